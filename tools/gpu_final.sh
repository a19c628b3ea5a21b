#!/usr/bin/env bash
# final verification on a fresh box: build check, smoke, GPU parity suite, default bench line
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "kx frac", round(d["roofline"]["frac"], 4),
      "pipeline frac", round(d["roofline"]["pipeline"]["frac"], 4), "traffic", d["roofline"]["traffic"],
      "e2e", round(d["e2e"]["value"]), "cpu", round(d["cpu_baseline"]["value"]), "launches", d["gpu_launches"])
print("scan", d["hamming"]["scan"]["by_n_query"]["1"]["frac_of_hbm_peak"], "pairs", d["hamming"]["all_pairs"]["pair_comparisons_per_s"])
print("luma", d["luma_frames"]["frames_per_s_device_resident"], d["luma_frames"]["kernels"])
PY
