#!/usr/bin/env bash
# one gpurun call: smoke, GPU parity suite, default bench line (+ reference arm), ncu launch list
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "kx frac", round(d["roofline"]["frac"], 4),
      "pipeline frac", round(d["roofline"]["pipeline"]["frac"], 4), "e2e", round(d["e2e"]["value"]),
      "roof", round(d["e2e"]["copy_roof"]["frames/s_all_ranks"]), "cpu", round(d["cpu_baseline"]["value"]))
print("api", {k: (round(v["frames/s"]), v["bit_identical_to_device_path"]) for k, v in d["e2e"]["video_hasher_api"].items() if isinstance(v, dict)})
h = d["hamming"]
print("sharded", h["sharded_all_pairs"]["pair_comparisons_per_s"], "gather us", h["sharded_all_pairs"]["all_gather_us"])
print("scan", {k: round(v["frac_of_hbm_peak"], 3) for k, v in h["scan"]["by_n_query"].items()}, "pairs", h["all_pairs"]["pair_comparisons_per_s"])
print("video-like", {k: round(v["frac_of_hbm_peak"], 3) for k, v in h["video_like_db"]["scan_by_n_query"].items()}, h["video_like_db"]["all_pairs"])
print("luma", d["luma_frames"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cut -c1-300 gpurun_out/bench_reference.json
# ncu: launch list of the same bench command (cold-cache, serialised: shares must agree, not absolutes) + full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(kx_|k5_|k_)' -c 80 --csv --log-file gpurun_out/launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-hasher-api > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches_ncu.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kx_systolic_jarosz|k5_finalize' -s 4 -c 2 -o gpurun_out/prof_pdq -f python tools/prof_pdq.py 4096 > gpurun_out/prof_pdq.log 2>&1
tail -2 gpurun_out/prof_pdq.log
