#!/usr/bin/env bash
# one gpurun call: parity tests, the default bench line, the reference arm, ncu launch list + full captures
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "kx ms", round(d["roofline"]["kernel_ms_per_launch"], 4),
      "frac", round(d["roofline"]["frac"], 4), "pipeline frac", round(d["roofline"]["pipeline"]["frac"], 4),
      "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-hamming --no-luma --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kx_fused_jarosz2|k5_finalize' -s 4 -c 2 -o gpurun_out/prof_final_pdq2 -f python tools/prof_pdq.py fused2 > gpurun_out/prof_full.log 2>&1
tail -2 gpurun_out/prof_full.log
