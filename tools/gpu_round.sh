#!/usr/bin/env bash
# one gpurun call: parity tests, A/B bench lines (finalize kernel k5 vs k4), kernel-only timings
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for fin in k5 k4; do
  VPDQ_B200_FINALIZE=$fin timeout 300 python bench.py --steps 10 --warmup 3 --no-hamming --no-luma --no-cpu-baseline > gpurun_out/bench_fin_$fin.json 2> gpurun_out/bench_fin_$fin.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_fin_$fin.json"))
print("$fin", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "kx ms", round(d["roofline"]["kernel_ms_per_launch"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]))
PY
done
python tools/kx_bench.py 4096 10
