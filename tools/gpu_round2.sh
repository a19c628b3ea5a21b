#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --no-hamming --no-cpu-baseline > gpurun_out/bench_luma.json 2> gpurun_out/bench_luma.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_luma.json"))
print("value", round(d["value"]), "kx frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "e2e", round(d["e2e"]["value"]))
print(d.get("luma_frames"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(kx_|k[0-9]_|k_)' -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_final.csv | cut -c1-300
