#!/usr/bin/env bash
# one gpurun call: parity tests, A/B bench lines, ncu launch list + full capture of the dominant kernel
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for impl in fused2 fused; do
  VPDQ_B200_PDQ_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-hamming --no-luma --no-cpu-baseline > gpurun_out/bench_ab_$impl.json 2> gpurun_out/bench_ab_$impl.err
  cat gpurun_out/bench_ab_$impl.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fused2.csv python tools/prof_pdq.py fused2 > gpurun_out/prof_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kx_fused_jarosz2 -s 2 -c 1 -o gpurun_out/prof_fused2 -f python tools/prof_pdq.py fused2 > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log
