#!/usr/bin/env bash
# one gpurun --gpus N call: NCCL parity test, bench at N (and at the smaller powers of two given), e2e dedupe at N and 1
# usage: bash tools/gpu_multi.sh N [smaller N ...]
set -u
N=${1:-2}
shift || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( timeout 600 python -m pytest tests/test_dist_nccl.py -q -m gpu 2>&1 | tail -4 )
for n in $N "$@"; do
  port=$((29500 + n))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}.json 2> gpurun_out/bench_n${n}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_n${n}.json"))
    sp = d["hamming"]["sharded_all_pairs"]
    print("N=${n} value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["e2e"]["copy_roof"]["frames/s_all_ranks"]),
          "api", {k: round(v["frames/s"]) for k, v in d["e2e"]["video_hasher_api"].items() if isinstance(v, dict)},
          "sharded pairs/s", f'{sp["pair_comparisons_per_s"]:.3e}', "gather us", round(sp["all_gather_us"], 1),
          "luma e2e", round(d["luma_frames"]["frames_per_s_e2e"]))
except Exception as e:
    print("N=${n} failed", e)
    print(open("gpurun_out/bench_n${n}.err").read()[-1500:])
PY
done
port=29611
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    tools/e2e_dedupe.py --clips 1000 --frames 300 --out gpurun_out/e2e_dedupe_n${N}.json 2> gpurun_out/e2e_dedupe_n${N}.err | cut -c1-900
timeout 900 python tools/e2e_dedupe.py --clips 1000 --frames 300 --out gpurun_out/e2e_dedupe_n1.json 2> gpurun_out/e2e_dedupe_n1.err | cut -c1-900
