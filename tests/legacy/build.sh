#!/usr/bin/env bash
# TEST-ONLY: builds tests/legacy/libvpdq_b200_legacy.so (the round-1 PDQ pipelines) for sm_100a.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-fvisibility=hidden,-O2 --shared -cudart static \
    -o libvpdq_b200_legacy.so pdq_lines.cu pdq_fused.cu pdq_fused2.cu legacy_support.cu
echo "built $(realpath libvpdq_b200_legacy.so)"
