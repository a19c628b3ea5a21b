#!/usr/bin/env bash
# Builds libvpdq_b200.so IN-TREE for sm_100a (B200).  nvcc cross-compiles without a GPU.
#   -fmad=false : no implicit FMA contraction anywhere (bit-exactness); the kernels use
#                 explicit __fmaf_rn where a fused op is provably exact.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libvpdq_b200.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false
       -Xcompiler -fPIC,-fvisibility=hidden,-O2 --shared -cudart static)
if [[ "${VPDQ_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
# dev: extra -D switches and an alternative output path (tools/kx_libs.py times several builds side by side)
if [[ -n "${VPDQ_EXTRA_DEFS:-}" ]]; then FLAGS+=(${VPDQ_EXTRA_DEFS}); fi
OUT=${VPDQ_OUT:-$OUT}
"$NVCC" "${FLAGS[@]}" -o "$OUT" pdq_kernels.cu pdq_systolic.cu resize_kernels.cu hamming_kernels.cu capi.cu
echo "built $(realpath "$OUT")"
