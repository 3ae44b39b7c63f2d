#!/bin/bash
# Build libcdrhead.so in-tree for sm_100a.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
OUT=${CDR_OUT:-../libcdrhead.so}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v"
SRCS="api.cu geometry.cu heatmap.cu layout.cu pack.cu gemm_ffma.cu gemm_tc.cu stem.cu backward.cu train_ops.cu"
BUILD=${CDR_BUILD_DIR:-build}
mkdir -p $BUILD
pids=()
for s in $SRCS; do
  $NVCC $FLAGS "$@" -c -o $BUILD/${s%.cu}.o $s > $BUILD/${s%.cu}.log 2>&1 &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then cat $BUILD/*.log; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $BUILD/*.o -lcudart_static -lpthread -ldl -lrt
echo "built $(realpath $OUT)"
