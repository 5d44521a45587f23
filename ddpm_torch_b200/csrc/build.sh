#!/bin/bash
# Build libddpm_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libddpm_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -shared -Xcompiler -fPIC \
    -Xptxas -v -o $OUT ddpm_b200.cu -lcudart 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning|spill" build.log | grep -v "0 bytes spill" | head -20 || true
echo "built $(realpath $OUT)"
