#!/bin/bash
# Builds libsqgpu.so in-tree for sm_100a (cross-compiles without a GPU). Usage: [OUT=name.so] ./build.sh [extra nvcc flags]
# (OUT + -D switches build the kernel-experiment variants that profiles/variants.py compares on the GPU)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo \
      -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function,-Wno-unknown-pragmas -shared -cudart static \
      -ccbin /usr/bin/g++ "$@" -o "${OUT:-libsqgpu.so}" sqgpu.cu
