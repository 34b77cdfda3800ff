#!/bin/sh
# Builds libbridgemc.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --shared \
      -o ../libbridgemc.so bridgemc.cu "$@"
