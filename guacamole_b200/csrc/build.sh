#!/bin/sh
# Builds libguac_b200.so (sm_100a only) in-tree next to the Python package.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared ${GUAC_NVCC_EXTRA} -o "$HERE/../libguac_b200.so" "$HERE/guac_api.cu" -lcudart -lnccl -lz
${CXX:-g++} -O2 -std=c++17 -fPIC -shared -pthread -o "$HERE/../libguac_synth.so" "$HERE/guac_synth.cpp"
