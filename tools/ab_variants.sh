#!/bin/bash
# Build compile-time variants of libcdfgpu.so into variants/ (git-ignored .so files), for A/B timing on the GPU box:
#   tools/ab_variants.sh name1 "-DFOO=1" name2 "-DBAR -DBAZ=2" ...   then   CDFGPU_LIB=variants/libcdfgpu_name1.so python tools/gpu_probe.py
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 \
      -shared -cudart static -ccbin /usr/bin/g++ $flags -o variants/libcdfgpu_$name.so cdftools_b200/csrc/api.cu && echo "built $name" ) &
done
wait
