#!/bin/bash
# Build A/B variants of libfa_b200.so: scripts/build_variants.sh name1 "-DFLAGS" name2 "-DFLAGS" ...
# (scripts/gpu_ab.sh runs the harness once per variant with LD_LIBRARY_PATH pointing at it; variant "T*" = trace build)
cd "$(dirname "$0")/.."
V=flashattention.c_b200/variants
rm -rf $V; mkdir -p $V
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p $V/$name
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --shared $flags \
      -o $V/$name/libfa_b200.so flashattention.c_b200/csrc/fa_api.cu && echo "built $name ($flags)" ) &
done
wait
