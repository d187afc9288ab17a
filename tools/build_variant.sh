#!/bin/bash
# build one experimental variant of libcvr_b200.so into tools/ab_libs/<name>.so:  tools/build_variant.sh name [nvcc -D flags...]
# (A/B on the GPU box: copy it over cvr_b200/lib/libcvr_b200.so, run tools/kernel_ab.py, restore)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
src=${CVR_VARIANT_SRC:-cvr_b200/csrc}
unset CC CXX
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-fopenmp -Xptxas -v \
  -I include "$@" -shared -o tools/ab_libs/$name.so $src/cvr_api.cu $src/cvr_convert.cu $src/cvr_spmv.cu $src/cvr_check.cu $src/cvr_sharded.cu $src/cvr_mm_reader.cpp -lgomp -ldl 2> tools/ab_libs/$name.ptxas.log
grep -A2 "tile_kernelILi7ELi7ELb0" tools/ab_libs/$name.ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo
