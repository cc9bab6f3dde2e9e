#!/bin/sh
# Builds an experimental libweedgpu with extra -D switches next to the product library:
#   tools/build_variant.sh <tag> [-DNAME=VALUE ...]   ->  multithreadedgameengine_b200/exp/libweedgpu_<tag>.so
# tools/ab_kernels.py --lib <tag> measures it.  Diagnostic only; nothing in the product loads these.
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p multithreadedgameengine_b200/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -shared -Xcompiler -fPIC "$@" \
  -o multithreadedgameengine_b200/exp/libweedgpu_$tag.so multithreadedgameengine_b200/csrc/weed_ctx.cu
