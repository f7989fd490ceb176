#!/bin/bash
# Developer tool: A/B builds of the attention translation unit with different compile-time knobs, linked against the
# other objects of the regular build.  usage: tools/build_attn_variants.sh name "-DFOO=1 ..." [name flags ...]
set -e
cd "$(dirname "$0")/.."
python -m tokensgen_b200.build > /dev/null
B=tokensgen_b200/build
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -cudart static $flags \
       -c tokensgen_b200/csrc/attn.cu -o $B/attn_$name.o
  nvcc -shared -cudart static -o tokensgen_b200/libtg_$name.so $B/common.o $B/gemm.o $B/attn_$name.o $B/elementwise.o $B/conv.o $B/vae.o
  echo built tokensgen_b200/libtg_$name.so
done
