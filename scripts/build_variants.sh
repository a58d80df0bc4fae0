#!/bin/bash
# Variant builds of libsphb200.so for tuning sweeps on the GPU box (selected with SPHB200_LIB=<path>).
#   scripts/build_variants.sh name "-DSPH_RING_NCW=12 ..." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p sphexample_b200/lib
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
      -o sphexample_b200/lib/libsphb200_$name.so sphexample_b200/csrc/sphb200.cu -ldl &
done
wait
ls -la sphexample_b200/lib/
