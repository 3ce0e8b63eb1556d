#!/bin/bash
# tools/experiments.sh — build timing-experiment variants of the library (never shipped):
#   1 = one shared atomic per event, 2 = conflict-free azimuth look-up, 3 = 3 Philox rounds
set -e
mkdir -p tiny_mc_b200/lib/exp
for e in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -cudart static -shared \
       -DTMC_EXPERIMENT=$e -o tiny_mc_b200/lib/exp/libtinymc_exp$e.so tiny_mc_b200/csrc/tmc_api.cu -ldl
done
