#!/bin/bash
# tools/experiments.sh — build timing-experiment variants of the library (never shipped):
#   usage: tools/experiments.sh name1 "flags1" name2 "flags2" ...   ->  tiny_mc_b200/lib/exp/libtinymc_<name>.so
#   e.g.   tools/experiments.sh ppl4 "-DTMC_PPL=4 -DTMC_ONLY_BLOCK=256 -DTMC_DEFAULT_BLOCK_PRIVATE=256 -DTMC_DEFAULT_BLOCK_PLAIN=256"
set -e
mkdir -p tiny_mc_b200/lib/exp
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -cudart static -shared \
       $2 -o tiny_mc_b200/lib/exp/libtinymc_$1.so tiny_mc_b200/csrc/tmc_api.cu -ldl &
  shift 2
done
wait
