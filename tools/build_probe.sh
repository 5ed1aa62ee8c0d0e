#!/bin/bash
# Builds tools/bin/tc_probe[_PE] (developer probe for conv_tc.cu) for sm_100a.
# Variants: tc_probe = library defaults; tc_probe_PE = P producer warps, E epilogue warps.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I include"
C=smart-vocoder_b200/csrc
mkdir -p tools/bin $C/build
nvcc $F -c $C/conv_ffma.cu -o tools/bin/conv_ffma.o &
nvcc $F -c $C/conv_tc.cu -o tools/bin/conv_tc_default.o &
wait
nvcc $F tools/tc_probe.cu tools/bin/conv_tc_default.o tools/bin/conv_ffma.o -o tools/bin/tc_probe
for v in ${PROBE_VARIANTS:-}; do
  p=${v:0:1}; e=${v:1:1}
  nvcc $F -DSVK_TC_PROD_WARPS=$p -DSVK_TC_EPI_WARPS=$e -c $C/conv_tc.cu -o tools/bin/conv_tc_$v.o
  nvcc $F tools/tc_probe.cu tools/bin/conv_tc_$v.o tools/bin/conv_ffma.o -o tools/bin/tc_probe_$v
done
