#!/bin/bash
# Builds tools/bin/tc_probe[_PE] (developer probe for conv_tc.cu) for sm_100a.
# Variants: tc_probe = library defaults; tc_probe_PE = P producer warps, E epilogue warps; tc_probe_tE = E epilogue
# warps in the TMA-input kernel.
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
  if [ "${v:0:1}" = "t" ]; then D="-DSVK_TC_EPI_WARPS_TMA=${v:1}"; else D="-DSVK_TC_PROD_WARPS=${v:0:1} -DSVK_TC_EPI_WARPS=${v:1:1}"; fi
  nvcc $F $D -c $C/conv_tc.cu -o tools/bin/conv_tc_$v.o
  nvcc $F tools/tc_probe.cu tools/bin/conv_tc_$v.o tools/bin/conv_ffma.o -o tools/bin/tc_probe_$v
done
nvcc $F tools/store_bench.cu -o tools/bin/store_bench
