#!/bin/bash
# Builds tools/bin/tc_probe (developer probe for conv_tc.cu) for sm_100a.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I include"
C=smart-vocoder_b200/csrc
mkdir -p tools/bin $C/build
nvcc $F -Xcompiler -fPIC -c $C/conv_tc.cu -o $C/build/conv_tc.o
[ -f $C/build/conv_ffma.o ] || nvcc $F -Xcompiler -fPIC -c $C/conv_ffma.cu -o $C/build/conv_ffma.o
nvcc $F tools/tc_probe.cu $C/build/conv_tc.o $C/build/conv_ffma.o -o tools/bin/tc_probe
