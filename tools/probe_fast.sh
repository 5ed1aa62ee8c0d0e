#!/bin/bash
# A/B of the lean STORE epilogue on the decoder ResBlock shapes (conv1 = image -> image, conv2 = image + residual -> fp32 + image):
#   tools/build_probe.sh && tools/probe_fast.sh      ($SVK_EPI_FAST=0 runs the generic epilogue)
cd "$(dirname "$0")/.."
P=tools/bin/tc_probe
run() { timeout 60 $P "$@" 2>&1 | grep -o "B=.*K=[0-9]* \|res=[01]\|maxabs [0-9.e+-]*\|bad [0-9]*\|[0-9.]* ms [0-9.]* TFLOP" | paste -sd' '; }
for f in 0 1; do echo "== SVK_EPI_FAST=$f"; export SVK_EPI_FAST=$f
for c in "0 tsn" "1 ts"; do set -- $c
run 16 32 32 3 1 262144 32 ffma $1 3 $2
run 16 32 32 7 1 262144 32 ffma $1 3 $2
run 16 32 32 11 1 262144 32 ffma $1 3 $2
run 16 64 64 3 1 131072 64 ffma $1 3 $2
run 16 64 64 7 1 131072 64 ffma $1 3 $2
run 16 64 64 11 1 131072 64 ffma $1 3 $2
run 16 128 128 3 1 65536 128 ffma $1 3 $2
run 16 128 128 11 1 65536 128 ffma $1 3 $2
done; done
