#!/bin/bash
# Runs the tcgen05 conv probe over correctness and timing shapes; each case in its own process
# under a timeout so a trap or hang in one case does not take the others down.
#   probe_suite.sh [all|perf|check|img] [binary suffix, e.g. _84]   (img = operand-image timing cases only)
cd "$(dirname "$0")/.."
P=tools/bin/tc_probe$2
mkdir -p gpurun_out
run() { timeout 120 $P "$@" 2>&1 | tail -16; rc=${PIPESTATUS[0]}; [ $rc -ne 0 ] && echo "  [exit $rc] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== $P"
if [ "$1" != "perf" ]; then
echo "== correctness (cpu fp64 reference)"
#   B Cin Cout K dil L N ref res reps
run 1 32 32 1 1 128 32 cpu 0 1
run 1 32 32 3 3 200 32 cpu 0 1
run 2 32 32 11 1 30000 32 cpu 1 1
run 1 64 64 7 5 300 64 cpu 1 1
run 2 64 64 7 1 20000 64 cpu 1 1
run 3 128 128 11 5 10000 128 cpu 1 1
run 1 256 256 7 3 520 128 cpu 1 1
run 1 192 512 7 1 300 128 cpu 0 1
run 2 96 192 1 1 9000 96 cpu 0 1
run 1 192 96 1 1 500 96 cpu 1 1
echo "== operand-image input (TMA) / output"
run 1 32 32 3 3 200 32 cpu 0 1 t
run 2 32 32 11 1 30000 32 cpu 1 1 ts
run 2 64 64 7 5 20000 64 cpu 1 1 tsn
run 3 128 128 11 5 10000 128 cpu 1 1 ts
run 1 256 256 7 3 520 128 cpu 1 1 tsn
run 2 96 192 1 1 9001 96 cpu 0 1 t
fi
if [ "$1" != "check" ] && [ "$1" != "img" ]; then
echo "== timing (ffma reference), decoder stage shapes at 16x1024 frames"
run 16 256 256 3 1 8192 128 ffma 1 3
run 16 256 256 7 3 8192 128 ffma 1 3
run 16 256 256 11 5 8192 128 ffma 1 3
run 16 128 128 3 1 65536 128 ffma 1 3
run 16 128 128 7 1 65536 128 ffma 1 3
run 16 128 128 11 1 65536 128 ffma 1 3
run 16 128 128 11 5 65536 128 ffma 1 3
run 16 64 64 3 1 131072 64 ffma 1 3
run 16 64 64 7 1 131072 64 ffma 1 3
run 16 64 64 11 1 131072 64 ffma 1 3
run 16 32 32 3 1 262144 32 ffma 1 3
run 16 32 32 7 1 262144 32 ffma 1 3
run 16 32 32 11 1 262144 32 ffma 1 3
run 16 192 384 5 1 1024 128 ffma 0 3
run 16 192 384 1 1 1024 128 ffma 1 3
fi
if [ "$1" != "check" ]; then
echo "== timing, operand-image path as the resblocks use it: conv1 = image in -> image out, conv2 = image in + residual -> fp32 + image"
for c in "0 tsn" "1 ts"; do set -- $c
run 16 128 128 3 1 65536 128 ffma $1 3 $2
run 16 128 128 11 1 65536 128 ffma $1 3 $2
run 16 64 64 3 1 131072 64 ffma $1 3 $2
run 16 64 64 11 1 131072 64 ffma $1 3 $2
run 16 32 32 3 1 262144 32 ffma $1 3 $2
run 16 32 32 7 1 262144 32 ffma $1 3 $2
run 16 32 32 11 1 262144 32 ffma $1 3 $2
done
fi
