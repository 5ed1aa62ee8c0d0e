#!/bin/bash
# Runs the tcgen05 conv probe over correctness and timing shapes; each case in its own process
# under a timeout so a trap or hang in one case does not take the others down.
cd "$(dirname "$0")/.."
P=tools/bin/tc_probe
mkdir -p gpurun_out
run() { timeout 120 $P "$@" 2>&1 | tail -16; rc=${PIPESTATUS[0]}; [ $rc -ne 0 ] && echo "  [exit $rc] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ "$1" != "perf" ]; then
echo "== correctness (cpu fp64 reference)"
#   B Cin Cout K dil L    N  nsub sep nw ref res reps
run 1 32  32   1 1   128  32  1 0 2 cpu 0 1
run 1 32  32   3 3   200  32  1 1 3 cpu 0 1
run 1 64  64   7 5   300  64  1 1 3 cpu 1 1
run 2 128 128  11 5  1000 128 2 1 4 cpu 1 1
run 1 256 256  7 3   520  256 1 1 4 cpu 1 1
run 1 192 512  7 1   300  128 1 1 4 cpu 0 1
fi
echo "== timing (ffma reference), decoder stage shapes at 16x1024 frames"
run 16 256 256 3  1 8192   128 1 1 4 ffma 1 3
run 16 256 256 7  3 8192   128 1 1 4 ffma 1 3
run 16 256 256 11 5 8192   128 1 1 4 ffma 1 3
run 16 256 256 11 5 8192   256 1 1 3 ffma 1 3
run 16 128 128 3  1 65536  128 1 1 4 ffma 1 3
run 16 128 128 7  1 65536  128 1 1 4 ffma 1 3
run 16 128 128 11 1 65536  128 1 1 4 ffma 1 3
run 16 128 128 11 5 65536  128 1 1 4 ffma 1 3
run 16 128 128 11 1 65536  128 2 1 4 ffma 1 3
run 16 64  64  3  1 131072 64  1 1 4 ffma 1 3
run 16 64  64  11 1 131072 64  1 1 4 ffma 1 3
run 16 64  64  11 1 131072 64  2 1 4 ffma 1 3
run 16 32  32  3  1 262144 32  1 1 4 ffma 1 3
run 16 32  32  11 1 262144 32  1 1 4 ffma 1 3
run 16 32  32  11 1 262144 32  2 1 4 ffma 1 3
