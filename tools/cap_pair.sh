#!/bin/bash
# ncu --set full of one fused conv pair (k = 7, C = 32) inside the model with every C = 32 block fused ($SVK_FUSE_PAIRS=3).
cd "$(dirname "$0")/.."
O=gpurun_out
export SVK_FUSE_PAIRS=3
S=$(python tools/ncu_inmodel.py --layer resblock_pair --cin 32 --k 7 --nth 0 2>/dev/null)
echo "pair c32 k7 ordinal $S"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair_kernel -s $S -c 1 -f -o $O/r2_pair_c32_k7 python tools/ncu_target.py > $O/r2_ncu_pair7.log 2>&1
tail -2 $O/r2_ncu_pair7.log
