#!/bin/bash
# ncu --set full captures, inside the model, of the narrow-stage kernels as they were BEFORE the second MMA issuer and before
# every C = 32 block went to the pair kernel (profiles/r2_ncu_*_before_dual_issue.txt).  To repeat them on the current build:
#   SVK_FUSE_PAIRS=1 SVK_DUAL_ISSUE=0 bash tools/cap_narrow.sh
cd /root/repo
O=gpurun_out
S=$(python tools/ncu_inmodel.py --layer resblock_pair --cin 32 --k 3 --nth 0 2>/dev/null)
echo "pair c32 k3 ordinal $S"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair_kernel -s $S -c 1 -f -o $O/r2_pair_c32_k3 python tools/ncu_target.py > $O/r2_ncu_pair.log 2>&1
S=$(python tools/ncu_inmodel.py --layer resblock_conv1 --cin 32 --k 7 --nth 0 2>/dev/null)
echo "conv1 c32 k7 ordinal $S"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $S -c 1 -f -o $O/r2_c32_k7_conv1 python tools/ncu_target.py > $O/r2_ncu_c32c1.log 2>&1
S=$(python tools/ncu_inmodel.py --layer resblock_conv1 --cin 64 --k 11 --nth 0 2>/dev/null)
echo "conv1 c64 k11 ordinal $S"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $S -c 1 -f -o $O/r2_c64_k11_conv1 python tools/ncu_target.py > $O/r2_ncu_c64c1.log 2>&1
ls -la $O/*.ncu-rep | tail -4
