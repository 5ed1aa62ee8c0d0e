#!/bin/bash
# Round-end measurement set (run under gpurun on one B200): GPU tests, bench lines, per-layer table, ncu launch list,
# ncu --set full captures IN THE MODEL of the dominant tensor-bound family, a narrow HBM-bound conv2 and the WN stack launch, smoke().
# Afterwards, on the CPU box: python tools/make_traffic_json.py <tag>  (profiles/dominant_kernel_traffic.json)
#   tools/round_records.sh <tag, e.g. r2>
cd "$(dirname "$0")/.."
T=${1:-r2}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $O/${T}_gpu_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --dump-profile $O/${T}_layers.csv 2>$O/${T}_bench_n1.err | tail -1 > $O/${T}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 40 2>/dev/null | tail -1 > $O/${T}_reference_arm.json
timeout 300 python bench.py --impl torch-gpu --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/${T}_torch_gpu_arm.json
timeout 300 python bench.py --engine bf16 --batch-per-gpu 64 --frames 512 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $O/${T}_bench_bf16_c4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv python tools/ncu_target.py --iters 2 > $O/${T}_ncu_list.log 2>&1
S=$(python tools/ncu_inmodel.py --layer resblock_conv1 --cin 128 --k 11 --nth 0 2>/dev/null)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $S -c 6 -f -o $O/${T}_c128_k11_block python tools/ncu_target.py > $O/${T}_ncu_c128.log 2>&1
S=$(python tools/ncu_inmodel.py --layer resblock_conv2 --cin 64 --k 7 --nth 0 2>/dev/null)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $S -c 1 -f -o $O/${T}_narrow_conv2 python tools/ncu_target.py > $O/${T}_ncu_narrow.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wn_layer_kernel -s 0 -c 1 -f -o $O/${T}_wn_layer python tools/ncu_target.py > $O/${T}_ncu_wn.log 2>&1
timeout 120 python tools/mel_bench.py > $O/${T}_mel_bench.json 2>/dev/null
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/${T}_smoke.log
cat $O/${T}_gpu_tests.log $O/${T}_smoke.log
python - <<P
import json
for f in ("bench_n1","reference_arm","torch_gpu_arm","bench_bf16_c4"):
    try:
        d=json.load(open("$O/${T}_%s.json"%f)); print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e: print(f, "ERR", e)
P
