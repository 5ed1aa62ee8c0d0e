#!/bin/bash
# Like tools/ab.sh for any number of values of one environment switch:  tools/ab3.sh <VAR> <rounds> <v1> <v2> [...]
cd "$(dirname "$0")/.."
V=$1; R=$2; shift 2
for r in $(seq $R); do
  for x in "$@"; do
    env $V=$x python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V=$x', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['ms_per_step'],3))"
  done
done | tee /tmp/ab.log
python - <<P
import collections
s=collections.defaultdict(list)
for l in open('/tmp/ab.log'):
    k,ms,clk,e2e=l.split(); s[k].append((float(ms),float(clk),float(e2e)))
for k,v in s.items(): print(k,'mean ms %.3f  clk %.0f  e2e ms %.3f'%tuple(sum(x[i] for x in v)/len(v) for i in range(3)))
P
