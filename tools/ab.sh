#!/bin/bash
# A/B of an environment switch on the bench step, alternating runs so that box / clock drift hits both arms alike.
#   tools/ab.sh <VAR> <value A> <value B> [rounds] [extra bench args...]
# Prints ms_per_step (device-resident, unprofiled) and the SM clock under load for every run, then the means.
cd "$(dirname "$0")/.."
V=$1; A=$2; B=$3; R=${4:-3}; shift 4
for r in $(seq $R); do
  for x in "$A" "$B"; do
    env $V=$x python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | \
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
