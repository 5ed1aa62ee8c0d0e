#!/bin/bash
# A/B of an environment switch on the per-launch table: alternating profiled runs, sums per layer family.
#   tools/ab_layers.sh <VAR> <value A> <value B> [rounds]
cd "$(dirname "$0")/.."
V=$1; A=$2; B=$3; R=${4:-2}
for r in $(seq $R); do
  for x in "$A" "$B"; do
    env $V=$x python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --dump-profile /tmp/abl_${x}_$r.csv >/dev/null 2>&1
  done
done
python - "$V" "$A" "$B" "$R" <<'P'
import csv, sys, collections
V, A, B, R = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
tab = {}
for x in (A, B):
    acc = collections.defaultdict(float)
    for r in range(1, R + 1):
        for row in csv.DictReader(open(f"/tmp/abl_{x}_{r}.csv")):
            acc[(row["layer"], row["cin"], row["k"])] += float(row["ms_per_step"]) / R
    tab[x] = acc
print(f"{'layer,cin,k':28s} {V}={A:>4s} {V}={B:>4s}   (ms per step, mean of {R} runs)")
for k in sorted(tab[A], key=lambda k: -tab[A][k]):
    print(f"{','.join(k):28s} {tab[A][k]:8.3f} {tab[B].get(k, 0):8.3f}")
print(f"{'total':28s} {sum(tab[A].values()):8.3f} {sum(tab[B].values()):8.3f}")
P
