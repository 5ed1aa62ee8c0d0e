"""Top stall sites of an .ncu-rep source page (SASS view): developer helper."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
i = [k for k, l in enumerate(out) if l.startswith('"Address"')][0]
rows = list(csv.DictReader(out[i:]))
stalls = [c for c in rows[0].keys() if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r['# Samples'] or 0) for r in rows)
print('total samples', tot, 'instructions', len(rows))
agg = {}
for s in stalls:
    agg[s] = sum(int(r[s] or 0) for r in rows)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for idx, r in enumerate(rows): r['_i'] = idx
for r in sorted(rows, key=lambda r: -int(r['# Samples'] or 0))[:topn]:
    top = sorted(((int(r[s] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{r['_i']:5d} {int(r['# Samples']):7d} {100*int(r['# Samples'])/tot:5.1f}%  exec={r['Instructions Executed']:>9}  {r['Source'][:70]:70s} {top}")
