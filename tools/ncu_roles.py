"""Split the stall samples of a conv_tc .ncu-rep by warp role (uses the barrier offsets each role waits on)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
i = [k for k, l in enumerate(out) if l.startswith('"Address"')][0]
rows = list(csv.DictReader(out[i:]))
stalls = [c for c in rows[0].keys() if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r['# Samples'] or 0) for r in rows)
print('total samples', tot, 'instrs', len(rows))
# key instructions
for k, r in enumerate(rows):
    s = r['Source']
    if ('SYNCS.PHASECHK' in s or 'UTCHMMA' in s or 'UBLKCP' in s or 'LDTM' in s) and int(r['Instructions Executed'] or 0) > 0:
        print(f"{k:5d} samples={r['# Samples']:>6} exec={r['Instructions Executed']:>9} {s[:80]}")
top = sorted(rows, key=lambda r: -int(r['# Samples'] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
idx = {id(r): k for k, r in enumerate(rows)}
for r in top:
    t2 = sorted(((int(r[s] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{idx[id(r)]:5d} {int(r['# Samples']):6d} {100*int(r['# Samples'])/tot:5.1f}% exec={r['Instructions Executed']:>9} {r['Source'][:64]:64s} {t2}")
