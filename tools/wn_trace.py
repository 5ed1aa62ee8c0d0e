"""Timeline of the multi-layer WN launch from a $SVK_WN_TRACE dump (clock64() stamps per tile and layer; wn_layer.cu)."""
import csv
import sys

import numpy as np

rows = list(csv.reader(open(sys.argv[1])))[1:]
a = np.array(rows, dtype=np.int64)
nt, nl = a[:, 0].max() + 1, a[:, 1].max() + 1
E = a[:, 2:].reshape(nt, nl, -1)
names = {0: 'tma: layer start', 1: 'tma: own x stored', 2: 'tma: neighbour flags ok', 3: 'tma: loads issued', 4: 'mma: in_layer start',
         5: 'mma: in_layer issued', 6: 'mma: res_skip acts ready', 7: 'mma: res_skip issued', 8: 'epi: gate 0 acc ready',
         14: 'epi: gate 1 acc ready', 13: 'epi: gate 2 acc ready', 9: 'epi: gates done', 10: 'epi: res_skip 0 acc ready',
         11: 'epi: layer done', 12: 'epi: tile published'}
per = E[:, 1:, 4] - E[:, :-1, 4]
print(f'{nl} layers x {nt} tiles; layer period (in_layer start -> next in_layer start): mean {per.mean():.0f} min {per.min()} max {per.max()} cycles')
for ev in [4, 8, 14, 13, 5, 9, 6, 10, 7, 12, 11]:
    rel = E[:, 1:nl - 1, ev] - E[:, 1:nl - 1, 4]
    print(f'{names[ev]:28s} mean {rel.mean():8.0f}  p10 {np.percentile(rel, 10):8.0f}  p90 {np.percentile(rel, 90):8.0f}')
for ev in [1, 2, 4]:
    rel = E[:, 2:nl, ev] - E[:, 1:nl - 1, 4]
    print(f'next layer {names[ev]:17s} mean {rel.mean():8.0f}  p10 {np.percentile(rel, 10):8.0f}  p90 {np.percentile(rel, 90):8.0f}')

if E.shape[2] >= 32:  # per-job stamps of the first epilogue warp: N-tile mt, job k -> accumulators loaded / results stored
    for mt in range(2):
        for k in range(4):
            a0, a1 = 16 + mt * 8 + 2 * k, 17 + mt * 8 + 2 * k
            if (E[:, 1:nl - 1, a0] == 0).all():
                continue
            r0, r1 = E[:, 1:nl - 1, a0] - E[:, 1:nl - 1, 4], E[:, 1:nl - 1, a1] - E[:, 1:nl - 1, 4]
            print(f'epi warp 4: res_skip N-tile {mt} job {k}: accumulators in registers {r0.mean():8.0f}   stored {r1.mean():8.0f}')
