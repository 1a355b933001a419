"""Print every launch of one training step (ncu launch-list csv) in order: index, us, cumulative us, grid, kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
gi = hdr.index('Grid Size')
recs = []
for r in data:
    if len(r) <= vi: continue
    t = float(r[vi].replace(',', ''))
    t = t / 1e3 if r[ui] == 'ns' else (t * 1e3 if r[ui] == 'ms' else t)
    recs.append((r[ki].split('(')[0].replace('void ', '').replace('b200seg::', ''), t, r[gi]))
marks = [i for i, (n, *_) in enumerate(recs) if n.startswith('softmax_dice_fwd')]
recs = recs[marks[-2]:marks[-1]]
st = [i for i, r in enumerate(recs) if r[0].startswith('pack_weights')][0]
recs = recs[st:] + recs[:st]
cum = 0
for i, (n, t, g) in enumerate(recs):
    cum += t
    print(f"{i:4d} {t:7.1f} {cum:7.0f} {g:>16s} {n[:70]}")
