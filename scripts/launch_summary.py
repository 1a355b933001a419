"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: isolates ONE training step
(between two consecutive softmax_dice_fwd launches) and prints per-kernel time shares."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
recs = []
for r in data:
    if len(r) <= vi: continue
    t = float(r[vi].replace(',', ''))
    t = t / 1e3 if r[ui] == 'ns' else (t * 1e3 if r[ui] == 'ms' else t)
    recs.append((r[ki].split('(')[0].replace('void ', '').replace('b200seg::', ''), t))
marks = [i for i, (n, _) in enumerate(recs) if n.startswith('softmax_dice_fwd')]
if len(marks) >= 2:
    recs = recs[marks[-2]:marks[-1]]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in recs:
    agg[n][0] += 1; agg[n][1] += t
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(recs)} launches, {tot:.1f} us summed kernel time (serialised, cold-cache under ncu)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  {k[:90]}")
