#!/usr/bin/env python
"""Per-kernel shares of one predict from an ncu launch list (gpu__time_duration.sum CSV).  usage: launch_shares.py file.csv [which]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); idc = h.index('ID')
L = [(int(r[idc]), r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv and r[mv]]
names = [x[1] for x in L]
starts = [i for i, n in enumerate(names) if n.startswith('set_y_kernel')]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3          # 4th predict = the timed step after 3 warm-ups
seg = L[starts[which]:starts[which + 1]] if which + 1 < len(starts) else L[starts[which]:]
# cut at the first kernel of the next problem build (e2e phase)
cut = [i for i, x in enumerate(seg) if x[1].startswith('lattice_ids')]
if cut: seg = seg[:cut[0]]
agg = collections.OrderedDict()
for _, n, t in seg:
    n = n.split('(')[0].replace('void ', '')
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (n[:80], a[0], a[1] / 1e3, 100 * a[1] / tot))
print("total %.3f ms, %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
