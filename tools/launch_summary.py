"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (last 1/n-th of the launches = one step).
usage: launch_summary.py launches.csv [n_steps] [--list]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
data = [r for r in rows[hi + 1:] if len(r) > vi]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1
per = len(data) // n
step = data[-per:]
agg = collections.OrderedDict()
for r in step:
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('scd::', '')
    us = float(r[vi].replace(',', '')) / 1e3
    if '--list' in sys.argv:
        print("%-46s %9.1f us grid %s" % (name[:46], us, r[gi]))
    agg.setdefault(name, [0.0, 0]); agg[name][0] += us; agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-58s %9.1f us %4d  %5.1f%%" % (k[:58], v[0], v[1], 100 * v[0] / tot))
print("total %.1f us over %d launches" % (tot, len(step)))
