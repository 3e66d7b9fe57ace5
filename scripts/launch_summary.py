"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last step only).
usage: python scripts/launch_summary.py gpurun_out/launches.csv [nsteps]"""
import collections, csv, re, sys
path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
rows = rows[len(rows) - len(rows) // nsteps:]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in rows:
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    agg[name][0] += 1
    agg[name][1] += float(row['Metric Value'])
tot = sum(v[1] for v in agg.values())
print("launches %d  total %.3f ms (cold-cache, serialised)" % (len(rows), tot / 1e6))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-52s n=%4d %9.3f ms %5.1f%%" % (k[:52], v[0], v[1] / 1e6, 100 * v[1] / tot))
