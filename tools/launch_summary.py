"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per kernel and (optionally) per grid.
python tools/launch_summary.py gpurun_out/launches.csv [kernel-substring ...]
python tools/launch_summary.py gpurun_out/launches.csv --count REGEX     -> number of launches whose kernel name matches"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, vi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
if len(sys.argv) > 3 and sys.argv[2] == '--count':
    pat = re.compile(sys.argv[3])
    print(sum(1 for r in rows[start:] if len(r) > vi and pat.search(r[ki])))
    sys.exit(0)
agg, per_grid, tot = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0]
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
    per_grid[(name, r[gi])][0] += 1; per_grid[(name, r[gi])][1] += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{100 * t / tot:5.1f}%  {k:32s} n={n:4d} avg {t / n:7.1f} us  sum {t:8.1f}")
for pat in sys.argv[2:]:
    print("--", pat)
    for k, (n, t) in sorted(per_grid.items(), key=lambda x: -x[1][1]):
        if pat in k[0]:
            print(f"   {k[0]:24s} grid {k[1]:16s} n={n:3d} avg {t / n:7.1f} sum {t:7.1f}")
