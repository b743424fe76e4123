"""Summarise the ncu capture of the HBM-bound kernels (tools/measure_all.sh: <tag>_hbm_kernels.csv) per kernel and grid:
launches, average time, DRAM bytes (read + write) per launch, achieved DRAM GB/s against MEASURED_PEAKS.json, ncu's own
dram / L2 throughput percentages.   python tools/hbm_summary.py gpurun_out/<tag>_hbm_kernels.csv > profiles/..."""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6530.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
except Exception:
    pass
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ix = {k: hdr.index(k) for k in ('ID', 'Kernel Name', 'Grid Size', 'Block Size', 'Metric Name', 'Metric Unit', 'Metric Value')}
launch = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= ix['Metric Value']:
        continue
    d = launch.setdefault(r[ix['ID']], {'name': r[ix['Kernel Name']].split('(')[0], 'grid': r[ix['Grid Size']], 'block': r[ix['Block Size']]})
    v = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    if u in ('ns', 'nsecond'):
        v /= 1000.0
    elif u in ('ms', 'msecond'):
        v *= 1000.0
    elif u == 'Kbyte':
        v *= 1e3
    elif u == 'Mbyte':
        v *= 1e6
    elif u == 'Gbyte':
        v *= 1e9
    d[r[ix['Metric Name']]] = v
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault((d['name'], d['grid'], d['block']), collections.defaultdict(float))
    a['n'] += 1
    for k, v in d.items():
        if isinstance(v, float):
            a[k] += v
print(f"# ncu per-launch metrics of the HBM-bound kernels of one eager training step (B = 32), averaged per (kernel, grid);")
print(f"# GB/s = (dram read + write bytes) / gpu__time_duration; peak = {peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs); tensors the")
print(f"# producing kernel has just written are served by the 126 MB L2, so DRAM GB/s understates what the kernel moves: lts% = L2 throughput")
print(f"{'kernel':28s} {'grid':14s} {'blk':5s} {'n':>3s} {'us':>7s} {'dramMB':>8s} {'GB/s':>7s} {'of peak':>7s} {'dram%':>6s} {'lts%':>6s} {'L2 MB':>7s} {'warps%':>6s}")
tot_t = tot_b = 0.0
for (name, grid, block), a in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    n = a['n']
    t = a['gpu__time_duration.sum'] / n
    b = (a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']) / n
    tot_t += a['gpu__time_duration.sum']
    tot_b += a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']
    gbs = b / (t * 1e-6) / 1e9 if t > 0 else 0.0
    print(f"{name[:28]:28s} {grid.replace(' ', ''):14s} {block.split(',')[0].strip('('):5s} {int(n):3d} {t:7.1f} {b / 1e6:8.2f} {gbs:7.0f} {gbs / peak:7.2f} "
          f"{a['dram__throughput.avg.pct_of_peak_sustained_elapsed'] / n:6.1f} {a['lts__throughput.avg.pct_of_peak_sustained_elapsed'] / n:6.1f} "
          f"{a['lts__t_bytes.sum'] / n / 1e6:7.1f} {a['sm__warps_active.avg.pct_of_peak_sustained_active'] / n:6.1f}")
print(f"# all launches: {tot_t:.0f} us, {tot_b / 1e6:.0f} MB of DRAM traffic -> {tot_b / (tot_t * 1e-6) / 1e9:.0f} GB/s = {tot_b / (tot_t * 1e-6) / 1e9 / peak:.2f} of peak")
