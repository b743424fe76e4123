"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> per-kernel averages.
python tools/traffic_summary.py gpurun_out/<tag>_tc_traffic.csv profiles/<name>.json"""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, mi, vi, ui, idi = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}
d = collections.defaultdict(dict)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    d[r[idi]][r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1)
    d[r[idi]]['name'] = r[ki].split('(')[0].replace('void unnamed>::', '')
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for m in d.values():
    a = agg[m['name']]
    a[0] += 1; a[1] += m.get('gpu__time_duration.sum', 0); a[2] += m.get('dram__bytes_read.sum', 0); a[3] += m.get('dram__bytes_write.sum', 0)
out = {n: {"launches": c, "avg_us": round(t / c, 2), "dram_read_bytes_per_launch": round(r / c), "dram_write_bytes_per_launch": round(w / c)}
       for n, (c, t, r, w) in agg.items()}
out["_how"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the "
               "tcgen05 launches of one eager B=32 training step (tools/measure_all.sh); DRAM writes land in the 126 MB L2 "
               "and are written back after the kernel, so the write counter reads ~0")
# stamp: bench.py reports this capture only while the kernel sources are the ones it was taken from
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
out["_kernel_source_hash"] = bench.kernel_source_hash()
json.dump(out, open(sys.argv[2], 'w'), indent=1)
for n, v in out.items():
    print(n, v)
