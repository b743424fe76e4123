"""Device timeline of the captured training step (the CUDA graph as it runs, branches overlapped): CUPTI kernel
activity records through torch.profiler (no nsys in the image).

    python tools/timeline.py [B] [out.tsv]

Writes one row per kernel of ONE steady-state step (start relative to the step's first kernel, duration, stream, grid,
name) and prints: step span, busy time per stream, the main stream's kernels by name (sum / count / average) and its
idle gaps, and -- when data parallel -- where the NCCL kernels sit.  Times are CUPTI's (concurrent-kernel tracing adds
~1 us per launch; compare shares, the bench is the step time)."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge  # noqa: E402
import unet_oracle as O  # noqa: E402

ub = ge.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "timeline.tsv")
# under torchrun: data parallel over NCCL, rank 0 reports (the all-reduce kernels show up as ncclDevKernel_*)
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
cfg = O.UNetConfig()
tr = ub.Trainer(B=B, device=local_rank, seed=1234 + rank)
tr.set_params(O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy())
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    idt = torch.zeros(ub.UB_NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(ub.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    tr.attach_dp(rank, world, bytes(idt.cpu().numpy().tobytes()))
x = (torch.rand(B, 3, 64, 64) * 2 - 1).cuda()
for _ in range(5):
    tr.train_step_device(x.data_ptr())
tr.sync()
NSTEP = 3
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        tr.train_step_device(x.data_ptr())
    tr.sync()
    torch.cuda.synchronize()
import json  # noqa: E402
import tempfile  # noqa: E402

if rank != 0:
    tr.close()
    sys.exit(0)
tmp = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(tmp)
evs = []
for e in json.load(open(tmp))["traceEvents"]:
    if e.get("ph") == "X" and e.get("cat", "").lower() in ("kernel", "gpu_memcpy", "gpu_memset"):
        evs.append((float(e["ts"]), float(e["dur"]), int(e.get("args", {}).get("stream", -1)), e["name"]))
evs.sort()
kern = [e for e in evs if "memcpy" not in e[3].lower() and "memset" not in e[3].lower()]
per = len(kern) // NSTEP
print(f"{len(evs)} device activities, {len(kern)} kernels over {NSTEP} steps ({per} per step)")
# the last step: from its first kernel (diffusion_t / prepare) to the end
names = [k[3] for k in kern]
starts = [i for i, n in enumerate(names) if "diffusion_t_kernel" in n]
i0 = starts[-1] if starts else len(kern) - per
step = kern[i0:]
t0 = step[0][0]
span = max(s + d for s, d, _, _ in step) - t0
with open(out, "w") as f:
    for s, d, st, n in step:
        f.write(f"{s - t0:.2f}\t{d:.2f}\t{st}\t{n[:90]}\n")
print(f"step span {span:.1f} us, {len(step)} kernels -> {out}")
by_stream = collections.defaultdict(list)
for s, d, st, n in step:
    by_stream[st].append((s - t0, d, n))
main = max(by_stream, key=lambda k: sum(d for _, d, _ in by_stream[k]))
for st, lst in sorted(by_stream.items(), key=lambda kv: -sum(d for _, d, _ in kv[1])):
    print(f"  stream {st}: {len(lst):4d} kernels, busy {sum(d for _, d, _ in lst):8.1f} us, first {lst[0][0]:8.1f} last end "
          f"{max(s + d for s, d, _ in lst):8.1f}" + ("   <- main" if st == main else ""))


def short(n):
    n = n.replace("void ", "").replace("ub::", "").replace("(anonymous namespace)::", "")
    return n.split("(")[0][:48]


agg = collections.defaultdict(lambda: [0, 0.0])
for s, d, n in by_stream[main]:
    agg[short(n)][0] += 1
    agg[short(n)][1] += d
print("main stream by kernel:")
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {d:8.1f} us  n={c:3d} avg {d / c:6.1f}  {n}")
lst = sorted(by_stream[main])
gaps = []
for (s0, d0, n0), (s1, d1, n1) in zip(lst, lst[1:]):
    g = s1 - (s0 + d0)
    if g > 0:
        gaps.append((g, s0 + d0, short(n0), short(n1)))
print(f"main stream idle between kernels: {sum(g for g, *_ in gaps):.1f} us in {len(gaps)} gaps (overlap counts as 0); largest:")
for g, at, a, b in sorted(gaps, reverse=True)[:12]:
    print(f"  {g:7.1f} us at {at:8.1f}  after {a}  before {b}")
nc = [(s_, d_, n_) for st_, lst_ in by_stream.items() for s_, d_, n_ in lst_ if "nccl" in n_.lower()]
if nc:
    print(f"NCCL kernels: {len(nc)}, busy {sum(d for _, d, _ in nc):.1f} us; start / duration / end (us from step start):")
    for s_, d_, n_ in sorted(nc):
        print(f"  {s_:8.1f} {d_:7.1f} {s_ + d_:8.1f}  {n_[:60]}")
ov = sum(max(0.0, (s0 + d0) - s1) for (s0, d0, _), (s1, _, _) in zip(lst, lst[1:]))
print(f"main stream kernel-to-kernel overlap (programmatic dependent launch): {ov:.1f} us")
tr.close()
