"""Per-op device times of one eager training step (CUDA events around every tape op), grouped by op label.
python tools/profile_ops.py [B]   ->  gpurun_out/ops.tsv + a summary on stdout"""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
out = os.path.join(ROOT, "gpurun_out", "ops.tsv")
os.makedirs(os.path.dirname(out), exist_ok=True)
os.environ["UB_PROFILE_DUMP"] = out
import __graft_entry__ as ge
import unet_oracle as O
ub = ge.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = O.UNetConfig()
tr = ub.Trainer(B=B)
tr.set_params(O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy())
x = (torch.rand(B, 3, 64, 64) * 2 - 1).cuda()
for _ in range(3):
    tr.train_step_device(x.data_ptr())
tr.sync()
prof = tr.profile(reps=3)
print({k: round(v["ms"], 3) for k, v in prof.items() if isinstance(v, dict)}, "total", round(prof["total_ms"], 3))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = {"fwd": 0.0, "bwd": 0.0}
for ln in open(out):
    ph, i, kind, us, fl, by, label = ln.rstrip("\n").split("\t")
    key = (ph, label or f"kind{kind}")
    agg[key][0] += 1; agg[key][1] += float(us); agg[key][2] += float(fl)
    tot[ph] += float(us)
print("eager fwd us", round(tot["fwd"]), "bwd us", round(tot["bwd"]))
for (ph, label), (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    tf = f"{fl / us / 1e6:7.1f} TF/s" if fl > 0 else ""
    print(f"{us:8.1f} us  n={n:3d} avg {us / n:6.1f}  {ph} {label} {tf}")
tr.close()
