"""B=4 vs 8x-replicated B=32 gradients, per parameter tensor (finds the layer a batch-size dependent bug lives in).
UB_DEBUG_SYNC=1 python tools/debug_b32.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge
import unet_oracle as O
ub = ge.load_package()
cfg = O.UNetConfig()
flat = O.flatten_params(cfg, O.init_params(cfg, seed=0))
x0, t, noise = O.synthetic_batch(cfg, 4)
tr4 = ub.Trainer(B=4, use_cuda_graph=0); tr4.set_params(flat.numpy())
l4 = tr4.forward_backward(x0.numpy(), t.numpy(), noise.numpy()); g4 = tr4.get_grads(); tr4.close()
rep = lambda a: np.concatenate([a.numpy()] * 8, axis=0)
tr = ub.Trainer(B=32, use_cuda_graph=0); tr.set_params(flat.numpy())
l32 = tr.forward_backward(rep(x0), rep(t), rep(noise)); g32 = tr.get_grads()
print("loss", l4, l32, "grad rel", np.linalg.norm(g32 - g4) / np.linalg.norm(g4))
off = 0; rows = []
for name, shape in O.param_spec(cfg):
    n = int(np.prod(shape)); a, b = g32[off:off + n], g4[off:off + n]; off += n
    rows.append((np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30), name, shape))
for i, (r, name, shape) in enumerate(rows):
    if r > 2e-2: print(f"{i:4d} {r:9.3e} {name} {shape}")
print("worst", sorted(rows, reverse=True)[:5])
try:
    tr.train_step(rep(x0), rep(t), rep(noise)); print("eager train_step ok")
except Exception as e:
    print("train_step:", e)
tr.close()
tr = ub.Trainer(B=32, use_cuda_graph=1); tr.set_params(flat.numpy())
try:
    print("graph step loss", tr.train_step(rep(x0), rep(t), rep(noise)))
except Exception as e:
    print("graph train_step:", e)
