"""A few eager (no CUDA graph) training steps at B=32 for ncu: python tools/profile_step.py [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge
import unet_oracle as O
ub = ge.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = O.UNetConfig()
flat = O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy()
tr = ub.Trainer(B=32, use_cuda_graph=0)
tr.set_params(flat)
x = (torch.rand(32, 3, 64, 64) * 2 - 1).cuda()
for _ in range(steps):
    tr.train_step_device(x.data_ptr())
tr.sync()
print("loss", tr.last_loss(), "launches/step", tr.launches_per_step())
tr.close()
