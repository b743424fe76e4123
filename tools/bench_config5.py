"""BASELINE.json configs[4] on ONE GPU: 128x128 U-Net, channel_mult (1,1,2,3,4), attention at 16x16 and 8x8
(21 082 755 parameters), batch 64 per GPU, full training step.   python tools/bench_config5.py [B] [steps]"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge
import unet_oracle as O
ub = ge.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = O.UNetConfig(channel_mult=(1, 1, 2, 3, 4), attn_start_level=3, H=128, W=128)
tr = ub.Trainer(B=B, H=128, W=128, channel_mult=(1, 1, 2, 3, 4), att_start_level=3)
tr.set_params(O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy())
x = (torch.rand(B, 3, 128, 128) * 2 - 1).cuda()
for _ in range(5):
    tr.train_step_device(x.data_ptr())
tr.sync()
stream = torch.cuda.ExternalStream(tr.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    tr.train_step_device(x.data_ptr())
e1.record(stream)
tr.sync()
ms = e0.elapsed_time(e1) / steps
prof = tr.profile(reps=2)
print(json.dumps({"workload": "128x128 U-Net mult (1,1,2,3,4), attention at 16x16 and 8x8, full train step", "batch": B,
                  "ms_per_step": ms, "images_per_s": B / ms * 1e3, "loss": tr.last_loss(),
                  "model_tflops": B / ms * 1e3 * 88.54e9 / 1e12,
                  "classes_ms": {k: round(v["ms"], 3) for k, v in prof.items() if isinstance(v, dict)},
                  "mem_GB": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)}))
tr.close()
