"""Development check (GPU): full training step of the CUDA trainer vs the CPU oracle, then a timing loop.
Usage: python tools/trainer_check.py [--B 4] [--steps 10] [--time-B 32]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge  # noqa: E402
import unet_oracle as O  # noqa: E402

ub = ge.load_package()


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--time-B", type=int, default=32)
    ap.add_argument("--time-steps", type=int, default=30)
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())

    cfg = O.UNetConfig()
    P = O.init_params(cfg, seed=0)
    flat = O.flatten_params(cfg, P)
    x0, t, noise = O.synthetic_batch(cfg, a.B)

    tr = ub.Trainer(B=a.B, use_cuda_graph=0 if a.no_graph else 1)
    print("nparams", tr.nparams, "launches/step", tr.launches_per_step())
    assert tr.nparams == flat.numel()
    tr.set_params(flat.numpy())

    t0 = time.time()
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise)
    print(f"oracle fwd+bwd {time.time() - t0:.2f}s  loss {float(loss_ref):.6f}")
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out = tr.get_output()
    g = tr.get_grads()
    print(f"cuda loss {loss:.6f}  |dloss| {abs(loss - float(loss_ref)):.3e}")
    print(f"out rel-inf err {rel(out, out_ref.numpy()):.3e}")
    g_ref = g_ref.numpy()
    print(f"grads: global rel-inf {rel(g, g_ref):.3e}  rel-L2 {np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref):.3e}"
          f"  cos {float(g @ g_ref / (np.linalg.norm(g) * np.linalg.norm(g_ref))):.6f}")
    off = 0
    worst = []
    for name, shape in O.param_spec(cfg):
        n = int(np.prod(shape))
        gr, gm = g_ref[off:off + n], g[off:off + n]
        worst.append((rel(gm, gr), name, float(np.abs(gr).max())))
        off += n
    worst.sort(reverse=True)
    for w in worst[:12]:
        print("   worst tensor rel-inf %.3e  %-40s ref max %.3e" % w)
    nan = int(np.isnan(g).sum())
    print("nan grads:", nan)

    # loss trace over n steps with injected inputs
    batches = [O.synthetic_batch(cfg, a.B, seed=1234 + i) for i in range(a.steps)]
    t0 = time.time()
    losses_ref, flat_ref = O.train_steps(cfg, flat, batches, lr=1e-4)
    print(f"oracle {a.steps} steps {time.time() - t0:.1f}s")
    tr.set_params(flat.numpy())
    tr2 = tr
    losses = [tr2.train_step(b[0].numpy(), b[1].numpy(), b[2].numpy(), lr=1e-4) for b in batches]
    for i, (lr_, lc) in enumerate(zip(losses_ref, losses)):
        print(f"  step {i + 1}: oracle {lr_:.6f}  cuda {lc:.6f}  diff {lc - lr_:+.2e}")
    p = tr.get_params()
    print(f"params after {a.steps} steps: rel-inf {rel(p, flat_ref.numpy()):.3e}, max |dp| "
          f"{np.abs(p - flat_ref.numpy()).max():.3e}")
    tr.close()

    # timing
    if a.time_B:
        tr = ub.Trainer(B=a.time_B, use_cuda_graph=0 if a.no_graph else 1)
        tr.set_params(flat.numpy())
        xb = torch.rand(a.time_B, 3, 64, 64).mul_(2).sub_(1).pin_memory()
        xd = xb.cuda()
        for _ in range(5):
            tr.train_step_device(xd.data_ptr())
        tr.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(tr.stream())
        e0.record(st)
        for _ in range(a.time_steps):
            tr.train_step_device(xd.data_ptr())
        e1.record(st)
        tr.sync()
        ms = e0.elapsed_time(e1) / a.time_steps
        print(f"B={a.time_B}: {ms:.3f} ms/step device-resident -> {a.time_B / ms * 1e3:.0f} img/s "
              f"(loss {tr.last_loss():.4f})")
        t0 = time.time()
        for _ in range(a.time_steps):
            tr.train_step(xb.numpy(), want_loss=True)
        ms2 = (time.time() - t0) / a.time_steps * 1e3
        print(f"B={a.time_B}: {ms2:.3f} ms/step host->device->loss e2e -> {a.time_B / ms2 * 1e3:.0f} img/s")
        tr.close()


if __name__ == "__main__":
    main()
