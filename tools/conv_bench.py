"""conv2d_k3 forward / dgrad / wgrad microbenchmark (BASELINE.json configs[1]),
through the C ABI.  Two numbers per shape and pass:
  core : the tcgen05 kernels on operands resident in their native layout (NHWC bf16, packed weights)
  api  : the drop-in NCHW fp32 operators (ub_conv2d_k3_forward3 / backward2) including the layout conversions
Timing: CUDA events, >= 10 warm-up, L2 flushed (256 MB memset) between repetitions as dev/common.h:99-107 does.

    python tools/conv_bench.py [--quick] [--json out.json] [--only C,H] [--reps 30]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ub = ge.load_package()
L = ub.lib()

REAL = [(64, 64, 64), (64, 128, 32), (128, 128, 32), (128, 192, 16), (192, 192, 16), (192, 256, 8), (256, 256, 8),
        (512, 256, 8), (448, 256, 8), (448, 192, 16), (384, 192, 16), (320, 192, 16), (320, 128, 32), (256, 128, 32),
        (192, 128, 32), (192, 64, 64), (128, 64, 64)]  # (Cin, Cout, H=W) of the default U-Net


def p(t):
    return C.c_void_p(t.data_ptr())


def timeit(fn, reps, flush):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def bench_shape(B, Cin, Cout, H, reps, flush, peaks):
    W = H
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device="cuda")
    dout = torch.randn(B, Cout, H, W, device="cuda")
    out = torch.empty(B, Cout, H, W, device="cuda")
    dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
    xb = torch.empty(B * H * W * Cin, dtype=torch.bfloat16, device="cuda")
    dyb = torch.empty(B * H * W * Cout, dtype=torch.bfloat16, device="cuda")
    ob = torch.empty(B * H * W * Cout, dtype=torch.bfloat16, device="cuda")
    dxb = torch.empty(B * H * W * Cin, dtype=torch.bfloat16, device="cuda")
    wf = torch.empty(9 * Cout * Cin, dtype=torch.bfloat16, device="cuda")
    wd = torch.empty(9 * Cout * Cin, dtype=torch.bfloat16, device="cuda")
    assert L.ub_nchw_to_nhwc_bf16(p(x), p(xb), B, Cin, H, W) == 0
    assert L.ub_nchw_to_nhwc_bf16(p(dout), p(dyb), B, Cout, H, W) == 0
    assert L.ub_pack_conv_weight(p(w), p(wf), p(wd), Cin, Cout, 3) == 0
    flops = 2.0 * 9 * B * H * W * Cin * Cout
    byts = 2.0 * B * H * W * (Cin + Cout) + 2.0 * 9 * Cin * Cout   # bf16 operands, each tensor touched once
    res = {"B": B, "Cin": Cin, "Cout": Cout, "H": H, "gflop": flops / 1e9}
    runs = {
        "fwd_core": lambda: L.ub_conv2d_nhwc_forward(p(xb), p(wf), p(b), p(ob), B, H, W, Cin, Cout, 3),
        "dgrad_core": lambda: L.ub_conv2d_nhwc_dgrad(p(dyb), p(wd), p(dxb), B, H, W, Cin, Cout, 3),
        "wgrad_core": lambda: L.ub_conv2d_nhwc_wgrad(p(dyb), p(xb), p(dw), p(db), B, H, W, Cin, Cout, 3),
        "fwd_api": lambda: L.ub_conv2d_k3_forward3(p(x), p(w), p(b), p(out), B, Cin, Cout, H, W),
        "bwd_api": lambda: L.ub_conv2d_k3_backward2(p(dout), p(x), p(w), None, None, p(dx), p(dw), p(db), B, Cin, Cout,
                                                    H, W),
    }
    for name, fn in runs.items():
        if name == "wgrad_core" and (Cin % 64 or Cout % 64):
            continue
        ms = timeit(fn, reps, flush)
        f = flops * (2 if name == "bwd_api" else 1)
        tf = f / ms / 1e9
        t_tensor = f / (peaks["tf"] * 1e12) * 1e3
        t_hbm = byts * (2 if name == "bwd_api" else 1) / (peaks["hbm"] * 1e9) * 1e3
        res[name] = {"ms": round(ms, 4), "tflops": round(tf, 1), "frac_of_tensor_peak": round(tf / peaks["tf"], 3),
                     "frac_of_min_roofline": round(max(t_tensor, t_hbm) / ms, 3)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="Cin,Cout,H")
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    peaks = {"tf": pk["bf16_tflops"], "hbm": pk["hbm_gbs"]}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if a.only:
        shapes = [tuple(int(v) for v in a.only.split(","))]
    elif a.quick:
        shapes = [(64, 64, 64), (192, 64, 64), (128, 128, 32), (192, 192, 16), (256, 256, 8)]
    else:
        shapes = [(c, c, h) for c in (64, 128, 192, 256, 384, 512) for h in (8, 16, 32, 64)]
        shapes += [s for s in REAL if s not in shapes]
    out = []
    print(f"{'shape':>22} | " + " | ".join(f"{n:>22}" for n in ("fwd_core", "dgrad_core", "wgrad_core", "fwd_api", "bwd_api")))
    for (ci, co, h) in shapes:
        r = bench_shape(a.batch, ci, co, h, a.reps, flush, peaks)
        out.append(r)
        cells = []
        for n in ("fwd_core", "dgrad_core", "wgrad_core", "fwd_api", "bwd_api"):
            cells.append(f"{r[n]['ms']:8.4f}ms {r[n]['tflops']:7.1f}TF" + f" {r[n]['frac_of_min_roofline']:4.2f}"
                         if n in r else " " * 22)
        print(f"B{a.batch} {ci:>3}->{co:<3} @{h:>2}x{h:<2}     | " + " | ".join(cells), flush=True)
    if a.json:
        json.dump({"peaks": peaks, "note": "frac_of_min_roofline = max(FLOPs/peak_bf16, bf16 bytes/HBM peak) / measured",
                   "results": out}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
