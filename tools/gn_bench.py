"""GroupNorm(+SiLU) NHWC bf16 microbenchmark through the C ABI: single-pass slab kernels (impl 0) vs the two-pass
kernels (impl 1), forward and backward, at the U-Net's (channels, resolution) pairs, B = 32.

Timing: CUDA events around `reps` back-to-back calls (the tensors of one shape total <= 70 MB: L2 resident, as inside the
training step where the producer conv has just written them) and, with --flush, a 256 MB memset between calls (HBM).
GB/s = algorithmic bytes (fwd: 1R + 1W, bwd: 2R + 1W) / time.

    python tools/gn_bench.py [--flush] [--reps 50]
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ub = ge.load_package()
L = ub.lib()

SHAPES = [(64, 64), (128, 64), (192, 64), (128, 32), (256, 32), (320, 32), (192, 16), (384, 16), (448, 16), (256, 8),
          (512, 8)]


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, reps, flush):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush is None:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    B, G = a.batch, 32
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if a.flush else None
    print(f"# B={B}, {'HBM (L2 flushed)' if a.flush else 'L2 resident'}; us per call and algorithmic GB/s")
    print(f"{'shape':>14} | {'fwd slab':>16} | {'fwd 2-pass':>16} | {'bwd slab':>16} | {'bwd 2-pass':>16}")
    for Cc, H in SHAPES:
        x = torch.randn(B, H, H, Cc, device="cuda").bfloat16()
        dy = torch.randn(B, H, H, Cc, device="cuda").bfloat16()
        y, dx = torch.empty_like(x), torch.empty_like(x)
        gam, bet = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
        cs, scr = torch.zeros(B, Cc, 2, device="cuda"), torch.zeros(B, Cc, 2, device="cuda")
        dg, db = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
        nbytes = x.numel() * 2
        cols = []
        for kind in ("fwd", "bwd"):
            for impl in (0, 1):
                if kind == "fwd":
                    fn = lambda: L.ub_groupnorm_nhwc_forward(p(x), p(gam), p(bet), p(y), p(cs), B, H, H, Cc, G, 1, impl)
                    traffic = 2 * nbytes
                else:
                    fn = lambda: L.ub_groupnorm_nhwc_backward(p(x), p(dy), p(cs), p(gam), p(bet), None, p(dx), p(dg),
                                                              p(db), p(scr), B, H, H, Cc, G, 1, impl)
                    traffic = 3 * nbytes
                assert fn() == 0, L.ub_last_error()
                ms = timeit(fn, a.reps, flush)
                cols.append(f"{ms * 1e3:7.1f} {traffic / ms * 1e-6:7.0f}")
        print(f"{Cc:4d} @{H:3d}x{H:<3d}  | " + " | ".join(f"{c:>16}" for c in cols))


if __name__ == "__main__":
    main()
