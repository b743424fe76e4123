"""Helpers for the GPU tests: torch CUDA tensors provide device memory; every call goes through the C ABI."""
import ctypes as C

import numpy as np
import torch


def dev(a):
    """numpy / torch CPU tensor -> contiguous fp32 CUDA tensor."""
    t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).float().contiguous()
    return t.cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def call(ub, name, *args):
    """Call lib.ub_<name>(...) with torch CUDA tensors / ints, check the status, synchronise."""
    fn = getattr(ub.lib(), "ub_" + name)
    cargs = []
    for a in args:
        if a is None:
            cargs.append(None)
        elif isinstance(a, torch.Tensor):
            assert a.is_cuda and a.is_contiguous() and a.dtype == torch.float32
            cargs.append(C.c_void_p(a.data_ptr()))
        else:
            cargs.append(C.c_int(int(a)))
    rc = fn(*cargs)
    torch.cuda.synchronize()
    assert rc == 0, (name, rc, ub.lib().ub_last_error())


def rel_inf(a, b):
    a = a.detach().float().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().float().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# Stated tolerances (north_star: fp32 paths 1e-3 relative; bf16 tensor paths a looser, stated bound).
# bf16 operands carry 2^-9 relative rounding each; with fp32 accumulation over K terms the error of a dot product is
# ~ sqrt(K) * 2^-9 * |x||w| (random signs), i.e. a few 1e-3 of the output scale for K up to 4608; we allow 2e-2 of the
# tensor's max magnitude, the same order as the reference's own 1e-2 absolute tolerance on O(1) tensors
# (dev/conv2d_k3.cu:2662,2712-2726).
TOL_F32 = 1e-3
TOL_BF16 = 2e-2
