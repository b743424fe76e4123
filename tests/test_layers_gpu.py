"""Drop-in layer operators (include/unet_b200.h section 1) vs the oracle, at the shapes the reference's own unit
tests use (batch reduced where the CPU oracle would take too long) plus ragged / edge
shapes.  Every call goes through the C ABI with fp32 NCHW device buffers, like a caller of dev/*.cuh would."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import TOL_BF16, TOL_F32, call, dev, rel_inf

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.fixture(autouse=True, scope="module")
def _threads():
    torch.set_num_threads(os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------ conv 3x3
CONV3 = [  # (B, Cin, Cout, H, W, tensor-core path?)
    (8, 192, 64, 64, 64, True),    # dev/conv2d_k3.cu:2586-2590 shape (B reduced from 32 for the CPU oracle)
    (4, 64, 64, 64, 64, True),
    (4, 128, 128, 32, 32, True),
    (4, 192, 192, 16, 16, True),
    (6, 256, 256, 8, 8, True),
    (3, 448, 256, 8, 8, True),
    (2, 64, 128, 20, 12, True),    # ragged spatial size
    (4, 3, 64, 64, 64, False),     # first layer: exact fp32 path
    (4, 64, 3, 64, 64, False),     # last layer
    (1, 8, 16, 5, 7, True),        # smallest tensor-path shape
]


@pytest.mark.parametrize("B,Cin,Cout,H,W,tc", CONV3)
def test_conv2d_k3_forward_backward(ub, oracle, B, Cin, Cout, H, W, tc):
    x = rnd(B, Cin, H, W, seed=1).requires_grad_(True)
    w = (rnd(Cout, Cin, 3, 3, seed=2) / math.sqrt(9 * Cin)).requires_grad_(True)
    b = rnd(Cout, seed=3).requires_grad_(True)
    dout = rnd(B, Cout, H, W, seed=4)
    y = oracle.conv3x3(x, w, b)
    y.backward(dout)
    tol = TOL_BF16 if tc else TOL_F32
    dx_, dw_, db_ = dev(x.detach()), dev(w.detach()), dev(b.detach())
    out = torch.empty(B, Cout, H, W, device="cuda")
    call(ub, "conv2d_k3_forward3", dx_, dw_, db_, out, B, Cin, Cout, H, W)
    assert rel_inf(out, y) <= tol
    gdx, gdw, gdb = torch.empty_like(dx_), torch.empty_like(dw_), torch.empty_like(db_)
    call(ub, "conv2d_k3_backward2", dev(dout), dx_, dw_, None, None, gdx, gdw, gdb, B, Cin, Cout, H, W)
    assert rel_inf(gdx, x.grad) <= tol
    assert rel_inf(gdw, w.grad) <= tol
    assert rel_inf(gdb, b.grad) <= TOL_F32


def test_conv2d_k3_is_linear_at_full_size(ub):
    """Config-2 size (B=32, 192->64 @64x64): conv(x1 + x2) == conv(x1) + conv(x2) - bias, no oracle needed."""
    B, Cin, Cout, H, W = 32, 192, 64, 64, 64
    x1, x2 = dev(rnd(B, Cin, H, W, seed=1)), dev(rnd(B, Cin, H, W, seed=2))
    w, b = dev(rnd(Cout, Cin, 3, 3, seed=3) / math.sqrt(9 * Cin)), dev(rnd(Cout, seed=4))
    o1, o2, o12 = (torch.empty(B, Cout, H, W, device="cuda") for _ in range(3))
    call(ub, "conv2d_k3_forward3", x1, w, b, o1, B, Cin, Cout, H, W)
    call(ub, "conv2d_k3_forward3", x2, w, b, o2, B, Cin, Cout, H, W)
    call(ub, "conv2d_k3_forward3", (x1 + x2).contiguous(), w, b, o12, B, Cin, Cout, H, W)
    ref = o1 + o2 - b.view(1, -1, 1, 1)
    assert rel_inf(o12, ref) <= TOL_BF16
    # translation structure: a one-pixel impulse reproduces the (flipped) filter taps
    imp = torch.zeros(1, Cin, 8, 8)
    imp[0, 5, 4, 4] = 1.0
    o = torch.empty(1, Cout, 8, 8, device="cuda")
    call(ub, "conv2d_k3_forward3", dev(imp), w, torch.zeros(Cout, device="cuda"), o, 1, Cin, Cout, 8, 8)
    taps = w[:, 5].flip(-1).flip(-2)
    assert rel_inf(o[0, :, 3:6, 3:6], taps) <= 1e-2


# ------------------------------------------------------------------------------------------ conv 1x1 / linear
@pytest.mark.parametrize("B,Cin,Cout,H,W", [(8, 64, 128, 64, 64), (4, 448, 192, 16, 16), (2, 512, 256, 8, 8),
                                            (2, 24, 40, 6, 6)])
def test_conv2d_k1_forward_backward(ub, oracle, B, Cin, Cout, H, W):
    x = rnd(B, Cin, H, W, seed=1).requires_grad_(True)
    w = (rnd(Cout, Cin, seed=2) / math.sqrt(Cin)).requires_grad_(True)
    b = rnd(Cout, seed=3).requires_grad_(True)
    dout = rnd(B, Cout, H, W, seed=4)
    y = oracle.conv1x1(x, w, b)
    y.backward(dout)
    tc = Cin % 8 == 0 and Cout % 16 == 0
    tol = TOL_BF16 if tc else TOL_F32
    dx_, dw_, db_ = dev(x.detach()), dev(w.detach()), dev(b.detach())
    out = torch.empty(B, Cout, H, W, device="cuda")
    call(ub, "conv2d_k1_forward2", dx_, dw_, db_, out, B, Cin, H, W, Cout)
    assert rel_inf(out, y) <= tol
    gdx, gdw, gdb = torch.empty_like(dx_), torch.empty_like(dw_), torch.empty_like(db_)
    call(ub, "conv2d_k1_backward1", dev(dout), dx_, dw_, gdx, gdw, gdb, B, Cin, Cout, H, W)
    assert rel_inf(gdx, x.grad) <= TOL_BF16
    assert rel_inf(gdw, w.grad) <= TOL_BF16
    assert rel_inf(gdb, b.grad) <= TOL_F32


@pytest.mark.parametrize("N,C,OC", [(32, 64, 128), (32, 256, 192), (2048, 256, 768), (8192, 192, 192), (5, 7, 3)])
def test_matmul_forward_backward(ub, N, C, OC):
    """dev/linear.cu:162-164 shape (N32 64->128) and the attention qkv / proj shapes."""
    x = rnd(N, C, seed=1).requires_grad_(True)
    w = (rnd(OC, C, seed=2) / math.sqrt(C)).requires_grad_(True)
    b = rnd(OC, seed=3).requires_grad_(True)
    dout = rnd(N, OC, seed=4)
    y = F.linear(x, w, b)
    y.backward(dout)
    tc = N >= 128
    tol = TOL_BF16 if tc else TOL_F32
    out = torch.empty(N, OC, device="cuda")
    call(ub, "matmul_forward2", out, dev(x.detach()), dev(w.detach()), dev(b.detach()), N, C, OC)
    assert rel_inf(out, y) <= tol
    dinp, dw, db = torch.empty(N, C, device="cuda"), torch.empty(OC, C, device="cuda"), torch.empty(OC, device="cuda")
    call(ub, "matmul_backward1", dinp, dw, db, dev(dout), dev(x.detach()), dev(w.detach()), N, C, OC)
    assert rel_inf(dinp, x.grad) <= tol
    assert rel_inf(dw, w.grad) <= tol
    assert rel_inf(db, b.grad) <= TOL_F32


# ------------------------------------------------------------------------------------------ groupnorm / silu
@pytest.mark.parametrize("B,C,H,W,G", [(16, 128, 4, 8, 32), (4, 192, 16, 16, 32), (2, 64, 64, 64, 32), (3, 96, 5, 7, 8)])
def test_groupnorm_forward_backward(ub, oracle, B, C, H, W, G):
    """dev/groupnorm.cu:275-278 shape (B16 C128 4x8, 32 groups) + U-Net shapes."""
    x = (rnd(B, C, H, W, seed=1) * 2 + 0.5).requires_grad_(True)
    w = (rnd(C, seed=2) * 0.5 + 1).requires_grad_(True)
    b = rnd(C, seed=3).requires_grad_(True)
    dout = rnd(B, C, H, W, seed=4)
    y = oracle.groupnorm(x, w, b, G)
    y.backward(dout)
    xd, wd, bd = dev(x.detach()), dev(w.detach()), dev(b.detach())
    out, mean, rstd = torch.empty_like(xd), torch.empty(B * G, device="cuda"), torch.empty(B * G, device="cuda")
    call(ub, "groupnorm_forward", xd, wd, bd, out, mean, rstd, B, C, H, W, G)
    assert rel_inf(out, y) <= TOL_F32
    xg = x.detach().reshape(B, G, -1)
    assert rel_inf(mean, xg.mean(2).reshape(-1)) <= TOL_F32
    assert rel_inf(rstd, torch.rsqrt(xg.var(2, unbiased=False) + 1e-5).reshape(-1)) <= TOL_F32
    dx = torch.empty_like(xd)
    dw = torch.full((C,), 0.25, device="cuda")   # backward ACCUMULATES into dweight / dbias (reference atomics)
    db = torch.full((C,), -0.5, device="cuda")
    call(ub, "groupnorm_backward", dev(dout), xd, mean, rstd, wd, dx, dw, db, B, C, H, W, G)
    assert rel_inf(dx, x.grad) <= TOL_F32
    assert rel_inf(dw - 0.25, w.grad) <= TOL_F32
    assert rel_inf(db + 0.5, b.grad) <= TOL_F32


def test_silu_add_elementwise(ub, oracle):
    n = 100003
    x = rnd(n, seed=1).requires_grad_(True)
    dout = rnd(n, seed=2)
    y = oracle.silu(x)
    y.backward(dout)
    xd, out, dx = dev(x.detach()), torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    call(ub, "silu_forward", xd, out, n)
    call(ub, "silu_backward", dev(dout), xd, dx, n)
    assert rel_inf(out, y) <= TOL_F32 and rel_inf(dx, x.grad) <= TOL_F32
    a, b = dev(rnd(n, seed=3)), dev(rnd(n, seed=4))
    s = torch.empty(n, device="cuda")
    call(ub, "add_forward", a, b, s, n)
    assert torch.equal(s, a + b)
    b2 = b.clone()
    call(ub, "add_inplace_forward", a, b2, n)
    assert torch.equal(b2, a + b)
    call(ub, "silu_forward", xd, out, 0)   # empty input is a no-op


# ------------------------------------------------------------------------------------------ resampling / concat
def test_upsample_avgpool_concat_broadcast(ub, oracle):
    B, C, H, W = 3, 10, 6, 8
    x = rnd(B, C, H, W, seed=1).requires_grad_(True)
    up = oracle.upsample2(x)
    dup = rnd(B, C, 2 * H, 2 * W, seed=2)
    up.backward(dup)
    xd = dev(x.detach())
    o = torch.empty(B, C, 2 * H, 2 * W, device="cuda")
    call(ub, "upsample_forward1", o, xd, B, C, H, W)
    assert torch.equal(o.cpu(), up.detach())
    dx = torch.empty_like(xd)
    call(ub, "upsample_backward1", dx, dev(dup), B, C, H, W)
    assert rel_inf(dx, x.grad) <= 1e-6
    x2 = rnd(B, C, H, W, seed=3).requires_grad_(True)
    pl = oracle.avgpool2(x2)
    dpl = rnd(B, C, H // 2, W // 2, seed=4)
    pl.backward(dpl)
    o = torch.empty(B, C, H // 2, W // 2, device="cuda")
    call(ub, "avgpool_2d_forward1", o, dev(x2.detach()), B, C, H, W)
    assert rel_inf(o, pl) <= 1e-6
    dx = torch.empty(B, C, H, W, device="cuda")
    call(ub, "avgpool_2d_backward1", dev(dpl), dx, B, C, H, W)
    assert rel_inf(dx, x2.grad) <= 1e-6
    a, b = rnd(B, 5, H, W, seed=5), rnd(B, 7, H, W, seed=6)
    cat = torch.empty(B, 12, H, W, device="cuda")
    call(ub, "concat_channel_forward", dev(a), dev(b), cat, B, 5, 7, H, W)
    assert torch.equal(cat.cpu(), torch.cat([a, b], 1))
    da, db = torch.empty(B, 5, H, W, device="cuda"), torch.empty(B, 7, H, W, device="cuda")
    call(ub, "concat_channel_backward", cat, da, db, B, 5, 7, H, W)
    assert torch.equal(da.cpu(), a) and torch.equal(db.cpu(), b)
    v = rnd(B * C, seed=7)
    bo = torch.empty(B * C, H, W, device="cuda")
    call(ub, "broadcast_last_dims_forward", dev(v), bo, B * C, H, W)
    assert torch.equal(bo.cpu(), v.view(-1, 1, 1).expand(-1, H, W))
    dv = torch.empty(B * C, device="cuda")
    call(ub, "broadcast_last_dims_backward", dev(dup[:, :, :H, :W].reshape(B * C, H, W).contiguous()), dv, B * C, H, W)
    assert rel_inf(dv, dup[:, :, :H, :W].reshape(B * C, -1).sum(1)) <= 1e-5


def test_mse_and_timestep_embedding(ub, oracle):
    n = 4 * 3 * 64 * 64
    a, y = rnd(n, seed=1), rnd(n, seed=2)
    loss = torch.zeros(1, device="cuda")
    call(ub, "mse_forward", dev(a), dev(y), loss, n)
    assert abs(float(loss) - float(oracle.mse_loss(a, y))) <= 1e-5
    d = torch.empty(n, device="cuda")
    call(ub, "mse_backward", dev(a), dev(y), d, n)
    assert rel_inf(d, 2 * (a - y) / n) <= 1e-6
    t = torch.tensor([0.0, 1.0, 17.0, 500.0, 999.0])
    o = torch.empty(5, 64, device="cuda")
    call(ub, "get_timestep_embeddings", dev(t), o, 5, 64, 1000)
    ref = oracle.timestep_embedding(t, 64, 1000)
    assert float((o.cpu() - ref).abs().max()) <= 1e-3   # dev/timestep_embedding.cu:64-95 tolerance


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,T,C,NH", [(2, 256, 256, 8), (4, 64, 256, 8), (2, 256, 64, 8), (1, 1024, 64, 2)])
def test_attention_forward_backward(ub, B, T, C, NH):
    """dev/attention.cu:371-374 (B4 T1024 C256 HS32, reduced) in the reference's (B,T,3,NH,HS) layout."""
    HS = C // NH
    inp = rnd(B, T, 3 * C, seed=1).requires_grad_(True)
    dout = rnd(B, T, C, seed=2)
    qkv = inp.view(B, T, 3, NH, HS)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))          # (B, NH, T, HS)
    pre = q @ k.transpose(-1, -2) / math.sqrt(HS)
    att = torch.softmax(pre, dim=-1)
    y = (att @ v).permute(0, 2, 1, 3).reshape(B, T, C)
    y.backward(dout)
    out = torch.empty(B, T, C, device="cuda")
    qkvr = torch.empty(3, B, NH, T, HS, device="cuda")
    preatt = torch.empty(B, NH, T, T, device="cuda")
    attd = torch.empty(B, NH, T, T, device="cuda")
    call(ub, "attention_forward1", out, qkvr, preatt, attd, dev(inp.detach()), B, T, C, NH)
    assert rel_inf(out, y) <= TOL_F32
    assert rel_inf(attd, att) <= TOL_F32 and rel_inf(preatt, pre) <= TOL_F32
    assert rel_inf(qkvr[0], q) <= 1e-6
    dinp = torch.empty(B, T, 3 * C, device="cuda")
    scratch = [torch.empty_like(qkvr), torch.empty_like(preatt), torch.empty_like(attd)]
    call(ub, "attention_backward", dinp, scratch[0], scratch[1], scratch[2], None, dev(dout), qkvr, attd, B, T, C, NH)
    assert rel_inf(dinp, inp.grad) <= TOL_F32


def test_shape_errors_are_reported_not_fatal(ub):
    x = torch.zeros(1, 10, 4, 4, device="cuda")
    rc = ub.lib().ub_groupnorm_forward(None, None, None, None, None, None, 1, 10, 4, 4, 3)
    assert rc != 0 and b"n_groups" in ub.lib().ub_last_error()
    rc = ub.lib().ub_avgpool_2d_forward1(None, None, 1, 1, 3, 4)
    assert rc != 0
