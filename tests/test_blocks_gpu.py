"""Composite block operators of the C ABI (ub_resblock_*, ub_attention_block_*: the launchers of dev/resblock.cuh and
dev/attention_block.cuh) against the oracle with autograd, at the reference's own unit-test scale reduced to a CPU
budget.  Every tensor is fp32 NCHW device memory owned by the caller, as in dev/resblock.cu:430-630; the final
output, the input / embedding gradients and every parameter gradient are compared (bf16 tensor-core bound for the
contractions, tests/gpu_util.py)."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import TOL_BF16, rel_inf

pytestmark = pytest.mark.gpu


def _struct(names):
    class S(C.Structure):
        _fields_ = [(n, C.c_void_p) for n in names]
    return S


RES_P = "gn1_w gn1_b cv3_1_w cv3_1_b l_emb_w l_emb_b gn2_w gn2_b cv3_2_w cv3_2_b res_cv1_w res_cv1_b".split()
RES_A = ("gn1 gn1_mean gn1_rstd silu1 ud_h ud_x cv3_1 silu_emb l_emb broad_emb add1 gn2 gn2_mean gn2_rstd silu2 cv3_2 "
         "res_cv1 add2 input emb").split()
RES_K = "buf_BCemb buf_BCHoWo buf1_BCHW buf2_BCHW dout dx demb".split()
ATT_P = "gn_w gn_b qkv_w qkv_b proj_w proj_b".split()
ATT_A = "gn gn_mean gn_rstd perm1 qkv1 qkv2 preatt att att_out proj perm2 add input".split()
ATT_K = "buf1_BCHW buf2_BCHW buf_B3CHW dqkvr dpreatt datt dout dinp".split()


def _fill(S, tensors):
    s = S()
    for n, _ in S._fields_:
        t = tensors.get(n)
        setattr(s, n, t.data_ptr() if t is not None else None)
    return s


def _dev(n):
    return torch.zeros(int(n), device="cuda")


@pytest.mark.parametrize("C_,Co,H,up,down", [(64, 64, 16, 0, 0), (64, 128, 16, 0, 0), (128, 64, 8, 1, 0),
                                             (64, 64, 16, 0, 1)])
def test_resblock_forward_backward(ub, oracle, C_, Co, H, up, down):
    B, Cemb, G, W = 2, 256, 32, H
    Ho = H * 2 if up else (H // 2 if down else H)
    g = torch.Generator().manual_seed(3)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).requires_grad_(True)
    x, emb = rn(B, C_, H, W), rn(B, Cemb)
    P = {"gn1_w": rn(C_, sc=0.3), "gn1_b": rn(C_, sc=0.3), "cv3_1_w": rn(Co, C_, 3, 3, sc=1 / math.sqrt(9 * C_)),
         "cv3_1_b": rn(Co, sc=0.1), "l_emb_w": rn(Co, Cemb, sc=1 / math.sqrt(Cemb)), "l_emb_b": rn(Co, sc=0.1),
         "gn2_w": rn(Co, sc=0.3), "gn2_b": rn(Co, sc=0.3), "cv3_2_w": rn(Co, Co, 3, 3, sc=1 / math.sqrt(9 * Co)),
         "cv3_2_b": rn(Co, sc=0.1)}
    if C_ != Co:
        P["res_cv1_w"], P["res_cv1_b"] = rn(Co, C_, sc=1 / math.sqrt(C_)), rn(Co, sc=0.1)
    # oracle: dev/resblock.py:107-160 (GN -> SiLU -> resample -> conv; emb; GN -> SiLU -> conv; skip)
    rs = (lambda t: oracle.upsample2(t)) if up else ((lambda t: oracle.avgpool2(t)) if down else (lambda t: t))
    h = oracle.conv3x3(rs(oracle.silu(oracle.groupnorm(x, P["gn1_w"], P["gn1_b"], G))), P["cv3_1_w"], P["cv3_1_b"])
    h = h + F.linear(oracle.silu(emb), P["l_emb_w"], P["l_emb_b"])[:, :, None, None]
    h = oracle.conv3x3(oracle.silu(oracle.groupnorm(h, P["gn2_w"], P["gn2_b"], G)), P["cv3_2_w"], P["cv3_2_b"])
    xs = rs(x)
    if C_ != Co:
        xs = F.conv2d(xs, P["res_cv1_w"][:, :, None, None], P["res_cv1_b"])
    out = xs + h
    dout = torch.randn(B, Co, Ho, Ho, generator=g)
    out.backward(dout)

    dp = {k: v.detach().cuda().contiguous() for k, v in P.items()}
    dg = {k: torch.zeros_like(v) for k, v in dp.items()}
    n_in, n_mid, n_out, ng = B * C_ * H * W, B * C_ * Ho * Ho, B * Co * Ho * Ho, B * G
    sizes = dict(gn1=n_in, gn1_mean=ng, gn1_rstd=ng, silu1=n_in, ud_h=n_mid, ud_x=n_mid, cv3_1=n_out, silu_emb=B * Cemb,
                 l_emb=B * Co, broad_emb=n_out, add1=n_out, gn2=n_out, gn2_mean=ng, gn2_rstd=ng, silu2=n_out, cv3_2=n_out,
                 res_cv1=n_out, add2=n_out)
    acts = {k: _dev(v) for k, v in sizes.items()}
    acts["input"], acts["emb"] = x.detach().cuda().contiguous(), emb.detach().cuda().contiguous()
    back = dict(buf_BCemb=_dev(B * Cemb), buf_BCHoWo=_dev(n_mid), buf1_BCHW=_dev(n_in), buf2_BCHW=_dev(n_in),
                dout=dout.cuda().contiguous(), dx=_dev(n_in), demb=_dev(B * Cemb))
    SP, SA, SK = _struct(RES_P), _struct(RES_A), _struct(RES_K)
    p, gr, a, k = _fill(SP, dp), _fill(SP, dg), _fill(SA, acts), _fill(SK, back)
    L = ub.lib()
    rc = L.ub_resblock_forward(C_, Cemb, Co, B, H, W, up, down, G, C.byref(p), C.byref(a))
    torch.cuda.synchronize()
    assert rc == 0, L.ub_last_error()
    assert rel_inf(acts["add2"].view(B, Co, Ho, Ho), out) <= TOL_BF16
    rc = L.ub_resblock_backward(C_, Cemb, Co, B, H, W, up, down, G, C.byref(p), C.byref(gr), C.byref(a), C.byref(k))
    torch.cuda.synchronize()
    assert rc == 0, L.ub_last_error()
    assert rel_inf(back["dx"].view_as(x), x.grad) <= TOL_BF16
    assert rel_inf(back["demb"].view_as(emb), emb.grad) <= TOL_BF16
    for name, ref in P.items():
        assert rel_inf(dg[name].view_as(ref), ref.grad) <= TOL_BF16, name


@pytest.mark.parametrize("C_,H", [(64, 8), (192, 16)])
def test_attention_block_forward_backward(ub, oracle, C_, H):
    B, HS, G, W = 2, 32, 32, H
    T = H * W
    g = torch.Generator().manual_seed(4)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).requires_grad_(True)
    x = rn(B, C_, H, W)
    P = {"gn_w": rn(C_, sc=0.5), "gn_b": rn(C_, sc=0.3), "qkv_w": rn(3 * C_, C_, sc=1 / math.sqrt(C_)),
         "qkv_b": rn(3 * C_, sc=0.1), "proj_w": rn(C_, C_, sc=1 / math.sqrt(C_)), "proj_b": rn(C_, sc=0.1)}
    OP = {"a.gn.weight": P["gn_w"], "a.gn.bias": P["gn_b"], "a.qkv.weight": P["qkv_w"][:, :, None],
          "a.qkv.bias": P["qkv_b"], "a.proj.weight": P["proj_w"][:, :, None], "a.proj.bias": P["proj_b"]}
    out = oracle.attention_block(x, OP, "a", HS, G)
    dout = torch.randn(B, C_, H, W, generator=g)
    out.backward(dout)

    dp = {k: v.detach().cuda().contiguous() for k, v in P.items()}
    dg = {k: torch.zeros_like(v) for k, v in dp.items()}
    nx, ng, ntt = B * C_ * T, B * G, B * (C_ // HS) * T * T
    sizes = dict(gn=nx, gn_mean=ng, gn_rstd=ng, perm1=nx, qkv1=3 * nx, qkv2=3 * nx, preatt=ntt, att=ntt, att_out=nx,
                 proj=nx, perm2=nx, add=nx)
    acts = {k: _dev(v) for k, v in sizes.items()}
    acts["input"] = x.detach().cuda().contiguous()
    back = dict(buf1_BCHW=_dev(nx), buf2_BCHW=_dev(nx), buf_B3CHW=_dev(3 * nx), dqkvr=_dev(3 * nx), dpreatt=_dev(ntt),
                datt=_dev(ntt), dout=dout.cuda().contiguous(), dinp=_dev(nx))
    SP, SA, SK = _struct(ATT_P), _struct(ATT_A), _struct(ATT_K)
    p, gr, a, k = _fill(SP, dp), _fill(SP, dg), _fill(SA, acts), _fill(SK, back)
    L = ub.lib()
    rc = L.ub_attention_block_forward(B, C_, H, W, HS, G, C.byref(p), C.byref(a))
    torch.cuda.synchronize()
    assert rc == 0, L.ub_last_error()
    assert rel_inf(acts["add"].view_as(x), out) <= TOL_BF16
    rc = L.ub_attention_block_backward(B, C_, H, W, HS, G, C.byref(p), C.byref(a), C.byref(k), C.byref(gr))
    torch.cuda.synchronize()
    assert rc == 0, L.ub_last_error()
    assert rel_inf(back["dinp"].view_as(x), x.grad) <= TOL_BF16
    for name, ref in P.items():
        assert rel_inf(dg[name].view_as(ref), ref.grad) <= TOL_BF16, name
