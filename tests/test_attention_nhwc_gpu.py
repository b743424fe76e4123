"""Attention core in the native NHWC bf16 layout (what the trainer runs): the tcgen05 kernels (impl 0) and the SIMT
fallback (impl 1) against the oracle's QKVAttention (oracle.mha == dev/unet.py:75-87, dev/attention.py:6-24) on the
same bf16-rounded inputs.  Shapes: the two attention resolutions of the default U-Net (T=256 C=192 NH=6, T=64 C=256
NH=8), a ragged last tile (odd number of 8x8 images), T=128 and T=16."""
import ctypes as C

import pytest
import torch

from gpu_util import TOL_BF16, rel_inf

pytestmark = pytest.mark.gpu

SHAPES = [(4, 256, 192, 6), (4, 64, 256, 8), (3, 64, 128, 4), (2, 128, 64, 2), (5, 16, 64, 2), (32, 256, 192, 6)]


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("B,T,Cc,NH", SHAPES)
def test_attention_nhwc_forward_backward(ub, oracle, B, T, Cc, NH, impl):
    if impl == 1 and B == 32:
        pytest.skip("full-size case is for the tensor-core path")
    g = torch.Generator().manual_seed(7)
    qkv = (torch.randn(B, T, 3 * Cc, generator=g) * 1.5).bfloat16()
    dout = torch.randn(B, T, Cc, generator=g).bfloat16()
    # oracle: (B, 3C, T) layout, fp32, on the same bf16-rounded values
    x = qkv.float().permute(0, 2, 1).contiguous().requires_grad_(True)
    y = oracle.mha(x, NH)                      # (B, C, T)
    y.backward(dout.float().permute(0, 2, 1).contiguous())
    y_ref = y.detach().permute(0, 2, 1)        # (B, T, C)
    dqkv_ref = x.grad.permute(0, 2, 1)         # (B, T, 3C)

    L = ub.lib()
    d_qkv, d_dout = qkv.cuda(), dout.cuda()
    d_out = torch.zeros(B, T, Cc, dtype=torch.bfloat16, device="cuda")
    d_lse = torch.zeros(B, NH, T, dtype=torch.float32, device="cuda")
    rc = L.ub_attention_nhwc_forward(_p(d_qkv), _p(d_out), _p(d_lse), B, T, Cc, NH, impl)
    torch.cuda.synchronize()
    assert rc == 0, L.ub_last_error()
    assert rel_inf(d_out, y_ref) <= TOL_BF16
    d_dqkv = torch.zeros(B, T, 3 * Cc, dtype=torch.bfloat16, device="cuda")
    d_dsum = torch.zeros(B, NH, T, dtype=torch.float32, device="cuda")
    rc = L.ub_attention_nhwc_backward(_p(d_qkv), _p(d_out), _p(d_dout), _p(d_lse), _p(d_dqkv), _p(d_dsum), B, T, Cc, NH,
                                      impl)
    torch.cuda.synchronize()
    assert rc == 0, L.ub_last_error()
    for name, sl in (("dq", slice(0, Cc)), ("dk", slice(Cc, 2 * Cc)), ("dv", slice(2 * Cc, 3 * Cc))):
        assert rel_inf(d_dqkv[..., sl], dqkv_ref[..., sl]) <= TOL_BF16, name


def test_attention_tc_rejects_unsupported(ub):
    L = ub.lib()
    t = torch.zeros(8, device="cuda")
    assert L.ub_attention_nhwc_forward(_p(t), _p(t), _p(t), 1, 48, 64, 2, 0) != 0   # T not a supported length
    assert L.ub_attention_nhwc_forward(_p(t), _p(t), _p(t), 1, 64, 96, 3, 0) != 0   # odd number of heads
    assert L.ub_attention_nhwc_forward(_p(t), _p(t), _p(t), 1, 64, 64, 4, 1) != 0   # head size != 32
