"""Direct check of the oracle against the reference's own Python modules -- only where /root/reference exists
(the build container).  On the GPU box the committed golden vectors (test_oracle_golden.py) carry the pin."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dev")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_unet():
    sys.path.insert(0, os.path.join(REF, "dev"))
    import unet as ref_unet_mod  # dev/unet.py (imports dev/resblock.py, dev/utils.py)
    return ref_unet_mod


@pytest.mark.parametrize("B", [1, 2])
def test_forward_backward_equal(oracle, ref_unet, B):
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig()
    torch.manual_seed(0)
    model = ref_unet.UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32)
    P = O.init_params(cfg, seed=0)
    for n, p in model.named_parameters():
        assert torch.equal(P[n], p.detach()), n
    x0, t, noise = O.synthetic_batch(cfg, B)
    xt = O.q_sample(x0, t, noise)
    flat = O.flatten_params(cfg, P)
    loss, out, g = O.train_step_grads(cfg, flat, x0, t, noise)
    out_ref = model(xt, t)
    loss_ref = ((out_ref - noise) ** 2).mean()
    loss_ref.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert abs(float(loss) - float(loss_ref)) < 1e-6
    assert float((out - out_ref.detach()).abs().max()) < 1e-5
    assert float((g - g_ref).abs().max()) < 1e-6 + 1e-4 * float(g_ref.abs().max())


def test_five_level_config_matches_reference(oracle, ref_unet):
    """BASELINE config 5 shape (128x128, channel_mult 1-1-2-3-4, attention at 16x16 and 8x8), tiny batch."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(channel_mult=(1, 1, 2, 3, 4), attn_start_level=3, H=128, W=128)
    torch.manual_seed(0)
    model = ref_unet.UNetModel(3, 64, 3, 2, cfg.attention_resolutions(), channel_mult=cfg.channel_mult,
                               num_head_channels=32)
    names = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    assert names == O.param_spec(cfg)
    assert O.num_params(cfg) == 21082755
    P = O.init_params(cfg, seed=0)
    x = torch.randn(1, 3, 128, 128)
    t = torch.tensor([[17.0]])
    with torch.no_grad():
        assert float((O.unet_forward(cfg, P, x, t) - model(x, t)).abs().max()) < 1e-5


def test_sampler_matches_generate_py(oracle, ref_unet):
    """oracle.ddpm_sample vs the reference's own generate.sample_next_step (generate.py:29-52), three steps."""
    sys.path.insert(0, REF)
    gen = pytest.importorskip("generate")     # imports PIL, tqdm and train_unet (all __main__-guarded)
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig()
    torch.manual_seed(0)
    model = ref_unet.UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32).eval()
    P = O.init_params(cfg, seed=0)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 3, 64, 64, generator=g)
    noises = [torch.randn(1, 3, 64, 64, generator=g) for _ in range(3)]
    import train_unet as tu
    betas = tu.get_named_beta_schedule("linear", 1000)
    beta_t = torch.tensor(betas, dtype=torch.float32)
    acp = torch.tensor(tu.GaussianDiffusion(betas=betas).alphas_cumprod)
    x_ref = x.clone()
    with torch.no_grad():
        for i, t in enumerate([700, 699, 698]):
            torch.manual_seed(1000 + i)
            z = torch.randn_like(x_ref)        # what sample_next_step will draw
            noises[i] = z
            torch.manual_seed(1000 + i)
            x_ref = gen.sample_next_step(x_ref, torch.tensor([[t]]), model, 1000, beta_t, acp).float()
    x_or = O.ddpm_sample(cfg, P, x, 700, 698, noises)
    assert float((x_or - x_ref).abs().max()) < 1e-4


def test_class_conditional_matches_reference(oracle, ref_unet):
    """num_classes (dev/unet.py:142,174-175,301-303): parameter order / init bit-identical, forward and every gradient
    (the label-embedding rows included) against the reference's own UNetModel.  The zero-initialised conv2 / proj
    weights are perturbed first -- with them at zero nothing upstream of a ResBlock's second conv receives a gradient,
    the embedding path included."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(num_classes=10)
    torch.manual_seed(0)
    model = ref_unet.UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32, num_classes=10)
    names = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    assert names == O.param_spec(cfg)
    assert names[4] == ("label_emb.weight", (10, 256))
    P = O.init_params(cfg, seed=0)
    for n, p in model.named_parameters():
        assert torch.equal(P[n], p.detach()), n
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if not p.detach().any():
                p.add_(0.02 * torch.randn(p.shape, generator=g))
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    B = 2
    x0, t, noise = O.synthetic_batch(cfg, B)
    y = torch.tensor([3, 7])
    loss, out, gr = O.train_step_grads(cfg, flat, x0, t, noise, y)
    out_ref = model(O.q_sample(x0, t, noise), t, y)
    loss_ref = ((out_ref - noise) ** 2).mean()
    loss_ref.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert abs(float(loss) - float(loss_ref)) < 1e-6
    assert float((out - out_ref.detach()).abs().max()) < 1e-5
    assert float((gr - g_ref).abs().max()) < 1e-6 + 1e-4 * float(g_ref.abs().max())
    ge = model.label_emb.weight.grad
    assert float(ge[3].abs().max()) > 0 and float(ge[7].abs().max()) > 0 and float(ge[0].abs().max()) == 0


def test_resblock_updown_matches_reference(oracle, ref_unet):
    """resblock_updown=True (dev/unet.py:147,205-222,271-284): parameter order / init bit-identical, forward output and
    every gradient against the reference's own UNetModel (zero-initialised tensors perturbed, see
    perturb_zero_params)."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(resblock_updown=True)
    torch.manual_seed(0)
    model = ref_unet.UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32, resblock_updown=True)
    names = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    assert names == O.param_spec(cfg)
    P = O.init_params(cfg, seed=0)
    for n, p in model.named_parameters():
        assert torch.equal(P[n], p.detach()), n
    flat = O.perturb_zero_params(cfg, O.flatten_params(cfg, P))
    with torch.no_grad():
        off = 0
        for p in model.parameters():
            p.copy_(flat[off:off + p.numel()].reshape(p.shape))
            off += p.numel()
    B = 2
    x0, t, noise = O.synthetic_batch(cfg, B)
    loss, out, gr = O.train_step_grads(cfg, flat, x0, t, noise)
    out_ref = model(O.q_sample(x0, t, noise), t)
    loss_ref = ((out_ref - noise) ** 2).mean()
    loss_ref.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert abs(float(loss) - float(loss_ref.detach())) < 1e-6
    assert float((out - out_ref.detach()).abs().max()) < 1e-5
    assert float((gr - g_ref).abs().max()) < 1e-6 + 1e-4 * float(g_ref.abs().max())
