"""Pins oracle/unet_oracle.py against the golden vectors that oracle/gen_golden.py produced FROM THE REFERENCE
(dev/unet.py UNetModel, dev/resblock.py ResBlock, train_unet.py GaussianDiffusion + .bin writer)."""
import os

import numpy as np
import pytest
import torch


def _flat(O, cfg, P):
    return O.flatten_params(cfg, P)


def test_param_spec_and_init_match_reference(oracle, golden_dir):
    O = oracle
    g = np.load(os.path.join(golden_dir, "unet_step_B2.npz"))
    cfg = O.UNetConfig()
    spec = O.param_spec(cfg)
    assert len(spec) == 326 and O.num_params(cfg) == 20494211
    assert [n for n, _ in spec] == list(g["names"])
    assert [str(tuple(s)) for _, s in spec] == list(g["shapes"])
    P = O.init_params(cfg, seed=0)
    sums = np.array([float(P[n].double().sum()) for n, _ in spec])
    np.testing.assert_allclose(sums, g["param_sums"], rtol=0, atol=1e-9)
    head = np.stack([np.pad(P[n].reshape(-1)[:4].numpy(), (0, max(0, 4 - P[n].numel()))) for n, _ in spec])
    np.testing.assert_array_equal(head, g["param_head"])


def test_diffusion_tables_and_qsample(oracle, golden_dir):
    O = oracle
    g = np.load(os.path.join(golden_dir, "unet_step_B2.npz"))
    sa, sb = O.diffusion_tables()
    np.testing.assert_allclose(sa, g["sqrt_ac"], rtol=1e-6)
    np.testing.assert_allclose(sb, g["sqrt_1mac"], rtol=1e-6)
    x0, t, noise = O.synthetic_batch(O.UNetConfig(), 2)
    xt = O.q_sample(x0, t, noise)
    np.testing.assert_allclose(xt.reshape(-1)[::997].numpy(), g["x_t_slice"], rtol=1e-6, atol=1e-6)


def test_train_step_matches_reference(oracle, golden_dir):
    """Forward output, all 326 gradients and a 3-step AdamW loss trace (reference: torch.optim.AdamW; oracle:
    the adamw_kernel2 restatement) at B=2."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = np.load(os.path.join(golden_dir, "unet_step_B2.npz"))
    cfg = O.UNetConfig()
    flat = _flat(O, cfg, O.init_params(cfg, seed=0))
    x0, t, noise = O.synthetic_batch(cfg, 2)
    loss, out, grads = O.train_step_grads(cfg, flat, x0, t, noise)
    np.testing.assert_allclose(out.reshape(-1)[::37].numpy(), g["out_slice"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(grads[::4099].numpy(), g["grad_slice"], rtol=1e-3, atol=1e-6)
    off, norms = 0, []
    for _, s in O.param_spec(cfg):
        n = int(np.prod(s))
        norms.append(float(grads[off:off + n].double().norm()))
        off += n
    np.testing.assert_allclose(np.array(norms), g["grad_norms"], rtol=1e-3)
    batches = [O.synthetic_batch(cfg, 2, seed=1234 + i) for i in range(3)]
    losses, flat3 = O.train_steps(cfg, flat, batches, lr=1e-4, wd=0.0)
    np.testing.assert_allclose(np.array(losses), g["loss_trace"], rtol=2e-4)
    np.testing.assert_allclose(flat3[::4099].numpy(), g["params_after3_slice"], rtol=0, atol=2e-5)


def test_layers_match_reference_modules(oracle, golden_dir):
    O = oracle
    g = np.load(os.path.join(golden_dir, "layers.npz"))
    P = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("rb.")}
    P = {"b." + k: v for k, v in P.items()}
    y = O.resblock(torch.from_numpy(g["rb_x"]), torch.from_numpy(g["rb_emb"]), P, "b")
    np.testing.assert_allclose(y.numpy(), g["rb_y"], rtol=1e-4, atol=1e-5)
    A = {"a." + k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("ab.")}
    ya = O.attention_block(torch.from_numpy(g["ab_x"]), A, "a", head_size=32)
    np.testing.assert_allclose(ya.numpy(), g["ab_y"], rtol=1e-4, atol=1e-5)
    te = np.load(os.path.join(golden_dir, "timestep_embedding.npz"))
    np.testing.assert_allclose(O.timestep_embedding(torch.from_numpy(te["t"]), 64).numpy(), te["emb"], rtol=1e-5,
                               atol=1e-6)


def test_model_bin_layout_matches_reference_writer(oracle, golden_dir, tmp_path):
    """train_unet.py:768-795 wrote the golden file; the oracle's writer must produce the same bytes."""
    O = oracle
    g = np.load(os.path.join(golden_dir, "model_bin.npz"))
    cfg = O.UNetConfig()
    flat = _flat(O, cfg, O.init_params(cfg, seed=0)).numpy()
    path = str(tmp_path / "unet_init.bin")
    O.write_model_bin(path, cfg, flat, B=32)
    assert os.path.getsize(path) == int(g["file_bytes"][0]) == 81977868
    header, payload = O.read_model_bin(path)
    np.testing.assert_array_equal(header, g["header"])
    assert payload.size == int(g["n_floats"][0])
    np.testing.assert_array_equal(payload[::4099], g["payload_slice"])


def test_class_conditional_step_matches_reference(oracle, golden_dir):
    """num_classes = 10 (dev/unet.py:174-175, 301-303) from perturbed weights: the fixture came from the reference's own
    UNetModel(num_classes=10); loss, output, gradients incl. the complete label-embedding gradient."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = np.load(os.path.join(golden_dir, "class_cond_B2.npz"))
    cfg = O.UNetConfig(num_classes=10)
    assert O.param_spec(cfg)[4] == ("label_emb.weight", (10, 256))
    assert O.num_params(cfg) == 20494211 + 2560
    flat = O.perturb_zero_params(cfg, _flat(O, cfg, O.init_params(cfg, seed=0)))
    np.testing.assert_array_equal(flat[::4099].numpy(), g["param_slice"])
    x0, t, noise = O.synthetic_batch(cfg, 2)
    loss, out, grads = O.train_step_grads(cfg, flat, x0, t, noise, torch.from_numpy(g["labels"]))
    assert abs(float(loss) - float(g["loss"][0])) < 1e-6
    np.testing.assert_allclose(out.reshape(-1)[::37].numpy(), g["out_slice"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(grads[::4099].numpy(), g["grad_slice"], rtol=1e-3, atol=1e-6)
    P = O.unflatten_params(cfg, grads)
    np.testing.assert_allclose(P["label_emb.weight"].numpy(), g["label_emb_grad"], rtol=1e-3, atol=1e-7)
    off, norms = 0, []
    for _, s in O.param_spec(cfg):
        n = int(np.prod(s))
        norms.append(float(grads[off:off + n].double().norm()))
        off += n
    np.testing.assert_allclose(np.array(norms), g["grad_norms"], rtol=1e-3)


def test_resblock_updown_step_matches_reference(oracle, golden_dir):
    """resblock_updown=True (dev/unet.py:147,205-222,271-284): the fixture came from the reference's own
    UNetModel(resblock_updown=True) with the zero-initialised tensors perturbed."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = np.load(os.path.join(golden_dir, "updown_B2.npz"))
    cfg = O.UNetConfig(resblock_updown=True)
    assert [n for n, _ in O.param_spec(cfg)] == [str(n) for n in g["names"]]
    flat = O.perturb_zero_params(cfg, _flat(O, cfg, O.init_params(cfg, seed=0)))
    np.testing.assert_array_equal(flat[::4099].numpy(), g["param_slice"])
    x0, t, noise = O.synthetic_batch(cfg, 2)
    loss, out, grads = O.train_step_grads(cfg, flat, x0, t, noise)
    assert abs(float(loss) - float(g["loss"][0])) < 1e-6
    np.testing.assert_allclose(out.reshape(-1)[::37].numpy(), g["out_slice"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(grads[::4099].numpy(), g["grad_slice"], rtol=1e-3, atol=1e-6)
    off, norms = 0, []
    for _, s in O.param_spec(cfg):
        n = int(np.prod(s))
        norms.append(float(grads[off:off + n].double().norm()))
        off += n
    np.testing.assert_allclose(np.array(norms), g["grad_norms"], rtol=1e-3)


def test_scale_shift_resblock_matches_reference_original(oracle, golden_dir):
    """use_scale_shift_norm (dev/resblock.py:211,243-247): the fixture came from the reference's copy of the original
    ResBlock (ResBlockO) -- output and the gradients of x, emb and every parameter, plain and with down=True."""
    O = oracle
    g = np.load(os.path.join(golden_dir, "resblock_scale_shift.npz"))
    for tag, updown in (("a", None), ("d", "down")):
        P = {k: torch.from_numpy(g[k]).requires_grad_(True) for k in g.files
             if k.startswith(tag + ".") and not k.startswith(tag + ".grad.")}
        assert P[tag + ".l_emb.weight"].shape[0] == 2 * P[tag + ".cv3_1.weight"].shape[0]
        x = torch.from_numpy(g[tag + "_x"]).requires_grad_(True)
        emb = torch.from_numpy(g[tag + "_emb"]).requires_grad_(True)
        y = O.resblock(x, emb, P, tag, updown=updown)
        np.testing.assert_allclose(y.detach().numpy(), g[tag + "_y"], rtol=1e-4, atol=1e-5)
        y.backward(torch.from_numpy(g[tag + "_dy"]))
        np.testing.assert_allclose(x.grad.numpy(), g[tag + "_dx"], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(emb.grad.numpy(), g[tag + "_demb"], rtol=1e-3, atol=1e-5)
        for k, v in P.items():
            ref = g[tag + ".grad." + k[len(tag) + 1:]]
            # (a bias in front of a GroupNorm has a gradient that is a difference of nearly equal sums: fp32 noise of
            #  ~5e-6 absolute around values of 1e-6)
            np.testing.assert_allclose(v.grad.numpy(), ref, rtol=1e-3, atol=2e-5 + 1e-4 * float(np.abs(ref).max()))
