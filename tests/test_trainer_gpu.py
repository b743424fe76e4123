"""Full training step through the C ABI vs the oracle (BASELINE config 1: dev/unet_test.py semantics at B=4)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(oracle):
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig()
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0))
    return O, cfg, flat


def _per_tensor_rel(O, cfg, g, g_ref):
    off, out = 0, []
    for name, shape in O.param_spec(cfg):
        n = int(np.prod(shape))
        a, b = g[off:off + n], g_ref[off:off + n]
        out.append((float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)), name))
        off += n
    return out


# Stated per-tensor bound of the bf16 training path: every one of the gradient tensors within 6e-2 relative L2 of the
# oracle's (fp32) gradient -- bf16 operands / activations carry 2^-9 relative rounding per element and a gradient tensor
# accumulates it over 50-100 layers of backward.  Measured worst tensors on B200 (this test prints them): 4.4e-2 at B=4,
# 3.9e-2 at B=32, 5.0e-2 for the 128x128 five-level config -- always GroupNorm scale / embedding-projection gradients of
# the 8x8 middle blocks, whose norms are 1e-4 .. 1e-5 (the deepest point of backward); the conv weights sit at <= 4e-2.
# The global bounds (all gradients: rel-L2 2e-2 stated, 3-5e-3 measured; cosine 0.9995) are the tighter ones.  Tensors
# whose oracle gradient is exactly zero (everything upstream of the zero-initialised conv2 / proj weights, dev/unet.py
# zero_module) must be exactly zero here too.
TOL_TENSOR_L2 = 6e-2


def check_grads(O, cfg, g, g_ref, what=""):
    """Global rel-L2 / cosine + per-tensor rel-L2 over every gradient tensor; prints the five worst tensors."""
    assert not np.isnan(g).any()
    cos = float(g @ g_ref / (np.linalg.norm(g) * np.linalg.norm(g_ref)))
    glob = float(np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref))
    off, rows = 0, []
    for name, shape in O.param_spec(cfg):
        n = int(np.prod(shape))
        a, b = g[off:off + n].astype(np.float64), g_ref[off:off + n].astype(np.float64)
        nb = float(np.linalg.norm(b))
        rel = float(np.linalg.norm(a - b)) / nb if nb > 0 else (0.0 if not a.any() else float("inf"))
        rows.append((rel, name, n, nb))
        off += n
    rows.sort(reverse=True)
    print(f"[grads {what}] global rel-L2 {glob:.3e} cosine {cos:.6f}; worst tensors (rel-L2, name, size, |ref|):")
    for r in rows[:5]:
        print(f"    {r[0]:.3e}  {r[1]}  n={r[2]}  |ref|={r[3]:.3e}")
    assert cos > 0.9995, cos
    assert glob <= 2e-2, glob
    assert rows[0][0] <= TOL_TENSOR_L2, rows[:5]
    return rows


def test_forward_backward_matches_oracle_B4(ub, setup):
    """out, loss and all 326 gradient tensors (dev/unet_test.cu:2082-2107 checks the same set)."""
    O, cfg, flat = setup
    B = 4
    x0, t, noise = O.synthetic_batch(cfg, B)
    tr = ub.Trainer(B=B)
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out, g = tr.get_output(), tr.get_grads()
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise)
    g_ref = g_ref.numpy()
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)          # loss: 2e-3 relative
    assert np.abs(out - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()   # bf16 activations
    check_grads(O, cfg, g, g_ref, "B=4")        # global rel-L2 2e-2, cosine 0.9995, every tensor rel-L2 5e-2
    worst = max(_per_tensor_rel(O, cfg, g, g_ref))
    assert worst[0] <= 0.15, worst                                        # every tensor: max-norm relative
    tr.close()


def test_forward_backward_matches_oracle_B32(ub, setup):
    """BASELINE configs[2] as written: the default 64x64 U-Net at batch 32, one full forward + backward against the
    oracle (a 5-6 s CPU step), loss / output / global and per-tensor gradient bounds."""
    O, cfg, flat = setup
    B = 32
    x0, t, noise = O.synthetic_batch(cfg, B, seed=4321)
    tr = ub.Trainer(B=B)
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out, g = tr.get_output(), tr.get_grads()
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise)
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    assert np.abs(out - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()
    check_grads(O, cfg, g, g_ref.numpy(), "B=32")
    tr.close()


def test_golden_fixture_B2(ub, setup, golden_dir):
    """Same comparison against the committed vectors that came from the reference's own UNetModel."""
    O, cfg, flat = setup
    g = np.load(os.path.join(golden_dir, "unet_step_B2.npz"))
    x0, t, noise = O.synthetic_batch(cfg, 2)
    tr = ub.Trainer(B=2)
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    assert abs(loss - g["loss_trace"][0]) <= 2e-3 * g["loss_trace"][0]
    out = tr.get_output().reshape(-1)[::37]
    assert np.abs(out - g["out_slice"]).max() <= 3e-2 * np.abs(g["out_slice"]).max()
    grads = tr.get_grads()
    off, norms = 0, []
    for _, s in O.param_spec(cfg):
        n = int(np.prod(s))
        norms.append(float(np.linalg.norm(grads[off:off + n].astype(np.float64))))
        off += n
    np.testing.assert_allclose(np.array(norms), g["grad_norms"], rtol=0.1)
    tr.close()


def test_ten_step_loss_trace_and_params(ub, setup):
    """10 AdamW steps (trainer hyper-parameters of train_unet.cu:5037): loss trace within 1e-2, as the reference's
    own eyeballed cpu-vs-gpu loss print (dev/unet_test.cu:2117-2132)."""
    O, cfg, flat = setup
    B = 4
    batches = [O.synthetic_batch(cfg, B, seed=1234 + i) for i in range(10)]
    losses_ref, flat_ref = O.train_steps(cfg, flat, batches, lr=1e-4)
    tr = ub.Trainer(B=B)
    tr.set_params(flat.numpy())
    losses = [tr.train_step(b[0].numpy(), b[1].numpy(), b[2].numpy(), lr=1e-4) for b in batches]
    assert np.abs(np.array(losses) - np.array(losses_ref)).max() <= 1e-2
    p = tr.get_params()
    assert np.abs(p - flat_ref.numpy()).max() <= 2.5e-3   # 10 steps x lr 1e-4 bounds any weight's travel by 1e-3
    tr.close()


def test_eager_and_graph_steps_agree(ub, setup):
    O, cfg, flat = setup
    B = 2
    x0, t, noise = O.synthetic_batch(cfg, B)
    res = []
    for graph in (0, 1):
        tr = ub.Trainer(B=B, use_cuda_graph=graph)
        tr.set_params(flat.numpy())
        l1 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy())
        l2 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy())
        res.append((l1, l2, tr.get_params()))
        tr.close()
    assert abs(res[0][0] - res[1][0]) < 1e-4 and abs(res[0][1] - res[1][1]) < 1e-3
    # two AdamW steps move a weight by at most 2*lr; a near-zero gradient whose sign differs between the runs
    # (atomic summation order) can therefore differ by 4*lr
    assert np.abs(res[0][2] - res[1][2]).max() <= 4.5e-4
    # GroupNorm statistics are accumulated with fp32 atomics in the conv epilogues: the summation order (and with it a
    # few bf16 roundings downstream) differs from run to run, so the mean drift is bounded, not zero
    assert np.abs(res[0][2] - res[1][2]).mean() < 1e-5


def test_update_matches_oracle_adamw(ub, setup):
    """unet_update / adamw_kernel2 (train_unet.cu:4720-4757) on the gradients the CUDA path produced: fp32-exact."""
    O, cfg, flat = setup
    B = 2
    x0, t, noise = O.synthetic_batch(cfg, B)
    tr = ub.Trainer(B=B)
    tr.set_params(flat.numpy())
    tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    g = torch.from_numpy(tr.get_grads())
    tr.update(lr=1e-3, weight_decay=0.01)  # dev/unet_test.cu:2108 hyper-parameters
    p = tr.get_params()
    p_ref, _, _ = O.adamw_step(flat, g, torch.zeros_like(flat), torch.zeros_like(flat), 1, lr=1e-3, wd=0.01)
    np.testing.assert_allclose(p, p_ref.numpy(), rtol=0, atol=2e-6)
    assert np.abs(tr.get_grads()).max() == 0.0   # update zeroes the gradient arena (unet_zero_grad)
    tr.close()


def test_device_noise_statistics(ub, setup):
    """Device-side timestep / noise draws (Philox) replace cuRAND (train_unet.cu:3115-3254): sanity of the moments
    through the loss of an untrained net, and different steps draw different noise."""
    O, cfg, flat = setup
    tr = ub.Trainer(B=8)
    tr.set_params(flat.numpy())
    x0 = (torch.rand(8, 3, 64, 64) * 2 - 1).numpy()
    l1 = tr.forward_backward(x0)
    l2 = tr.forward_backward(x0)
    assert 0.8 < l1 < 1.5 and 0.8 < l2 < 1.5    # E[(out - eps)^2] ~ 1 for eps ~ N(0,1) and a near-zero initial output
    tr.close()


def test_random_flip_matches_oracle(ub, setup):
    """On-GPU augmentation (cfg.random_flip; the PyTorch loader's random_flip, train_unet.py:508-536): the step must
    equal the oracle's step on the batch mirrored by the same coins, the coins are fair and change from step to step."""
    O, cfg, flat = setup
    B = 8
    x0, t, noise = O.synthetic_batch(cfg, B, seed=77)
    tr = ub.Trainer(B=B, random_flip=1)
    tr.set_params(flat.numpy())
    seen = []
    for step in range(3):
        loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
        flips = tr.get_flips()
        assert set(np.unique(flips)) <= {0, 1}
        seen.append(flips.copy())
        loss_ref, out_ref, _ = O.train_step_grads(cfg, flat, O.random_flip(x0, flips), t, noise)
        assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref), (step, flips)
        assert np.abs(tr.get_output() - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()
        tr.update(lr=0.0)   # advances the step counter that keys the draws; lr 0 keeps the parameters
    allf = np.concatenate(seen)
    assert 0 < allf.sum() < allf.size                      # both outcomes occur in 24 fair coins (p = 1 - 2^-23)
    assert any((seen[0] != s).any() for s in seen[1:])     # not the same coins every step
    tr.close()
    # without the flag nothing is flipped and the call says so
    tr = ub.Trainer(B=2)
    with pytest.raises(ub.UbError):
        tr.get_flips()
    tr.close()


def test_checkpoint_roundtrip_and_reference_layout(ub, setup, tmp_path):
    O, cfg, flat = setup
    tr = ub.Trainer(B=2)
    ref_file = str(tmp_path / "unet_init.bin")
    O.write_model_bin(ref_file, cfg, flat.numpy(), B=32)     # the layout train_unet.py:768-795 writes
    tr.load(ref_file)
    np.testing.assert_array_equal(tr.get_params(), flat.numpy())
    x0, t, noise = O.synthetic_batch(cfg, 2)
    tr.train_step(x0.numpy(), t.numpy(), noise.numpy())
    out_file = str(tmp_path / "model_1.bin")
    tr.save(out_file, with_adamw=True)
    header, rest = O.read_model_bin(out_file)                 # generate.py:17-27 reads header + params
    n = tr.nparams
    assert list(header[:10]) == [12345678, 2, 3, 64, 3, 64, 64, 1000, 1, 0] and header[10] == 1
    assert header[11] == 0x55425354                           # marks header[10] (AdamW step count) as valid
    assert rest.size == 3 * n and os.path.getsize(out_file) == 1024 + 12 * n
    np.testing.assert_array_equal(rest[:n], tr.get_params())
    tr2 = ub.Trainer(B=2)
    tr2.load(out_file)                                        # resume: params + m + v + step counter
    b = O.synthetic_batch(cfg, 2, seed=7)
    la = tr.train_step(b[0].numpy(), b[1].numpy(), b[2].numpy())
    lb = tr2.train_step(b[0].numpy(), b[1].numpy(), b[2].numpy())
    assert abs(la - lb) < 1e-3
    assert np.abs(tr.get_params() - tr2.get_params()).max() <= 2.5e-4   # one step: 2*lr + rounding
    tr.close(), tr2.close()


def test_resume_from_reference_written_checkpoint_with_garbage_header(ub, setup, tmp_path):
    """The reference's C writer fills words 0..9 of an uninitialised int[256] (train_unet.cu:4764): words 10.. are stack
    garbage.  Loading such a checkpoint (with AdamW state) must not take the step count from it: a negative or huge count
    makes the bias correction NaN / a no-op.  The count restarts at 0, like the reference's own resume."""
    O, cfg, flat = setup
    tr = ub.Trainer(B=2)
    tr.set_params(flat.numpy())
    x0, t, noise = O.synthetic_batch(cfg, 2)
    tr.train_step(x0.numpy(), t.numpy(), noise.numpy())
    good = str(tmp_path / "good.bin")
    tr.save(good, with_adamw=True)
    raw = np.fromfile(good, dtype=np.int32).copy()
    rng = np.random.default_rng(3)
    raw[10:256] = rng.integers(-2**31, 2**31 - 1, size=246, dtype=np.int64).astype(np.int32)
    raw[10] = -123456789                                     # the dangerous case: 1 - beta2^t < 0 -> sqrt -> NaN
    bad = str(tmp_path / "reference_written.bin")
    raw.tofile(bad)
    tr2 = ub.Trainer(B=2)
    tr2.load(bad)
    l = tr2.train_step(x0.numpy(), t.numpy(), noise.numpy())
    p = tr2.get_params()
    assert np.isfinite(l) and np.isfinite(p).all()
    # step count 0 -> first-step bias correction: |update| = lr for every weight with a non-zero gradient
    assert np.abs(p - tr.get_params()).max() <= 2.5e-4
    assert ub.lib().ub_trainer_set_step(tr2._h, 7) == 0 and ub.lib().ub_trainer_set_step(tr2._h, -1) != 0
    tr.close(), tr2.close()


def test_predict_matches_oracle_forward(ub, setup):
    """Forward only (what generate.py:29-52 needs from the checkpoint consumer side)."""
    O, cfg, flat = setup
    P = O.unflatten_params(cfg, flat)
    xt = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    t = torch.tensor([[10.0], [900.0]])
    tr = ub.Trainer(B=2)
    tr.set_params(flat.numpy())
    out = tr.predict(xt.numpy(), t.numpy())
    with torch.no_grad():
        ref = O.unet_forward(cfg, P, xt, t).numpy()
    assert np.abs(out - ref).max() <= 3e-2 * np.abs(ref).max()
    tr.close()


def test_full_size_properties_B32(ub, setup):
    """BASELINE config 3 size (B=32): size-independent properties instead of an oracle run --
    (1) the loss of a batch made of 8 copies of a B=4 batch equals the B=4 loss, and its gradient equals the B=4
    gradient (mean reduction), (2) training on a fixed batch decreases the loss."""
    O, cfg, flat = setup
    x0, t, noise = O.synthetic_batch(cfg, 4)
    tr4 = ub.Trainer(B=4)
    tr4.set_params(flat.numpy())
    l4 = tr4.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    g4 = tr4.get_grads()
    tr4.close()
    rep = lambda a: np.concatenate([a.numpy()] * 8, axis=0)
    tr = ub.Trainer(B=32)
    tr.set_params(flat.numpy())
    l32 = tr.forward_backward(rep(x0), rep(t), rep(noise))
    g32 = tr.get_grads()
    assert abs(l32 - l4) < 1e-4
    assert np.linalg.norm(g32 - g4) <= 5e-3 * np.linalg.norm(g4)
    losses = [tr.train_step(rep(x0), rep(t), rep(noise), lr=1e-4) for _ in range(6)]
    assert losses[-1] < losses[0]
    tr.close()


def test_config5_128px_five_levels_matches_oracle(ub, oracle):
    """BASELINE config 5 architecture: 128x128 input, channel_mult (1,1,2,3,4), attention at
    16x16 and 8x8 -- outside the reference CUDA trainer's hard-coded 4 levels, inside dev/unet.py's envelope.  One
    forward+backward at B=2 against the oracle; same bounds as the 64x64 test."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(channel_mult=(1, 1, 2, 3, 4), attn_start_level=3, H=128, W=128)
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0))
    assert flat.numel() == 21082755
    x0, t, noise = O.synthetic_batch(cfg, 2)
    tr = ub.Trainer(B=2, H=128, W=128, channel_mult=(1, 1, 2, 3, 4), att_start_level=3)
    assert tr.nparams == flat.numel()
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    g = tr.get_grads()
    loss_ref, _, g_ref = O.train_step_grads(cfg, flat, x0, t, noise)
    g_ref = g_ref.numpy()
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    check_grads(O, cfg, g, g_ref, "config5 128px")
    tr.close()


def test_sampler_matches_oracle(ub, setup):
    """ub_trainer_sample (generate.py:29-79 on the device) vs oracle.ddpm_sample with the same injected draws: three
    iterations at small t (where the update weights eps the most), B = 2; then a run with device-side noise."""
    O, cfg, flat = setup
    P = O.unflatten_params(cfg, flat)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 64, 64, generator=g)
    noises = [torch.randn(2, 3, 64, 64, generator=g) for _ in range(3)]
    ref = O.ddpm_sample(cfg, P, x, 40, 38, noises).numpy()
    tr = ub.Trainer(B=2)
    tr.set_params(flat.numpy())
    out = tr.sample(x.numpy(), 40, 38, torch.stack(noises).numpy())
    assert np.abs(out - ref).max() <= 3e-2 * np.abs(ref).max()
    # device-side draws: finite, roughly unit scale after a few steps from N(0,1); a fixed seed reproduces the run up
    # to the summation order of the GroupNorm-statistics atomics (bf16 roundings downstream), a new seed does not
    a = tr.sample(None, 999, 990, None, seed=5)
    b = tr.sample(None, 999, 990, None, seed=5)
    c = tr.sample(None, 999, 990, None, seed=6)
    assert np.isfinite(a).all() and 0.5 < a.std() < 2.0
    assert np.abs(a - b).max() < 2e-2 and np.abs(a - c).max() > 0.5
    tr.close()


def test_train_from_dataloader(ub, setup, tmp_path):
    """The reference's loop (train_unet.cu:5019-5037) end to end: prepare_data.py-format file -> prefetching reader ->
    pinned buffer -> train step; the loss of a step on a loader batch equals the loss on the same images passed in."""
    O, cfg, flat = setup
    rng = np.random.default_rng(0)
    imgs = rng.uniform(-1, 1, (12, 3, 64, 64)).astype(np.float32)
    path = str(tmp_path / "train.bin")
    O.write_data_bin(path, imgs)
    dl = ub.DataLoader(path, B=4)
    _, t, noise = O.synthetic_batch(cfg, 4)
    tr = ub.Trainer(B=4)
    tr.set_params(flat.numpy())
    l_direct = tr.forward_backward(imgs[0:4], t.numpy(), noise.numpy())
    import ctypes as C
    loss = C.c_float()
    rc = ub.lib().ub_trainer_forward_backward(tr._h, C.cast(dl.next_ptr(), C.POINTER(C.c_float)),
                                              t.numpy().ctypes.data_as(C.POINTER(C.c_float)),
                                              noise.numpy().ctypes.data_as(C.POINTER(C.c_float)), C.byref(loss))
    assert rc == 0 and abs(loss.value - l_direct) < 1e-3
    tr.close(), dl.close()


def test_next_batch_prefetch(ub, setup):
    """ub_trainer_set_next_batch: the announced batch's H2D copy runs on the copy stream under the current step; the next
    train_step called with that pointer must consume exactly that batch (bit-for-bit on the device), a step called with
    another pointer must fall back to the ordinary copy, and the losses must be those of the synchronous path."""
    import ctypes as C
    O, cfg, flat = setup
    B = 2
    g = torch.Generator().manual_seed(5)
    params = (flat + 0.02 * torch.randn(flat.shape, generator=g)).numpy()  # (zero-initialised layers made non-zero)
    batches = [((torch.rand(B, 3, 64, 64, generator=g) * 2 - 1) * s).pin_memory() for s in (0.1, 1.0, 0.5, 0.8, 0.3)]

    def run(announce):
        tr = ub.Trainer(B=B)
        tr.set_params(params)
        loss, losses = C.c_float(), []
        for i in range(len(batches)):
            if announce is not None:
                tr.set_next_batch(batches[announce[i]].data_ptr())
            tr.train_step_ptr(batches[i].data_ptr(), loss_ref=C.byref(loss))
            np.testing.assert_array_equal(tr.get_batch(), batches[i].numpy())
            losses.append(loss.value)
        tr.close()
        return np.array(losses)

    plain = run(None)
    hit = run([1, 2, 3, 4, 0])    # step i announces batch i + 1: every step from the second on starts from the prefetch
    miss = run([2, 3, 4, 0, 1])   # always announces a batch the next step does not use: ordinary path, prefetch dropped
    mixed = run([1, 0, 3, 1, 2])  # hit, miss, hit, miss
    assert not np.isnan(plain).any()
    # (fp32 atomics: two runs of the same steps differ by summation order, bounded as in test_eager_and_graph_steps_agree)
    for other in (hit, miss, mixed):
        assert np.abs(other - plain).max() < 2e-3, (plain, other)
    # an un-pinned announcement is ignored
    tr = ub.Trainer(B=B)
    tr.set_params(params)
    pageable = batches[1].clone()
    tr.set_next_batch(pageable.data_ptr())
    loss = C.c_float()
    tr.train_step_ptr(batches[0].data_ptr(), loss_ref=C.byref(loss))
    tr.train_step_ptr(pageable.data_ptr(), loss_ref=C.byref(loss))
    np.testing.assert_array_equal(tr.get_batch(), pageable.numpy())
    tr.close()


def test_dinput_matches_oracle(ub, setup):
    """dL/d(x_t), the last tensor dev/unet_test.cu:2082-2107 compares (unet_backward writes it into its dinp buffer)."""
    O, cfg, flat = setup
    P = O.unflatten_params(cfg, flat)
    x0, t, noise = O.synthetic_batch(cfg, 2)
    xt = O.q_sample(x0, t, noise).requires_grad_(True)
    loss = O.mse_loss(O.unet_forward(cfg, P, xt, t), noise)
    loss.backward()
    tr = ub.Trainer(B=2, compute_dinput=1)
    tr.set_params(flat.numpy())
    tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    d = tr.get_dinput()
    ref = xt.grad.numpy()
    assert np.abs(d - ref).max() <= 3e-2 * np.abs(ref).max()
    tr.close()
    tr = ub.Trainer(B=2)
    with pytest.raises(ub.UbError):
        tr.get_dinput()
    tr.close()


@pytest.mark.parametrize("kw,okw,B", [
    (dict(H=32, W=32, channel_mult=(1, 2), att_start_level=1), dict(H=32, W=32, channel_mult=(1, 2), attn_start_level=1), 3),
    (dict(C_model=128, channel_mult=(1, 2, 2), att_start_level=2, n_res_blocks=1, H=32, W=32),
     dict(model_channels=128, channel_mult=(1, 2, 2), attn_start_level=2, num_res_blocks=1, H=32, W=32), 2),
    (dict(H=64, W=32, channel_mult=(1, 2, 4), att_start_level=1, n_res_blocks=1),
     dict(H=64, W=32, channel_mult=(1, 2, 4), attn_start_level=1, num_res_blocks=1), 2),
])
def test_other_configs_match_oracle(ub, oracle, kw, okw, B):
    """Config generality: other depths, widths, block counts, a non-square image, attention
    at 16x16 (T=256), 8x8 (T=64) and a 16x8 map (T=128), an odd batch -- one forward+backward against the oracle."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(**okw)
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0))
    x0, t, noise = O.synthetic_batch(cfg, B)
    tr = ub.Trainer(B=B, **kw)
    assert tr.nparams == flat.numel()
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    g = tr.get_grads()
    loss_ref, _, g_ref = O.train_step_grads(cfg, flat, x0, t, noise)
    g_ref = g_ref.numpy()
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    check_grads(O, cfg, g, g_ref, str(kw))
    l2 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-4)   # the captured-graph path runs too
    assert np.isfinite(l2)
    tr.close()


def test_class_conditional_matches_oracle(ub, oracle, golden_dir):
    """class-conditional embedding (dev/unet.py:174-175, 301-303).  Weights = the initialisation
    with the zero-initialised tensors perturbed (otherwise the embedding path receives no gradient); loss, output and
    every gradient against the oracle, the label-embedding gradient also against the fixture the reference's own
    UNetModel(num_classes=10) produced; rows of unused classes stay exactly zero; predict() takes the labels too."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(num_classes=10)
    flat = O.perturb_zero_params(cfg, O.flatten_params(cfg, O.init_params(cfg, seed=0)))
    gold = np.load(os.path.join(golden_dir, "class_cond_B2.npz"))
    B = 2
    x0, t, noise = O.synthetic_batch(cfg, B)
    y = gold["labels"]
    tr = ub.Trainer(B=B, num_classes=10)
    assert tr.nparams == flat.numel() == 20494211 + 10 * 256
    tr.set_params(flat.numpy())
    tr.set_labels(y)
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out, g = tr.get_output(), tr.get_grads()
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise, torch.from_numpy(y))
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    assert abs(loss - float(gold["loss"][0])) <= 2e-3 * float(gold["loss"][0])
    assert np.abs(out - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()
    rows = {name: rel for rel, name, _, _ in check_grads(O, cfg, g, g_ref.numpy(), "class-conditional")}
    assert rows["label_emb.weight"] <= 2e-2, rows["label_emb.weight"]
    off = sum(int(np.prod(s)) for _, s in O.param_spec(cfg)[:4])
    ge = g[off:off + 2560].reshape(10, 256)
    ref = gold["label_emb_grad"]
    assert np.linalg.norm(ge - ref) <= 2e-2 * np.linalg.norm(ref)
    unused = [k for k in range(10) if k not in set(y.tolist())]
    assert np.abs(ge[unused]).max() == 0.0
    # other labels -> another output; the same labels through predict() -> the oracle's forward
    xt = O.q_sample(x0, t, noise)
    with torch.no_grad():
        P = O.unflatten_params(cfg, flat)
        o_same = O.unet_forward(cfg, P, xt, t, torch.from_numpy(y)).numpy()
        o_other = O.unet_forward(cfg, P, xt, t, torch.tensor([1, 2])).numpy()
    assert np.abs(tr.predict(xt.numpy(), t.numpy()) - o_same).max() <= 3e-2 * np.abs(o_same).max()
    tr.set_labels([1, 2])
    assert np.abs(tr.predict(xt.numpy(), t.numpy()) - o_other).max() <= 3e-2 * np.abs(o_other).max()
    # the captured-graph training step reads the label buffer too
    l2 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-4)
    assert np.isfinite(l2)
    with pytest.raises(ub.UbError):
        tr.set_labels([1, 10])
    tr.close()
    plain = ub.Trainer(B=B)
    with pytest.raises(ub.UbError):
        plain.set_labels([0, 1])
    plain.close()


@pytest.mark.parametrize("graph", [0, 1])
def test_ema_matches_oracle(ub, setup, tmp_path, graph):
    """EMA of the parameters fused into the AdamW kernel (guided-diffusion update_ema,
    train_unet.py:708): after every step ema == rate * ema + (1 - rate) * params, fp32-exact on the parameters the CUDA
    path itself produced; saved as a parameters-only checkpoint of the reference layout."""
    O, cfg, flat = setup
    B, rate = 2, 0.9
    x0, t, noise = O.synthetic_batch(cfg, B)
    tr = ub.Trainer(B=B, ema_rate=rate, use_cuda_graph=graph)
    tr.set_params(flat.numpy())
    np.testing.assert_array_equal(tr.get_ema(), flat.numpy())      # starts as a copy of the parameters
    ema = flat.clone()
    for _ in range(3):
        tr.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-3)
        ema = O.ema_update(ema, torch.from_numpy(tr.get_params()), rate)
        np.testing.assert_allclose(tr.get_ema(), ema.numpy(), rtol=0, atol=1e-6)
    assert np.abs(tr.get_ema() - tr.get_params()).max() > 1e-4
    path = str(tmp_path / "ema.bin")
    tr.save_ema(path)
    hdr, payload = O.read_model_bin(path)[:2]
    assert int(hdr[0]) == O.MODEL_MAGIC and int(hdr[8]) == 0
    np.testing.assert_array_equal(np.asarray(payload[:flat.numel()]), tr.get_ema())
    tr.set_ema(flat.numpy())
    np.testing.assert_array_equal(tr.get_ema(), flat.numpy())
    tr.close()
    plain = ub.Trainer(B=B)
    with pytest.raises(ub.UbError):
        plain.get_ema()
    plain.close()


@pytest.mark.parametrize("kw,okw,B", [
    (dict(resblock_updown=1), dict(resblock_updown=True), 2),
    (dict(resblock_updown=1, H=32, W=32, channel_mult=(1, 2, 2), att_start_level=1, n_res_blocks=1),
     dict(resblock_updown=True, H=32, W=32, channel_mult=(1, 2, 2), attn_start_level=1, num_res_blocks=1), 3),
])
def test_resblock_updown_matches_oracle(ub, oracle, golden_dir, kw, okw, B):
    """ResBlock(down=True) / ResBlock(up=True) in place of Downsample / Upsample
    (dev/unet.py:147,205-222,271-284; dev/resblock.py:78-86,125-128).  Perturbed weights (every path carries a gradient);
    loss, output, all gradients against the oracle (pinned to the reference's UNetModel(resblock_updown=True) live and by
    fixture); the default-shape case also against the fixture's loss and per-tensor gradient norms."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(**okw)
    flat = O.perturb_zero_params(cfg, O.flatten_params(cfg, O.init_params(cfg, seed=0)))
    x0, t, noise = O.synthetic_batch(cfg, B)
    tr = ub.Trainer(B=B, **kw)
    assert tr.nparams == flat.numel()
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out, g = tr.get_output(), tr.get_grads()
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise)
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    assert np.abs(out - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()
    check_grads(O, cfg, g, g_ref.numpy(), "resblock_updown " + str(kw))
    if len(kw) == 1:
        gold = np.load(os.path.join(golden_dir, "updown_B2.npz"))
        assert abs(loss - float(gold["loss"][0])) <= 2e-3 * float(gold["loss"][0])
        off, norms = 0, []
        for _, s in O.param_spec(cfg):
            n = int(np.prod(s))
            norms.append(float(np.linalg.norm(g[off:off + n].astype(np.float64))))
            off += n
        np.testing.assert_allclose(np.array(norms), gold["grad_norms"], rtol=0.1)
    l2 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-4)   # the captured-graph path runs too
    assert np.isfinite(l2)
    tr.close()


@pytest.mark.parametrize("kw,okw,B", [
    (dict(use_scale_shift_norm=1), dict(use_scale_shift_norm=True), 2),
    (dict(use_scale_shift_norm=1, resblock_updown=1, num_classes=5, H=32, W=32, channel_mult=(1, 2, 2), att_start_level=1,
          n_res_blocks=1),
     dict(use_scale_shift_norm=True, resblock_updown=True, num_classes=5, H=32, W=32, channel_mult=(1, 2, 2),
          attn_start_level=1, num_res_blocks=1), 3),
])
def test_scale_shift_norm_matches_oracle(ub, oracle, kw, okw, B):
    """use_scale_shift_norm (dev/unet.py:146, dev/resblock.py:211,243-247): every ResBlock's
    embedding projection yields [scale | shift] and GroupNorm 2 computes gn(h) * (1 + scale) + shift.  Perturbed weights;
    loss, output and all gradients -- the doubled embedding projections included -- against the oracle, whose ResBlock is
    pinned to the reference's ResBlockO by fixture; the second case combines it with resblock_updown and class labels."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(**okw)
    flat = O.perturb_zero_params(cfg, O.flatten_params(cfg, O.init_params(cfg, seed=0)))
    x0, t, noise = O.synthetic_batch(cfg, B)
    y = np.arange(B) % 5 if cfg.num_classes else None
    tr = ub.Trainer(B=B, **kw)
    assert tr.nparams == flat.numel()
    tr.set_params(flat.numpy())
    if y is not None:
        tr.set_labels(y)
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out, g = tr.get_output(), tr.get_grads()
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise, None if y is None else torch.from_numpy(y))
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    assert np.abs(out - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()
    rows = {name: rel for rel, name, _, _ in check_grads(O, cfg, g, g_ref.numpy(), "scale-shift " + str(kw))}
    # (the doubled embedding projections are gradient tensors like any other: measured worst 4.3e-2, a middle-block l_emb
    #  whose gradient norm is ~1e-4 -- inside the per-tensor bound check_grads applies to all of them)
    assert max(v for k, v in rows.items() if ".l_emb." in k) <= TOL_TENSOR_L2
    l2 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-4)   # the captured-graph path runs too
    assert np.isfinite(l2)
    tr.close()


def _resblock_shapes(O, cfg):
    """(C, H, W) of every ResBlock's second-GroupNorm tensor, forward order (= the dropout masks' shapes)."""
    out = []
    for l in O.build_layers(cfg):
        if l.kind.startswith("res"):
            lev = l.level + (1 if l.kind == "res_down" else (-1 if l.kind == "res_up" else 0))
            out.append((l.cout, cfg.H >> lev, cfg.W >> lev))
    return out


@pytest.mark.parametrize("kw,okw,B", [
    (dict(dropout=0.1), dict(dropout=0.1), 2),
    (dict(dropout=0.25, use_scale_shift_norm=1, resblock_updown=1, H=32, W=32, channel_mult=(1, 2, 2), att_start_level=1,
          n_res_blocks=1),
     dict(dropout=0.25, use_scale_shift_norm=True, resblock_updown=True, H=32, W=32, channel_mult=(1, 2, 2),
          attn_start_level=1, num_res_blocks=1), 3),
])
def test_dropout_matches_oracle_with_the_device_masks(ub, oracle, kw, okw, B):
    """dropout between SiLU and conv2 of every ResBlock (guided-diffusion's out_layers; the
    reference carries the option commented out, dev/resblock.py:51,61 -- PARITY UNPINNED against the reference for this
    one feature: the oracle restates nn.Dropout's training-mode arithmetic, keep / (1 - p), with the masks the device
    drew).  The masks are Philox draws that are never stored; they are regenerated for the check, must keep ~(1 - p) of the
    elements, must differ between blocks and steps, and predict() must run without dropout."""
    O = oracle
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.UNetConfig(**okw)
    p = cfg.dropout
    flat = O.perturb_zero_params(cfg, O.flatten_params(cfg, O.init_params(cfg, seed=0)))
    x0, t, noise = O.synthetic_batch(cfg, B)
    tr = ub.Trainer(B=B, **kw)
    tr.set_params(flat.numpy())
    loss = tr.forward_backward(x0.numpy(), t.numpy(), noise.numpy())
    out, g = tr.get_output(), tr.get_grads()
    shapes = _resblock_shapes(O, cfg)
    masks = [tr.get_dropout_mask(i, *s) for i, s in enumerate(shapes)]
    total = sum(m.size for m in masks)
    kept = sum(int(m.sum()) for m in masks)
    assert abs(kept / total - (1 - p)) < 5e-3, kept / total
    assert all(set(np.unique(m)) <= {0, 1} for m in masks)
    same_shape = [(a, b) for i, a in enumerate(masks) for b in masks[i + 1:] if a.shape == b.shape]
    assert same_shape and all((a != b).mean() > 0.5 * p for a, b in same_shape)      # blocks draw different masks
    loss_ref, out_ref, g_ref = O.train_step_grads(cfg, flat, x0, t, noise,
                                                  drop_masks=[torch.from_numpy(m) for m in masks])
    assert abs(loss - float(loss_ref)) <= 2e-3 * float(loss_ref)
    assert np.abs(out - out_ref.numpy()).max() <= 3e-2 * np.abs(out_ref.numpy()).max()
    check_grads(O, cfg, g, g_ref.numpy(), "dropout " + str(kw))
    # without the masks the oracle is measurably elsewhere (the check above is not vacuous)
    _, _, g_nodrop = O.train_step_grads(cfg, flat, x0, t, noise)
    gap = float(np.linalg.norm(g_ref.numpy() - g_nodrop.numpy()) / np.linalg.norm(g_ref.numpy()))
    print(f"[dropout] oracle gradient with vs without the masks: rel-L2 {gap:.3f}")
    assert gap > 3 * 2e-2, gap     # several times the global gradient bound of check_grads
    # inference: no dropout
    xt = O.q_sample(x0, t, noise)
    with torch.no_grad():
        o_inf = O.unet_forward(cfg, O.unflatten_params(cfg, flat), xt, t).numpy()
    assert np.abs(tr.predict(xt.numpy(), t.numpy()) - o_inf).max() <= 3e-2 * np.abs(o_inf).max()
    # next step (captured graph path): other masks
    tr.update(lr=1e-4)
    l2 = tr.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-4)
    assert np.isfinite(l2)
    m2 = tr.get_dropout_mask(0, *shapes[0])
    assert (m2 != masks[0]).mean() > 0.5 * p
    tr.close()
