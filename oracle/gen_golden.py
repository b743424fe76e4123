"""Generate the committed golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py
It imports the reference's own Python modules (dev/unet.py -> UNetModel/AttentionBlock, dev/resblock.py -> ResBlock,
train_unet.py -> GaussianDiffusion / get_named_beta_schedule / save_model_params_to_bin), runs them on seeded
synthetic inputs on CPU and stores small slices of the results.  tests/test_oracle_golden.py then pins
oracle/unet_oracle.py against these vectors anywhere (the reference tree does not travel to the GPU box).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "dev"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import unet_oracle as O  # noqa: E402


def synthetic(B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (B, 1), generator=g).float()
    noise = torch.randn(B, 3, 64, 64, generator=g)
    return x0, t, noise


def main():
    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    from unet import UNetModel, AttentionBlock  # reference dev/unet.py
    from resblock import ResBlock  # reference dev/resblock.py
    import train_unet as ref_train  # reference train_unet.py (GaussianDiffusion, bin writer)

    # ---------------- full model: default 64x64 U-Net, B = 2
    B = 2
    torch.manual_seed(0)
    model = UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32)
    names = [n for n, _ in model.named_parameters()]
    shapes = [tuple(p.shape) for _, p in model.named_parameters()]
    flat0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    diffusion = ref_train.GaussianDiffusion(ref_train.get_named_beta_schedule("linear", 1000))
    x0, t, noise = synthetic(B)
    x_t = diffusion.q_sample(x0, t.view(B).long(), noise)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    losses, first = [], {}
    for step in range(3):
        xs, ts, ns = synthetic(B, seed=1234 + step)
        xt = diffusion.q_sample(xs, ts.view(B).long(), ns)
        opt.zero_grad()
        out = model(xt, ts)  # (B,1) timesteps as dev/unet_test.py:298
        loss = ((out - ns) ** 2).mean()
        loss.backward()
        if step == 0:
            first["out"] = out.detach().clone()
            first["grads"] = [p.grad.detach().clone() for p in model.parameters()]
        opt.step()
        losses.append(float(loss))
    flat3 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gflat = torch.cat([g.reshape(-1) for g in first["grads"]])
    np.savez_compressed(
        os.path.join(out_dir, "unet_step_B2.npz"),
        names=np.array(names), shapes=np.array([str(s) for s in shapes]),
        param_sums=np.array([float(flat0[o:o + n].double().sum()) for o, n in offsets(shapes)]),
        param_head=np.stack([pad4(flat0[o:o + n][:4]) for o, n in offsets(shapes)]),
        x_t_slice=x_t.reshape(-1)[::997].numpy(),
        sqrt_ac=diffusion.sqrt_alphas_cumprod.astype(np.float32),
        sqrt_1mac=diffusion.sqrt_one_minus_alphas_cumprod.astype(np.float32),
        out_slice=first["out"].reshape(-1)[::37].numpy(),
        loss_trace=np.array(losses, dtype=np.float64),
        grad_norms=np.array([float(g.double().norm()) for g in first["grads"]]),
        grad_head=np.stack([pad4(g.reshape(-1)[:4]) for g in first["grads"]]),
        grad_slice=gflat[::4099].numpy(),
        params_after3_slice=flat3[::4099].numpy(),
    )

    # .bin written by the reference's own writer: keep header + a strided slice of the payload
    with tempfile.TemporaryDirectory() as d:
        torch.manual_seed(0)
        m2 = UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32)
        path = os.path.join(d, "unet_init.bin")
        ref_train.save_model_params_to_bin(m2, path)
        raw = np.fromfile(path, dtype=np.int32, count=256)
        payload = np.fromfile(path, dtype=np.float32, offset=1024)
        np.savez_compressed(os.path.join(out_dir, "model_bin.npz"), header=raw, n_floats=np.array([payload.size]),
                            file_bytes=np.array([os.path.getsize(path)]), payload_slice=payload[::4099])

    # ---------------- layers: the reference's own ResBlock / AttentionBlock modules on small inputs
    torch.manual_seed(0)
    rb = ResBlock(64, 256, out_channels=32)  # channel-changing block (1x1 skip conv)
    x = torch.randn(2, 64, 8, 8)
    emb = torch.randn(2, 256)
    y = rb(x, emb)
    torch.manual_seed(1)
    ab = AttentionBlock(64, HS=32)
    xa = torch.randn(2, 64, 4, 4)
    ya = ab(xa)
    sd = {"rb." + k: v.detach().numpy() for k, v in rb.state_dict().items()}
    sd.update({"ab." + k: v.detach().numpy() for k, v in ab.state_dict().items()})
    np.savez_compressed(os.path.join(out_dir, "layers.npz"), rb_x=x.numpy(), rb_emb=emb.numpy(),
                        rb_y=y.detach().numpy(), ab_x=xa.numpy(), ab_y=ya.detach().numpy(), **sd)
    from unet import timestep_embedding
    tt = torch.tensor([[0.0], [1.0], [17.0], [999.0]])
    np.savez_compressed(os.path.join(out_dir, "timestep_embedding.npz"), t=tt.numpy(),
                        emb=timestep_embedding(tt, 64).numpy())
    print("golden fixtures written to", out_dir, "losses", losses)


def gen_class_cond(out_dir):
    """Class-conditional model (dev/unet.py num_classes=10), zero-initialised tensors perturbed (oracle.perturb_zero_params
    recipe applied to the reference model's own parameters), B = 2, labels (3, 7): loss, output slice, gradient slice,
    per-tensor gradient norms and the full label-embedding gradient.   python oracle/gen_golden.py class_cond"""
    from unet import UNetModel
    import train_unet as ref_train
    torch.manual_seed(0)
    model = UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32, num_classes=10)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if not p.detach().any():
                p.add_(0.02 * torch.randn(p.shape, generator=g))
    shapes = [tuple(p.shape) for _, p in model.named_parameters()]
    diffusion = ref_train.GaussianDiffusion(ref_train.get_named_beta_schedule("linear", 1000))
    B = 2
    x0, t, noise = synthetic(B)
    y = torch.tensor([3, 7])
    out = model(diffusion.q_sample(x0, t.view(B).long(), noise), t, y)
    loss = ((out - noise) ** 2).mean()
    loss.backward()
    gflat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    np.savez_compressed(
        os.path.join(out_dir, "class_cond_B2.npz"), labels=y.numpy(), loss=np.array([float(loss)]),
        out_slice=out.detach().reshape(-1)[::37].numpy(), grad_slice=gflat[::4099].numpy(),
        grad_norms=np.array([float(gflat[o:o + n].double().norm()) for o, n in offsets(shapes)]),
        label_emb_grad=model.label_emb.weight.grad.numpy(),
        param_slice=torch.cat([p.detach().reshape(-1) for p in model.parameters()])[::4099].numpy())
    print("class_cond_B2.npz written, loss", float(loss))


_RBO_NAMES = {"in_layers.0": "gn1", "in_layers.2": "cv3_1", "emb_layers.1": "l_emb", "out_layers.0": "gn2",
              "out_layers.2": "cv3_2", "skip_connection": "skip_connection"}


def gen_scale_shift(out_dir):
    """use_scale_shift_norm on the reference's copy of the original ResBlock (ResBlockO, dev/resblock.py:165-247 -- the
    explicit ResBlock's own scale-shift branch refers to an attribute it does not define): two blocks, 64 -> 32 plain and
    64 -> 64 with down=True; inputs, outputs and the gradients of x, emb and every parameter for a fixed dL/dy.
    python oracle/gen_golden.py scale_shift"""
    from resblock import ResBlockO
    out = {}
    for tag, kw, cout in (("a", {}, 32), ("d", {"down": True}, 64)):
        torch.manual_seed(3)
        rb = ResBlockO(64, 256, out_channels=cout, use_scale_shift_norm=True, **kw)
        x = torch.randn(2, 64, 8, 8, requires_grad=True)
        emb = torch.randn(2, 256, requires_grad=True)
        y = rb(x, emb)
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(4))
        y.backward(dy)
        out.update({f"{tag}_x": x.detach().numpy(), f"{tag}_emb": emb.detach().numpy(), f"{tag}_y": y.detach().numpy(),
                    f"{tag}_dy": dy.numpy(), f"{tag}_dx": x.grad.numpy(), f"{tag}_demb": emb.grad.numpy()})
        for k, v in rb.named_parameters():
            mod, leaf = k.rsplit(".", 1)
            out[f"{tag}.{_RBO_NAMES[mod]}.{leaf}"] = v.detach().numpy()
            out[f"{tag}.grad.{_RBO_NAMES[mod]}.{leaf}"] = v.grad.numpy()
    np.savez_compressed(os.path.join(out_dir, "resblock_scale_shift.npz"), **out)
    print("resblock_scale_shift.npz written")


def gen_updown(out_dir):
    """resblock_updown=True (dev/unet.py:147,205-222,271-284), zero-initialised tensors perturbed as in gen_class_cond,
    B = 2: loss, output slice, gradient slice, per-tensor gradient norms.   python oracle/gen_golden.py updown"""
    from unet import UNetModel
    import train_unet as ref_train
    torch.manual_seed(0)
    model = UNetModel(3, 64, 3, 2, (4, 8), num_head_channels=32, resblock_updown=True)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if not p.detach().any():
                p.add_(0.02 * torch.randn(p.shape, generator=g))
    shapes = [tuple(p.shape) for _, p in model.named_parameters()]
    diffusion = ref_train.GaussianDiffusion(ref_train.get_named_beta_schedule("linear", 1000))
    B = 2
    x0, t, noise = synthetic(B)
    out = model(diffusion.q_sample(x0, t.view(B).long(), noise), t)
    loss = ((out - noise) ** 2).mean()
    loss.backward()
    gflat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    np.savez_compressed(
        os.path.join(out_dir, "updown_B2.npz"), names=np.array([n for n, _ in model.named_parameters()]),
        loss=np.array([float(loss.detach())]), out_slice=out.detach().reshape(-1)[::37].numpy(),
        grad_slice=gflat[::4099].numpy(),
        grad_norms=np.array([float(gflat[o:o + n].double().norm()) for o, n in offsets(shapes)]),
        param_slice=torch.cat([p.detach().reshape(-1) for p in model.parameters()])[::4099].numpy())
    print("updown_B2.npz written, loss", float(loss.detach()), "params", int(gflat.numel()))


def offsets(shapes):
    o = 0
    for s in shapes:
        n = int(np.prod(s))
        yield o, n
        o += n


def pad4(v):
    a = np.zeros(4, dtype=np.float32)
    a[:min(4, v.numel())] = v[:4].numpy()
    return a


if __name__ == "__main__":
    later = {"class_cond": gen_class_cond, "updown": gen_updown, "scale_shift": gen_scale_shift}
    if len(sys.argv) > 1 and sys.argv[1] in later:   # (added later: leave the other fixtures untouched)
        torch.set_num_threads(os.cpu_count())
        later[sys.argv[1]](os.path.join(ROOT, "tests", "golden"))
    else:
        main()
        for fn in later.values():
            fn(os.path.join(ROOT, "tests", "golden"))
