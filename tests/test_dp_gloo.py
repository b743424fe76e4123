"""Data-parallel host logic on CPU: world_size 2, gloo.  Each rank runs the oracle on its contiguous batch shard
(ub.shard_batch), gradients are summed with all_reduce and scaled by 1/world exactly as the CUDA path does
(NCCL sum + grad_scale in AdamW); the result must equal the single-process full-batch gradient / update."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny_cfg(O):
    return O.UNetConfig(model_channels=32, channel_mult=(1, 2), attn_start_level=1, num_res_blocks=1, H=16, W=16)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unet_oracle as O
    import __graft_entry__ as ge
    ub = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = _tiny_cfg(O)
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0))
    x0, t, noise = O.synthetic_batch(cfg, 8)
    lo, hi = ub.shard_batch(8, rank, world)
    loss, _, g = O.train_step_grads(cfg, flat, x0[lo:hi], t[lo:hi], noise[lo:hi])
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    g = g / world
    lt = torch.tensor([float(loss)])
    dist.all_reduce(lt)
    new_flat, _, _ = O.adamw_step(flat, g, torch.zeros_like(flat), torch.zeros_like(flat), 1)
    if rank == 0:
        np.savez(os.path.join(out_dir, "dp.npz"), g=g.numpy(), loss=float(lt) / world, p=new_flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dp_equals_single_process(tmp_path, oracle):
    O = oracle
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(str(tmp_path / "dp.npz"))
    cfg = _tiny_cfg(O)
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0))
    x0, t, noise = O.synthetic_batch(cfg, 8)
    loss, _, g = O.train_step_grads(cfg, flat, x0, t, noise)
    new_flat, _, _ = O.adamw_step(flat, g, torch.zeros_like(flat), torch.zeros_like(flat), 1)
    assert abs(float(loss) - float(r["loss"])) < 1e-6
    gn = g.numpy()
    assert np.abs(gn - r["g"]).max() <= 1e-5 * np.abs(gn).max() + 1e-9
    # the first AdamW step moves every weight by ~lr*sign(g): compare where the gradient is not at rounding level
    mask = np.abs(gn) > 1e-4 * np.abs(gn).max()
    np.testing.assert_allclose(new_flat.numpy()[mask], r["p"][mask], rtol=0, atol=2e-6)
