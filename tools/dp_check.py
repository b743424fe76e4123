"""Data-parallel equivalence: R ranks, each training on its contiguous shard of a global batch
with NCCL gradient all-reduce, must produce the same parameters as one rank training on the whole batch.
    torchrun --nproc-per-node 2 tools/dp_check.py        (rank 0 prints the comparison)"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge
import unet_oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ub = ge.load_package()
cfg = O.UNetConfig()
flat = O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy()
Bg = 4 * world
x0, t, noise = O.synthetic_batch(cfg, Bg)
lo, hi = ub.shard_batch(Bg, rank, world)
tr = ub.Trainer(B=hi - lo, device=int(os.environ["LOCAL_RANK"]))
tr.set_params(flat)
idt = torch.zeros(ub.UB_NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(ub.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
tr.attach_dp(rank, world, bytes(idt.cpu().numpy().tobytes()))
losses = [tr.train_step(x0[lo:hi].numpy(), t[lo:hi].numpy(), noise[lo:hi].numpy(), lr=1e-4) for _ in range(3)]
p_dp = tr.get_params()
tr.close()
allp = [torch.zeros(p_dp.size, device="cuda") for _ in range(world)]
dist.all_gather(allp, torch.from_numpy(p_dp).cuda())
if rank == 0:
    ref = ub.Trainer(B=Bg, device=0)
    ref.set_params(flat)
    lref = [ref.train_step(x0.numpy(), t.numpy(), noise.numpy(), lr=1e-4) for _ in range(3)]
    p_ref = ref.get_params()
    ref.close()
    moved = np.abs(p_ref - flat).max()
    print(f"world {world}: replicas identical across ranks: "
          f"{all(bool(torch.equal(allp[0], a)) for a in allp[1:])}")
    print(f"  max |p_dp - p_single| = {np.abs(p_dp - p_ref).max():.3e}  (3 AdamW steps moved weights by up to {moved:.3e}); "
          f"mean |diff| = {np.abs(p_dp - p_ref).mean():.3e}")
    print(f"  shard-0 losses {['%.5f' % l for l in losses]}  full-batch losses {['%.5f' % l for l in lref]}")
    import json
    print("DPCHECK " + json.dumps({"world": world, "identical": all(bool(torch.equal(allp[0], a)) for a in allp[1:]),
                                   "max_diff": float(np.abs(p_dp - p_ref).max()), "moved": float(moved),
                                   "mean_diff": float(np.abs(p_dp - p_ref).mean())}))
dist.barrier()
dist.destroy_process_group()
