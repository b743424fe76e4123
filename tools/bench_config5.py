"""BASELINE.json configs[4]: 128x128 U-Net, channel_mult (1,1,2,3,4), attention at 16x16 and 8x8 (21 082 755 parameters),
batch 64 per GPU, full training step -- on one GPU, or data parallel under torchrun (weak scaling, NCCL all-reduce inside
the captured step graph; timed like bench.py: barrier + synchronize on both sides, CUDA events, max over ranks).
    python tools/bench_config5.py [B] [steps]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_config5.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as ge
import unet_oracle as O
ub = ge.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
cfg = O.UNetConfig(channel_mult=(1, 1, 2, 3, 4), attn_start_level=3, H=128, W=128)
tr = ub.Trainer(B=B, H=128, W=128, channel_mult=(1, 1, 2, 3, 4), att_start_level=3, device=local_rank, seed=1234 + rank)
tr.set_params(O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy())
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    idt = torch.zeros(ub.UB_NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(ub.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    tr.attach_dp(rank, world, bytes(idt.cpu().numpy().tobytes()))
x = (torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(7 + rank)) * 2 - 1).cuda()
for _ in range(5):
    tr.train_step_device(x.data_ptr())
tr.sync()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


stream = torch.cuda.ExternalStream(tr.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
barrier()
e0.record(stream)
for _ in range(steps):
    tr.train_step_device(x.data_ptr())
e1.record(stream)
tr.sync()
barrier()
t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
loss = tr.last_loss()  # (before the profile: its replays run every op several times, the loss of those steps is meaningless)
if rank == 0:
    prof = tr.profile(reps=2)
    print(json.dumps({"workload": "128x128 U-Net mult (1,1,2,3,4), attention at 16x16 and 8x8, full train step",
                      "n_gpus": world, "batch_per_gpu": B, "global_batch": B * world, "ms_per_step": ms,
                      "images_per_s": world * B / ms * 1e3, "loss": loss,
                      "model_tflops_per_gpu": B / ms * 1e3 * 88.54e9 / 1e12,
                      "classes_ms": {k: round(v["ms"], 3) for k, v in prof.items() if isinstance(v, dict)},
                      "mem_GB": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)}))
tr.close()
if world > 1:
    dist.destroy_process_group()
