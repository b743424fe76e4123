"""Data-parallel equivalence of the CUDA path, on >= 2 GPUs of one box: R ranks (torchrun, NCCL
over NVLink), each training on its contiguous shard of a global batch with the bucketed gradient all-reduce inside the
captured step graph, must (1) hold bit-identical parameters on every rank after 3 AdamW steps and (2) agree with ONE
GPU training on the whole global batch up to bf16 / atomic summation order: the maximum parameter difference is bounded
by the distance the weights travelled (a near-zero gradient may flip sign), the mean difference by 5e-5.
Skipped on a single-GPU box (tests/test_dp_gloo.py covers the host logic on CPU)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2])
def test_nccl_data_parallel_matches_single_gpu(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DPCHECK ")]
    assert line, r.stdout[-3000:]
    res = json.loads(line[-1][8:])
    assert res["world"] == world and res["identical"] is True, res
    assert res["max_diff"] <= 2.5 * res["moved"] and res["mean_diff"] <= 5e-5, res
