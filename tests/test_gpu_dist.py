"""Sharded path on real GPUs (needs >= 2 devices; skipped on a single-GPU box): the result of
`weighted_ntxent(..., group=WORLD)` equals the single-GPU result on the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,transport", [(256, "auto"), (1024, "fused"), (1024, "peer"), (1024, "nccl")])
def test_sharded_equals_single_gpu(n, transport):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs (the single-GPU box runs tests/test_gpu_shard_emulation.py instead)")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tools", "dist_check.py"), str(n),
           transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == world
