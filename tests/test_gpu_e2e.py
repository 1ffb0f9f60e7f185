"""End-to-end training step: ResNet-50 + projection head under bf16 autocast -> fused projection-space
transform -> fused global-batch loss -> backward -> optimizer step.  Checks that the step runs through the public API,
that gradients reach the backbone and that repeated steps on one batch lower the loss."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


def test_single_gpu_training_step():
    res = _run([sys.executable, os.path.join(ROOT, "examples", "e2e_step.py"), "--batch", "256", "--steps", "4",
                "--warmup", "2", "--image", "64"])
    assert res["finite"] and res["loss_decreasing"], res


def test_sharded_training_step():
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    res = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                "127.0.0.1", "--master-port", "29633", os.path.join(ROOT, "examples", "e2e_step.py"), "--batch", "512",
                "--steps", "4", "--warmup", "2", "--image", "64"])
    assert res["finite"] and res["loss_decreasing"] and res["world"] == 2, res
