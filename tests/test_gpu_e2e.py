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


def _check_dump_against_oracle(prefix, world):
    """The loss and d loss / d projections of the step's first iteration against the CPU oracle on the same projections
    and joints (gathered over the ranks): the training step computes the reference's global-batch loss."""
    import numpy as np
    from oracle import restate as R
    parts = [torch.load(f"{prefix}.{r}") for r in range(world)]
    b = parts[0]["p"].shape[0] // 2
    z1 = torch.cat([d["p"][:b] for d in parts])
    z2 = torch.cat([d["p"][b:] for d in parts])
    j1 = torch.cat([d["joints1"] for d in parts])
    j2 = torch.cat([d["joints2"] for d in parts])
    ref = R.c_step(z1, z2, j1, j2)
    for d in parts:
        assert abs(d["loss"] - ref["loss"]) <= 1e-5 * abs(ref["loss"]), (d["loss"], ref["loss"])
    g1 = torch.cat([d["dp"][:b] for d in parts]).numpy() / world          # the op was called with grad_scale = world
    g2 = torch.cat([d["dp"][b:] for d in parts]).numpy() / world
    cos, mx = R.grad_metrics(np.concatenate([g1, g2]), np.concatenate([ref["dz1"], ref["dz2"]]))
    assert cos >= 0.9999 and mx <= 1e-3, (cos, mx)


@pytest.mark.parametrize("fused_head", [False, True])
def test_single_gpu_training_step(fused_head, tmp_path):
    prefix = str(tmp_path / "dump")
    res = _run([sys.executable, os.path.join(ROOT, "examples", "e2e_step.py"), "--batch", "256", "--steps", "4",
                "--warmup", "2", "--image", "64", "--dump", prefix] + (["--fused-head"] if fused_head else []))
    assert res["finite"] and res["loss_decreasing"], res
    _check_dump_against_oracle(prefix, 1)


def test_sharded_training_step(tmp_path):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    prefix = str(tmp_path / "dump")
    res = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                "127.0.0.1", "--master-port", "29633", os.path.join(ROOT, "examples", "e2e_step.py"), "--batch", "512",
                "--steps", "4", "--warmup", "2", "--image", "64", "--fused-head", "--dump", prefix])
    assert res["finite"] and res["loss_decreasing"] and res["world"] == 2, res
    _check_dump_against_oracle(prefix, 2)
