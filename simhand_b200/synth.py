"""Synthetic inputs for the weighted NT-Xent hot path.

All generators run on the CPU with an explicit seed so that the CPU oracle and
the CUDA path see identical bits.  Shapes follow the batch contract of the
reference data side (`src/data_loader/data_set.py:646-691`): joints are
`[N, 21, 3]` fp32 in pixel units of the 128x128 crop and the loss callers pass
the non-contiguous `[:, :, :2]` slice (`src/models/unsupervised/simhand_w_model.py:103-104`).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

NUM_JOINTS = 21
CROP = 128.0


def make_embeddings(n: int, d: int = 128, seed: int = 5, noise: float = 0.3):
    """Two views of L2-normalised projections `[n, d]` (already normalised, as the
    caller hands them to the loss: `simhand_w_model.py:91-94`)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, d, generator=g)
    raw1 = base + noise * torch.randn(n, d, generator=g)
    raw2 = base + noise * torch.randn(n, d, generator=g)
    return F.normalize(raw1, dim=1), F.normalize(raw2, dim=1), raw1, raw2


def make_joints_hand(n: int, seed: int = 5):
    """Crop-centred hands after the 128x128 resize (`sample_augmenter.py:425-476,196-222`):
    view 2 is a "similar hand" (view 1 + N(0, 4 px))."""
    g = torch.Generator().manual_seed(seed + 1000)
    off = torch.randn(n, NUM_JOINTS, 2, generator=g)
    radius = off.norm(dim=-1).amax(dim=1, keepdim=True).unsqueeze(-1)
    m = 0.9 + 0.6 * torch.rand(n, 1, 1, generator=g)
    off = off / radius * (CROP / 2) / m
    centre = CROP / 2 + (torch.rand(n, 1, 2, generator=g) * 16 - 8)
    xy1 = centre + off
    xy2 = xy1 + 4.0 * torch.randn(n, NUM_JOINTS, 2, generator=g)
    one = torch.ones(n, NUM_JOINTS, 1)
    return torch.cat([xy1, one], -1).contiguous(), torch.cat([xy2, one], -1).contiguous()


def make_joints_uniform(n: int, seed: int = 5):
    """Adversarial set: joints uniform in the crop, D spans 0..~100 px."""
    g = torch.Generator().manual_seed(seed + 2000)
    j1 = torch.rand(n, NUM_JOINTS, 3, generator=g) * CROP
    j2 = torch.rand(n, NUM_JOINTS, 3, generator=g) * CROP
    return j1.contiguous(), j2.contiguous()


def make_joints_peclr(n: int, seed: int = 5):
    """peclr_w variant (BASELINE config 4): view 2 is view 1 rotated by an integer angle
    in [-45, 45] degrees about the crop centre and shifted by an integer jitter in [0, 15] px
    (`training_config.json:36-57`, `sample_augmenter.py:224-252,462-474`)."""
    j1, _ = make_joints_hand(n, seed)
    g = torch.Generator().manual_seed(seed + 3000)
    ang = torch.randint(-45, 46, (n, 1), generator=g).float() * (math.pi / 180.0)
    jit = torch.randint(0, 16, (n, 1, 2), generator=g).float()
    c, s = torch.cos(ang), torch.sin(ang)
    rel = j1[:, :, :2] - CROP / 2
    x = rel[..., 0] * c - rel[..., 1] * s
    y = rel[..., 0] * s + rel[..., 1] * c
    xy2 = torch.stack([x, y], -1) + CROP / 2 + jit
    j2 = torch.cat([xy2, torch.ones(n, NUM_JOINTS, 1)], -1).contiguous()
    return j1, j2


JOINT_SETS = {
    "hand": make_joints_hand,
    "uniform": make_joints_uniform,
    "peclr": make_joints_peclr,
}


def make_batch(n: int, d: int = 128, seed: int = 5, joints: str = "hand"):
    """Returns (z1, z2, joints1[n,21,3], joints2[n,21,3]); pass `joints[:, :, :2]` to the op."""
    z1, z2, _, _ = make_embeddings(n, d, seed)
    j1, j2 = JOINT_SETS[joints](n, seed)
    return z1, z2, j1, j2
