"""TEST INFRASTRUCTURE ONLY: writes tests/golden_transform/*.npz by running the reference's own
translate_encodings / rotate_encoding / get_rotation_2D_matrix (AST-loaded from /root/reference,
src/models/utils.py:606-684) composed exactly as HandCLR_W.get_transformed_projections does
(src/models/unsupervised/simhand_w_model.py:55-94): normalise, translate by -jitter, rotate by -angles,
normalise.  Run in the build container:  python -m oracle.gen_golden_transform

Each file holds the inputs, the reference's output and autograd gradient (of sum(out * cot) with respect to the raw
projections, for a stored cotangent) in fp32 -- the reference builds its rotation matrices as fp32 zeros
(utils.py:625), so its code only runs in fp32 -- and the same quantities from the oracle's restatement
(oracle/restate.py: port_transform) in fp64, which is what the kernel tolerances are measured against; the script
prints how far the restatement is from the reference.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle.ref_loader import load_reference_functions
from oracle.restate import port_transform

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_transform")

# (name, rows, d, crop, rotate, seed)
CASES = [
    ("r128_d128_crop_rot", 128, 128, True, True, 11),      # handclr_w / peclr_w default augmentations
    ("r50_d128_rot", 50, 128, False, True, 12),
    ("r50_d128_crop", 50, 128, True, False, 13),
    ("r33_d64_crop_rot", 33, 64, True, True, 14),
    ("r16_d6_crop_rot", 16, 6, True, True, 15),            # d not a multiple of 4
    ("r20_d128_plain", 20, 128, False, False, 16),         # two normalisations only
    ("r24_d64_crop", 24, 64, True, False, 17),             # crop only with d < 128: padding lanes must not move
    ("r12_d6_crop", 12, 6, True, False, 18),               # crop only, d % 4 == 2: half-used lane
    ("r12_d30_rot", 12, 30, False, True, 19),              # rotate only, d % 4 == 2
]


def reference_transform(ns, proj, jitter_x, jitter_y, angles, crop, rotate):
    """simhand_w_model.py:55-94 with the batch-dict plumbing removed; `proj` is [2B, d]."""
    rows = proj.shape[0]
    p = F.normalize(proj).view(rows, -1, 2)                                      # :56-60 (both halves, row-wise)
    if crop:
        p = ns["translate_encodings"](p, -jitter_x, -jitter_y, None)             # :79
    if rotate:
        p = ns["rotate_encoding"](p, -angles, None)                              # :85
    return F.normalize(p.reshape(rows, -1))                                      # :87-93


def run_case(ns, rows, d, crop, rotate, seed):
    gen = torch.Generator().manual_seed(seed)
    proj = torch.randn(rows, d, generator=gen) * 3.0
    jitter_x = torch.randint(0, 16, (rows,), generator=gen).float() / 128.0       # jitter / image size (:62-77)
    jitter_y = torch.randint(0, 16, (rows,), generator=gen).float() / 128.0
    angles = torch.randint(-45, 46, (rows,), generator=gen).float()               # training_config.json:36-57
    cot = torch.randn(rows, d, generator=gen)
    out = {}
    x = proj.clone().requires_grad_(True)
    y = reference_transform(ns, x, jitter_x, jitter_y, angles, crop, rotate)
    (y * cot).sum().backward()
    out["out_f32"], out["dx_f32"] = y.detach().numpy(), x.grad.numpy()
    x = proj.double().clone().requires_grad_(True)
    y = port_transform(x, -jitter_x.double() if crop else None, -jitter_y.double() if crop else None,
                       -angles.double() if rotate else None)
    (y * cot.double()).sum().backward()
    out["out_f64"], out["dx_f64"] = y.detach().numpy(), x.grad.numpy()
    out.update(proj=proj.numpy(), jitter_x=jitter_x.numpy(), jitter_y=jitter_y.numpy(), angles=angles.numpy(),
               cot=cot.numpy(), crop=np.bool_(crop), rotate=np.bool_(rotate))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference_functions(("translate_encodings", "rotate_encoding", "get_rotation_2D_matrix"))
    for name, rows, d, crop, rotate, seed in CASES:
        res = run_case(ns, rows, d, crop, rotate, seed)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **res)
        err = np.abs(res["out_f32"] - res["out_f64"]).max()
        gerr = np.abs(res["dx_f32"] - res["dx_f64"]).max() / np.abs(res["dx_f64"]).max()
        print(f"{name}: reference fp32 vs restatement fp64: out {err:.2e}, grad {gerr:.2e} (rel. to max) "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
