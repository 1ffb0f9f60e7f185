"""TEST INFRASTRUCTURE ONLY: writes tests/golden/*.npz by running the reference's own
functions (AST-loaded from /root/reference, see oracle/ref_loader.py) on seeded synthetic
inputs.  Run in the build container:  python -m oracle.gen_golden

Each file holds the inputs (so the GPU box does not depend on RNG reproducibility) and the
reference outputs: pos_w, neg_w (src/models/utils.py:218-261), loss and autograd gradients
(src/models/utils.py:391-427) in fp32 exactly as the reference computes them on torch-CPU,
plus the same loss/gradients recomputed by the reference code in fp64 on the fp32 weights
(the "true" value the tolerances of BASELINE.json are measured against).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.ref_loader import load_reference_functions
from simhand_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, per-view batch N, d, joint set, seed)
CASES = [
    ("n64_hand", 64, 128, "hand", 5),          # one 128x128 tile
    ("n96_uniform", 96, 128, "uniform", 6),    # ragged: 2N = 192
    ("n3_uniform", 3, 128, "uniform", 7),      # tiny
    ("n200_peclr", 200, 128, "peclr", 8),      # 2N = 400, not a multiple of the tile
    ("n256_hand", 256, 128, "hand", 5),        # BASELINE.json configs[0]
    ("n256_uniform", 256, 128, "uniform", 5),
    ("n130_hand_d64", 130, 64, "hand", 9),     # other projection width
]


def run_case(ns, n, d, jset, seed):
    z1, z2, j1, j2 = synth.make_batch(n, d, seed, jset)
    a, b = j1[:, :, :2], j2[:, :, :2]           # the strided views the callers pass
    pos_w, neg_w = ns["get_weights_linear"](a, b, "mpjpe")
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        x1 = z1.to(dt).clone().detach().requires_grad_(True)
        x2 = z2.to(dt).clone().detach().requires_grad_(True)
        loss = ns["vanila_weights_contrastive_loss"](x1, x2, pos_w.to(dt), neg_w.to(dt))
        loss.backward()
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"dz1_{tag}"] = x1.grad.numpy()
        out[f"dz2_{tag}"] = x2.grad.numpy()
    # the other weightings of the same loss (pos_neg == "pos" / "neg", and plain NT-Xent), fp64
    for tag, call in (("pos", lambda a, c: ns["vanila_pos_weights_contrastive_loss"](a, c, pos_w.double())),
                      ("neg", lambda a, c: ns["vanila_neg_weights_contrastive_loss"](a, c, neg_w.double())),
                      ("plain", lambda a, c: ns["vanila_contrastive_loss"](a, c))):
        x1 = z1.double().clone().detach().requires_grad_(True)
        x2 = z2.double().clone().detach().requires_grad_(True)
        loss = call(x1, x2)
        loss.backward()
        out[f"loss_{tag}_f64"] = loss.detach().numpy()
        out[f"dz1_{tag}_f64"] = x1.grad.numpy()
    out.update(z1=z1.numpy(), z2=z2.numpy(), joints1=j1.numpy(), joints2=j2.numpy(),
               pos_w=pos_w.numpy(), neg_w=neg_w.numpy(), temperature=np.float64(0.5))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference_functions()
    torch.set_num_threads(8)
    for name, n, d, jset, seed in CASES:
        res = run_case(ns, n, d, jset, seed)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: loss={float(res['loss_f32']):.7f} (fp64 {float(res['loss_f64']):.9f}) "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
