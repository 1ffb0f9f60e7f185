"""TEST INFRASTRUCTURE ONLY: writes tests/golden_variants/*.npz by running the reference's own
get_weights_linear (diff_type w_abs / w_o_abs, src/models/utils.py:218-261) and get_weights_nonlinear
(all three diff types, utils.py:304-346), AST-loaded from /root/reference, followed by its
vanila_weights_contrastive_loss (utils.py:391-427) in fp64 on those fp32 weights.
Run in the build container:  python -m oracle.gen_golden_variants
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.ref_loader import load_reference_functions
from simhand_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_variants")

# (name, N, joint set, seed, weight_type, diff_type, lambda_pos, lambda_neg); lambdas from src/experiments/utils.py:63-67
CASES = [
    ("lin_wabs_n64", 64, "hand", 21, "linear", "w_abs", 0.0, 0.0),
    ("lin_woabs_n100", 100, "uniform", 22, "linear", "w_o_abs", 0.0, 0.0),
    ("nl_mpjpe_n64", 64, "hand", 23, "non_linear", "mpjpe", 5.0, 0.05),
    ("nl_mpjpe_n100", 100, "uniform", 24, "non_linear", "mpjpe", 1.0, 0.01),
    ("nl_wabs_n100", 100, "peclr", 25, "non_linear", "w_abs", 2.5, 0.05),
    ("nl_woabs_n64", 64, "hand", 26, "non_linear", "w_o_abs", 1.0, 0.005),
    # config.use_pca: apply_pca (utils.py:192-215, randomised torch.pca_lowrank under a fixed seed) then *_with_pca; the
    # PCA coordinates are stored (joints1 / joints2 = [N, 14]) because the basis is not reproducible across machines
    ("pca_lin_mpjpe_n96", 96, "hand", 27, "linear", "pca:mpjpe", 0.0, 0.0),
    ("pca_nl_wabs_n64", 64, "uniform", 28, "non_linear", "pca:w_abs", 2.5, 0.05),
    # tensor-core sizes (2N > 256: engine "auto" resolves to the tcgen05 sweeps); neg_w is not stored (checked through
    # the loss, the gradients and neg_w_rows = its first 8 rows) to keep the fixtures small
    ("big_lin_wabs_n320", 320, "hand", 31, "linear", "w_abs", 0.0, 0.0),
    ("big_nl_mpjpe_n288", 288, "peclr", 32, "non_linear", "mpjpe", 2.5, 0.05),
    ("big_nl_woabs_n264", 264, "uniform", 33, "non_linear", "w_o_abs", 1.0, 0.01),
    ("big_pca_nl_n272", 272, "hand", 34, "non_linear", "pca:w_o_abs", 2.5, 0.05),
]


def run_case(ns, n, jset, seed, wtype, diff, lam_p, lam_n):
    z1, z2, j1, j2 = synth.make_batch(n, 128, seed, jset)
    a, b = j1[:, :, :2], j2[:, :, :2]
    if diff.startswith("pca:"):
        torch.manual_seed(seed)
        a, b = ns["apply_pca"](a, target_dim=14), ns["apply_pca"](b, target_dim=14)              # simhand_w_model.py:109-111
        j1, j2 = a, b
        if wtype == "linear":
            pos_w, neg_w = ns["get_weights_linear_with_pca"](a, b, diff[4:])
        else:
            pos_w, neg_w = ns["get_weights_nonlinear_with_pca"](a, b, lam_p, lam_n, diff[4:])
    elif wtype == "linear":
        pos_w, neg_w = ns["get_weights_linear"](a, b, diff)
    else:
        pos_w, neg_w = ns["get_weights_nonlinear"](a, b, lam_p, lam_n, diff)
    x1 = z1.double().clone().requires_grad_(True)
    x2 = z2.double().clone().requires_grad_(True)
    loss = ns["vanila_weights_contrastive_loss"](x1, x2, pos_w.double(), neg_w.double())
    loss.backward()
    big = n > 128
    return dict(z1=z1.numpy(), z2=z2.numpy(), joints1=j1.numpy(), joints2=j2.numpy(), pos_w=pos_w.numpy(),
                **({"neg_w_rows": neg_w[:8].numpy()} if big else {"neg_w": neg_w.numpy()}), loss_f64=loss.detach().numpy(), dz1_f64=x1.grad.numpy(), dz2_f64=x2.grad.numpy(),
                weight_type=np.str_(wtype), diff_type=np.str_(diff), lambda_pos=np.float64(lam_p),
                lambda_neg=np.float64(lam_n), temperature=np.float64(0.5))


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference_functions(("get_weights_linear", "get_weights_nonlinear", "vanila_weights_contrastive_loss",
                                   "apply_pca", "get_weights_linear_with_pca", "get_weights_nonlinear_with_pca"))
    for name, n, jset, seed, wtype, diff, lam_p, lam_n in CASES:
        res = run_case(ns, n, jset, seed, wtype, diff, lam_p, lam_n)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: loss {float(res['loss_f64']):.9f} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
