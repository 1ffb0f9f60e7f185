"""Projection-space transform: the oracle's restatement against the reference's golden vectors
(CPU), and the fused kernel against both (GPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_transform")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def _load(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    t = {k: torch.from_numpy(v) for k, v in g.items() if k not in ("crop", "rotate")}
    crop, rot = bool(g["crop"]), bool(g["rotate"])
    # what the reference hands to translate_encodings / rotate_encoding (simhand_w_model.py:79, :85)
    args = (-t["jitter_x"] if crop else None, -t["jitter_y"] if crop else None, -t["angles"] if rot else None)
    return t, args


def test_golden_files_present():
    assert len(NAMES) >= 6


@pytest.mark.parametrize("name", NAMES)
def test_restatement_matches_reference(name):
    """fp32 restatement vs the reference's own fp32 output and autograd gradient."""
    t, (tx, ty, an) = _load(name)
    x = t["proj"].clone().requires_grad_(True)
    y = R.port_transform(x, tx, ty, an)
    (y * t["cot"]).sum().backward()
    assert (y.detach() - t["out_f32"]).abs().max() <= 5e-7
    assert (x.grad - t["dx_f32"]).abs().max() <= 2e-6 * t["dx_f32"].abs().max()
    assert (y.detach().double() - t["out_f64"]).abs().max() <= 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_fused_transform_matches_golden(name):
    from simhand_b200 import ops
    t, (tx, ty, an) = _load(name)
    dev = torch.device("cuda:0")
    x = t["proj"].to(dev).requires_grad_(True)
    to = lambda v: None if v is None else v.to(dev)                      # noqa: E731
    y = ops.get_transformed_projections(x, to(tx), to(ty), to(an))
    (y * t["cot"].to(dev)).sum().backward()
    assert (y.detach().cpu().double() - t["out_f64"]).abs().max() <= 1e-6          # fp32 kernel vs fp64 truth
    assert (y.detach().cpu() - t["out_f32"]).abs().max() <= 1e-6                   # and vs the reference's fp32 run
    gerr = (x.grad.cpu().double() - t["dx_f64"]).abs().max() / t["dx_f64"].abs().max()
    assert gerr <= 5e-6, gerr


@pytest.mark.gpu
def test_fused_transform_full_size_properties():
    """2N = 16384 rows: unit-norm rows, zero translation + zero angle == double normalisation, and the gradient is
    orthogonal to the raw projection (scale invariance of the two normalisations)."""
    from simhand_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    x = (torch.randn(16384, 128, generator=gen) * 2).to(dev).requires_grad_(True)
    tx = (torch.randint(0, 16, (16384,), generator=gen).float() / 128).to(dev)
    an = torch.randint(-45, 46, (16384,), generator=gen).float().to(dev)
    y = ops.get_transformed_projections(x, -tx, -tx, -an)
    assert (y.norm(dim=1) - 1).abs().max() <= 1e-6
    y.sum().backward()
    assert ((x.grad * x.detach()).sum(1).abs() / (x.grad.norm(dim=1) * x.detach().norm(dim=1))).max() <= 1e-4
    zero = torch.zeros(16384, device=dev)
    y0 = ops.get_transformed_projections(x.detach(), zero, zero, zero)
    ref = torch.nn.functional.normalize(torch.nn.functional.normalize(x.detach()))
    assert (y0 - ref).abs().max() <= 1e-6


@pytest.mark.gpu
def test_fused_transform_rejects_bad_shapes():
    from simhand_b200 import ops
    dev = torch.device("cuda:0")
    with pytest.raises(ValueError):
        ops.get_transformed_projections(torch.zeros(4, 7, device=dev))
    with pytest.raises(ValueError):
        ops.get_transformed_projections(torch.zeros(4, 8, device=dev), torch.zeros(4, device=dev), None)
    with pytest.raises(RuntimeError):
        ops.get_transformed_projections(torch.zeros(4, 8))
