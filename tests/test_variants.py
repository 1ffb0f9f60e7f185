"""The other weightings of the reference: weight_type non_linear and diff_type w_abs / w_o_abs.
CPU: the oracle's restatement against golden vectors from the reference's own get_weights_linear / get_weights_nonlinear.
GPU: materialised weights and the fused loss + gradients against the same goldens."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_variants")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLD, "*.npz")))
W_ATOL = 2e-6          # the variants are fp32 reductions in library order, not bit-pinned like the MPJPE path
LOSS_RTOL = 1e-5
GRAD_COS, GRAD_MAXABS = 0.9999, 1e-3


def _load(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    cfg = dict(weight_type=str(g["weight_type"]), diff_type=str(g["diff_type"]), lambda_pos=float(g["lambda_pos"]),
               lambda_neg=float(g["lambda_neg"]))
    return g, cfg


def _joints(g, cfg, dev=None):
    """What the caller hands to get_weights_*: [N, 21, 2] slices, or [N, 14] PCA coordinates for the *_with_pca cases."""
    a, b = torch.from_numpy(g["joints1"]), torch.from_numpy(g["joints2"])
    if dev is not None:
        a, b = a.to(dev), b.to(dev)
    if cfg["diff_type"].startswith("pca:"):
        return a, b
    return a[:, :, :2], b[:, :, :2]


def _handles(ops, a, b, cfg):
    diff = cfg["diff_type"]
    if diff.startswith("pca:"):
        if cfg["weight_type"] == "linear":
            return ops.get_weights_linear_with_pca(a, b, diff[4:])
        return ops.get_weights_nonlinear_with_pca(a, b, cfg["lambda_pos"], cfg["lambda_neg"], diff[4:])
    if cfg["weight_type"] == "linear":
        return ops.get_weights_linear(a, b, diff)
    return ops.get_weights_nonlinear(a, b, cfg["lambda_pos"], cfg["lambda_neg"], diff)


def _check_neg_w(got, g, atol):
    if "neg_w" in g:
        assert np.abs(got - g["neg_w"]).max() <= atol
    else:
        assert np.abs(got[:8] - g["neg_w_rows"]).max() <= atol      # the tensor-core-size cases keep 8 rows of the matrix


def test_golden_files_present():
    assert len(NAMES) >= 12
    assert sum(n.startswith("big_") for n in NAMES) >= 4 and sum("pca" in n for n in NAMES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_restatement_matches_reference(name):
    g, cfg = _load(name)
    a, b = _joints(g, cfg)
    port_cfg = dict(cfg, diff_type="pca" if cfg["diff_type"].startswith("pca:") else cfg["diff_type"])
    pos_w, neg_w = R.port_get_weights(a, b, **port_cfg)
    assert np.abs(pos_w.numpy() - g["pos_w"]).max() <= 1e-6
    _check_neg_w(neg_w.numpy(), g, 1e-6)
    loss = R.port_loss(torch.from_numpy(g["z1"]).double(), torch.from_numpy(g["z2"]).double(), pos_w.double(),
                       neg_w.double(), 0.5)
    assert abs(float(loss) - float(g["loss_f64"])) <= 1e-6 * abs(float(g["loss_f64"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_materialised_variant_weights(name):
    from simhand_b200 import ops
    g, cfg = _load(name)
    dev = torch.device("cuda:0")
    a, b = _joints(g, cfg, dev)
    hp, hn = _handles(ops, a, b, cfg)
    atol = 4e-6 if "pca" in name else W_ATOL            # PCA coordinates are O(100): fp32 sum of 14 squares
    # the sigmoid amplifies the fp32 rounding of a distance of O(300) (ulp 3e-5) by lambda / 4
    atol_pos = atol + (1e-5 * abs(cfg["lambda_pos"]) if "pca" in name else 0.0)
    assert np.abs(hp.materialize().cpu().numpy() - g["pos_w"]).max() <= atol_pos
    _check_neg_w(hn.materialize().cpu().numpy(), g, atol)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["fp32", "auto", "fp16", "tf32", "bf16"])
@pytest.mark.parametrize("name", NAMES)
def test_fused_variant_step(name, engine):
    """Every weighting under every engine against the reference's own outputs.  The big_* cases have 2N > 256, so "auto"
    resolves to the tcgen05 sweeps there; fp16 / tf32 / bf16 force them on every case."""
    from simhand_b200 import ops
    g, cfg = _load(name)
    dev = torch.device("cuda:0")
    a, b = _joints(g, cfg, dev)
    z1 = torch.from_numpy(g["z1"]).to(dev).requires_grad_(True)
    z2 = torch.from_numpy(g["z2"]).to(dev).requires_grad_(True)
    if name.startswith("big_"):
        assert ops.resolve_engine("auto", z1.shape[0]) == "fp16"
    hp, hn = _handles(ops, a, b, cfg)
    loss = ops.vanila_weights_contrastive_loss(z1, z2, hp, hn, 0.5, engine=engine)
    loss.backward()
    ref = float(g["loss_f64"])
    assert abs(float(loss) - ref) <= (1e-3 if engine == "bf16" else LOSS_RTOL) * abs(ref), (float(loss), ref)
    for got, key in ((z1.grad, "dz1_f64"), (z2.grad, "dz2_f64")):
        cos, mx = R.grad_metrics(got.cpu().numpy(), g[key])
        assert cos >= GRAD_COS and mx <= GRAD_MAXABS, (key, cos, mx)


@pytest.mark.gpu
def test_variant_full_size_consistency():
    """2N = 16384, non_linear / w_abs: the fused tensor-core step against the fp32 engine (same tiles, fp32 logits) and
    the mean-distance reduction against the materialised matrix of a 2N = 2048 problem."""
    from simhand_b200 import ops, synth
    dev = torch.device("cuda:0")
    wt = ops.make_weighting("non_linear", "w_abs", 2.5, 0.05)
    z1, z2, j1, j2 = synth.make_batch(8192, 128, 31, "hand")
    z1, z2, a, b = z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    l_tc, g1, _ = ops.run_step(z1, z2, a, b, 0.5, "auto", True, weighting=wt)
    l_32, h1, _ = ops.run_step(z1, z2, a, b, 0.5, "fp32", True, weighting=wt)
    assert abs(float(l_tc) - float(l_32)) <= LOSS_RTOL * abs(float(l_32))
    cos, mx = R.grad_metrics(g1.cpu().numpy(), h1.cpu().numpy())
    assert cos >= GRAD_COS and mx <= GRAD_MAXABS
    hp, hn = ops.get_weights_nonlinear(a[:1024], b[:1024], 2.5, 0.05, "w_abs")
    pw, nw = R.port_get_weights(a[:1024].cpu(), b[:1024].cpu(), "non_linear", "w_abs", 2.5, 0.05)
    assert (hn.materialize().cpu() - nw).abs().max() <= W_ATOL
    assert (hp.materialize().cpu() - pw).abs().max() <= W_ATOL
