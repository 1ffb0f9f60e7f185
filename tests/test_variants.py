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


def test_golden_files_present():
    assert len(NAMES) >= 6


@pytest.mark.parametrize("name", NAMES)
def test_restatement_matches_reference(name):
    g, cfg = _load(name)
    a, b = torch.from_numpy(g["joints1"])[:, :, :2], torch.from_numpy(g["joints2"])[:, :, :2]
    pos_w, neg_w = R.port_get_weights(a, b, **cfg)
    assert np.abs(pos_w.numpy() - g["pos_w"]).max() <= 1e-6
    assert np.abs(neg_w.numpy() - g["neg_w"]).max() <= 1e-6
    loss = R.port_loss(torch.from_numpy(g["z1"]).double(), torch.from_numpy(g["z2"]).double(), pos_w.double(),
                       neg_w.double(), 0.5)
    assert abs(float(loss) - float(g["loss_f64"])) <= 1e-6 * abs(float(g["loss_f64"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_materialised_variant_weights(name):
    from simhand_b200 import ops
    g, cfg = _load(name)
    dev = torch.device("cuda:0")
    a, b = torch.from_numpy(g["joints1"]).to(dev)[:, :, :2], torch.from_numpy(g["joints2"]).to(dev)[:, :, :2]
    if cfg["weight_type"] == "linear":
        hp, hn = ops.get_weights_linear(a, b, cfg["diff_type"])
    else:
        hp, hn = ops.get_weights_nonlinear(a, b, cfg["lambda_pos"], cfg["lambda_neg"], cfg["diff_type"])
    assert np.abs(hp.materialize().cpu().numpy() - g["pos_w"]).max() <= W_ATOL
    assert np.abs(hn.materialize().cpu().numpy() - g["neg_w"]).max() <= W_ATOL


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["fp32", "auto"])
@pytest.mark.parametrize("name", NAMES)
def test_fused_variant_step(name, engine):
    from simhand_b200 import ops
    g, cfg = _load(name)
    dev = torch.device("cuda:0")
    a, b = torch.from_numpy(g["joints1"]).to(dev)[:, :, :2], torch.from_numpy(g["joints2"]).to(dev)[:, :, :2]
    z1 = torch.from_numpy(g["z1"]).to(dev).requires_grad_(True)
    z2 = torch.from_numpy(g["z2"]).to(dev).requires_grad_(True)
    if cfg["weight_type"] == "linear":
        hp, hn = ops.get_weights_linear(a, b, cfg["diff_type"])
    else:
        hp, hn = ops.get_weights_nonlinear(a, b, cfg["lambda_pos"], cfg["lambda_neg"], cfg["diff_type"])
    loss = ops.vanila_weights_contrastive_loss(z1, z2, hp, hn, 0.5, engine=engine)
    loss.backward()
    ref = float(g["loss_f64"])
    assert abs(float(loss) - ref) <= LOSS_RTOL * abs(ref), (float(loss), ref)
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
