"""GPU parity of the materialised-weights path (the reference's two-call API handed REAL tensors,
`src/models/utils.py:391-427`, :430-465, :468-501): `vanila_weights_contrastive_loss(z1, z2, pos_w, neg_w)` with
arbitrary fp32 weight tensors, checked against the oracle's op-for-op port run in fp64 on the CPU (autograd gradients,
so nothing is assumed about the symmetry of the weights)."""
import numpy as np
import pytest
import torch

from oracle import restate as R
from simhand_b200 import ops, synth

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_COS = 0.9999
GRAD_MAXABS = 1e-3


def _dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _ref(z1, z2, pos_w, neg_w, tau):
    a = z1.double().clone().requires_grad_(True)
    b = z2.double().clone().requires_grad_(True)
    loss = R.port_loss(a, b, pos_w.double(), neg_w.double(), tau)
    loss.backward()
    return float(loss), a.grad, b.grad


def _check(loss, g1, g2, ref, loss_rtol=LOSS_RTOL):
    rl, r1, r2 = ref
    assert abs(float(loss) - rl) <= loss_rtol * abs(rl), (float(loss), rl)
    g = torch.cat([g1, g2]).double().cpu()
    r = torch.cat([r1, r2])
    cos = float((g * r).sum() / (g.norm() * r.norm()))
    err = float((g - r).abs().max() / r.abs().max())
    assert cos >= GRAD_COS and err <= GRAD_MAXABS, (cos, err)


def _problem(n, d, seed, symmetric):
    gen = torch.Generator().manual_seed(seed)
    z1 = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
    z2 = torch.nn.functional.normalize(z1 + 0.3 * torch.randn(n, d, generator=gen), dim=-1)
    neg_w = torch.rand(2 * n, 2 * n, generator=gen) * 1.3 - 0.1          # outside [0, 1] on purpose
    if symmetric:
        neg_w = 0.5 * (neg_w + neg_w.t())
    pos_w = torch.rand(n, generator=gen)
    return z1, z2, pos_w, neg_w


@pytest.mark.parametrize("engine", ["fp32", "auto", "bf16"])
@pytest.mark.parametrize("n,d,symmetric", [(200, 128, False), (64, 128, True), (130, 64, False), (1024, 128, False)])
def test_dense_weights_match_reference(n, d, symmetric, engine):
    tau = 0.5
    z1, z2, pos_w, neg_w = _problem(n, d, 1234 + n, symmetric)
    ref = _ref(z1, z2, pos_w, neg_w, tau)
    dev = _dev()
    a = z1.to(dev).requires_grad_(True)
    b = z2.to(dev).requires_grad_(True)
    loss = ops.vanila_weights_contrastive_loss(a, b, pos_w.to(dev), neg_w.to(dev), tau, engine=engine)
    loss.backward()
    _check(loss, a.grad, b.grad, ref, 1e-3 if engine == "bf16" else LOSS_RTOL)


def test_dense_equals_fused_on_reference_weights(golden):
    """Materialising the weights (get_weights_linear handles -> tensors) and feeding them back gives the fused result."""
    if golden["z1"].shape[0] < 8:
        pytest.skip("fp16 logits over < 16 samples do not average to 1e-5")
    dev = _dev()
    z1, z2 = torch.from_numpy(golden["z1"]).to(dev), torch.from_numpy(golden["z2"]).to(dev)
    pos_w, neg_w = torch.from_numpy(golden["pos_w"]).to(dev), torch.from_numpy(golden["neg_w"]).to(dev)
    a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    loss = ops.vanila_weights_contrastive_loss(a, b, pos_w, neg_w, 0.5)
    loss.backward()
    ref = (float(golden["loss_f64"]), torch.from_numpy(golden["dz1_f64"]), torch.from_numpy(golden["dz2_f64"]))
    _check(loss, a.grad, b.grad, ref)


@pytest.mark.parametrize("which", ["pos", "neg"])
def test_single_sided_dense_variants(which):
    """vanila_pos_weights_contrastive_loss / vanila_neg_weights_contrastive_loss with a real tensor (utils.py:430, :468)."""
    n, d, tau = 96, 128, 0.5
    z1, z2, pos_w, neg_w = _problem(n, d, 77, False)
    dev = _dev()
    a, b = z1.to(dev).requires_grad_(True), z2.to(dev).requires_grad_(True)
    if which == "pos":
        ref = _ref(z1, z2, pos_w, torch.ones_like(neg_w), tau)
        loss = ops.vanila_pos_weights_contrastive_loss(a, b, pos_w.to(dev), tau)
    else:
        ref = _ref(z1, z2, torch.ones_like(pos_w), neg_w, tau)
        loss = ops.vanila_neg_weights_contrastive_loss(a, b, neg_w.to(dev), tau)
    loss.backward()
    _check(loss, a.grad, b.grad, ref)


def test_dense_no_grad_and_mixed_handles():
    """Loss only (no backward sweep), and one lazy handle mixed with one real tensor."""
    n, d, tau = 128, 128, 0.5
    z1, z2, pos_w, neg_w = _problem(n, d, 5, True)
    dev = _dev()
    with torch.no_grad():
        loss = ops.vanila_weights_contrastive_loss(z1.to(dev), z2.to(dev), pos_w.to(dev), neg_w.to(dev), tau)
    rl = float(R.port_loss(z1.double(), z2.double(), pos_w.double(), neg_w.double(), tau))
    assert abs(float(loss) - rl) <= LOSS_RTOL * abs(rl)
    z1s, z2s, j1, j2 = synth.make_batch(n, d, seed=3, joints="hand")
    j1d, j2d = j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    hp, hn = ops.get_weights_linear(j1d, j2d, "mpjpe")
    full = ops.vanila_weights_contrastive_loss(z1s.to(dev), z2s.to(dev), hp, hn, tau)
    mixed = ops.vanila_weights_contrastive_loss(z1s.to(dev), z2s.to(dev), hp.materialize(), hn, tau)
    assert abs(float(full) - float(mixed)) <= LOSS_RTOL * abs(float(full))
