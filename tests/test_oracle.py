"""CPU tests: the oracle (oracle/) against the reference's golden vectors and, where the
reference checkout is present, against the reference's own functions executed live."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import restate as R
from simhand_b200 import synth


def _views(g):
    j1, j2 = torch.from_numpy(g["joints1"]), torch.from_numpy(g["joints2"])
    return torch.from_numpy(g["z1"]), torch.from_numpy(g["z2"]), j1[:, :, :2], j2[:, :, :2]


def test_c_oracle_weights_bit_exact_vs_golden(golden):
    z1, z2, a, b = _views(golden)
    bj = R.pack_joints(a, b)
    m = bj.shape[0]
    dmax, dmin = R.c_minmax(bj)
    w = R.c_neg_weights_rows(bj, 0, m, dmax, dmin)
    assert R.ulp_distance(w, golden["neg_w"]).max() == 0
    res = R.c_step(z1, z2, a, b, want_grad=False)
    assert R.ulp_distance(res["pos_w"], golden["pos_w"]).max() == 0


def test_c_oracle_loss_and_grad_vs_golden(golden):
    z1, z2, a, b = _views(golden)
    res = R.c_step(z1, z2, a, b)
    ref = float(golden["loss_f64"])
    assert abs(res["loss"] - ref) <= 1e-12 * abs(ref)
    for k in ("dz1", "dz2"):
        cos, mx = R.grad_metrics(res[k], golden[k + "_f64"])
        assert cos > 1 - 1e-12 and mx < 1e-11
    # the fp32 reference itself sits within its own noise of the fp64 value
    assert abs(float(golden["loss_f32"]) - ref) <= 5e-7 * abs(ref)


def test_port_matches_golden(golden):
    """Same ATen ops as the reference => same bits on this class of CPU (AVX2/AVX512)."""
    z1, z2, a, b = _views(golden)
    loss, g1, g2, pw, nw = R.port_step(z1, z2, a, b)
    assert R.ulp_distance(nw.numpy(), golden["neg_w"]).max() == 0
    assert R.ulp_distance(pw.numpy(), golden["pos_w"]).max() == 0
    assert abs(float(loss) - float(golden["loss_f32"])) <= 2e-6 * abs(float(golden["loss_f32"]))
    cos, mx = R.grad_metrics(g1.numpy(), golden["dz1_f32"])
    assert cos > 1 - 1e-9 and mx < 1e-5


def test_closed_form_matches_golden(golden):
    z1, z2, a, b = _views(golden)
    loss, g1, g2, _ = R.closed_form_fp64(z1, z2, torch.from_numpy(golden["pos_w"]),
                                         torch.from_numpy(golden["neg_w"]))
    assert abs(float(loss) - float(golden["loss_f64"])) < 1e-12
    assert np.abs(g1.numpy() - golden["dz1_f64"]).max() < 1e-14
    assert np.abs(g2.numpy() - golden["dz2_f64"]).max() < 1e-14


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")
@pytest.mark.parametrize("jset", ["hand", "uniform", "peclr"])
@pytest.mark.parametrize("n", [2, 37, 160])
def test_c_oracle_vs_live_reference(jset, n):
    ns = ref_loader.load_reference_functions()
    z1, z2, j1, j2 = synth.make_batch(n, 128, 11 + n, jset)
    a, b = j1[:, :, :2], j2[:, :, :2]
    pw, nw = ns["get_weights_linear"](a, b, "mpjpe")
    x1, x2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    loss = ns["vanila_weights_contrastive_loss"](x1, x2, pw, nw)
    loss.backward()
    res = R.c_step(z1, z2, a, b)
    bj = R.pack_joints(a, b)
    w = R.c_neg_weights_rows(bj, 0, 2 * n, res["stats"]["dmax"], res["stats"]["dmin"])
    assert R.ulp_distance(w, nw.numpy()).max() == 0
    assert R.ulp_distance(res["pos_w"], pw.numpy()).max() == 0
    assert abs(res["loss"] - float(loss.detach())) <= 1e-6 * abs(float(loss.detach()))
    cos, mx = R.grad_metrics(res["dz1"], x1.grad.numpy())
    assert cos > 1 - 1e-10 and mx < 2e-6


def test_reference_edge_n1_is_nan():
    """N = 1: pos max == pos min => 0/0 (utils.py:235); the oracle reproduces the NaN."""
    z1, z2, j1, j2 = synth.make_batch(1, 128, 3, "uniform")
    res = R.c_step(z1, z2, j1[:, :, :2], j2[:, :, :2], want_grad=False)
    assert np.isnan(res["pos_w"]).all() and np.isnan(res["loss"])


def test_properties_symmetry_and_permutation():
    z1, z2, j1, j2 = synth.make_batch(48, 128, 21, "uniform")
    a, b = j1[:, :, :2], j2[:, :, :2]
    bj = R.pack_joints(a, b)
    d = R.c_mpjpe_rows(bj, 0, 96)
    assert np.array_equal(d.view(np.uint32), d.T.copy().view(np.uint32))       # bitwise symmetric
    assert (np.diag(d) == 0).all()
    base = R.c_step(z1, z2, a, b)
    perm = torch.randperm(48, generator=torch.Generator().manual_seed(1))
    p = R.c_step(z1[perm], z2[perm], a[perm], b[perm])
    assert abs(p["loss"] - base["loss"]) < 1e-12
    assert np.abs(p["dz1"] - base["dz1"][perm.numpy()]).max() < 1e-15
    sw = R.c_step(z2, z1, b, a)                                                # view swap
    assert abs(sw["loss"] - base["loss"]) < 1e-12
    # row-chunked MPJPE == full (the property the chunked paths rely on)
    assert np.array_equal(R.c_mpjpe_rows(bj, 10, 30), d[10:30])
    assert np.array_equal(R.port_mpjpe_rows(torch.from_numpy(bj).view(96, 21, 2), 10, 30).numpy(), d[10:30])
