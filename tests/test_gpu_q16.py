"""The relaxed-weights mode of the fused step (SMH_DIMS_Q16_TILES, `exact_weights=False`): the sweeps read a 16-bit
fixed-point image of the joint distances built from approximate square roots.

What is claimed (simhand_b200/ops.py: step_flags, include/simhand_b200.h) and checked here against the CPU oracle
(oracle/smh_oracle.c: utils.py:251-255):
    Dbound = 2 max_i D(i, 0) >= Dmax            (triangle inequality; no stored value can overflow 16 bits)
    |D_stored - D| <= Dbound / 130000 + 2e-6 D  (half a quantisation step + approximate sqrt + fp32 sum)
    |W_stored - W| <= 1.6e-5                    (NOT the 1-ulp contract of the weights API: that needs exact_weights=True)
    |Dmax_stored - Dmax| <= 4e-7 Dmax
and the loss / gradient tolerances of BASELINE.json hold on top of it.
"""
import numpy as np
import pytest
import torch

from oracle import restate as R
from simhand_b200 import _lib, layouts as L, ops, synth

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _line_batch(n, outlier):
    """Rigidly shifted copies of one hand along x: D_ij = |t_i - t_j| exactly representable structure.
    outlier=True: sample 0 sits at one end (Dbound = 2 Dmax: coarsest image); False: sample 0 in the middle
    (Dbound = Dmax (1 + 1e-4): the largest distance maps to the top of the 16-bit range)."""
    g = torch.Generator().manual_seed(11)
    hand = torch.rand(1, 21, 2, generator=g) * 40.0 + 10.0
    t = torch.linspace(-30.0, 30.0, 2 * n)
    t = t[torch.randperm(2 * n, generator=g)]
    if outlier:
        t[0] = 30.0
        t[1:] = torch.clamp(t[1:], max=29.0)
        t[1] = -30.0
    else:
        t[0] = 0.0
        t[1], t[2] = -30.0, 30.0
    j = hand + torch.stack([t, torch.zeros_like(t)], -1)[:, None, :]
    j = torch.cat([j, torch.ones(2 * n, 21, 1)], -1).contiguous()
    z1, z2, _, _ = synth.make_embeddings(n, 128, 3)
    return z1, z2, j[:n].contiguous(), j[n:].contiguous()


def _batches():
    return {
        "hand_n1024": lambda: synth.make_batch(1024, 128, 5, "hand"),
        "uniform_n777": lambda: synth.make_batch(777, 128, 9, "uniform"),
        "line_outlier0_n512": lambda: _line_batch(512, True),
        "line_centre0_n512": lambda: _line_batch(512, False),
    }


@pytest.mark.parametrize("name", sorted(_batches()))
def test_stored_q16_tiles_against_oracle(name):
    dev = _dev()
    z1, z2, j1, j2 = _batches()[name]()
    n = z1.shape[0]
    m = 2 * n
    a, b = j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    loss, dz1, dz2, aux = ops.run_step(z1.to(dev), z2.to(dev), a, b, 0.5, "fp16", True, return_aux=True,
                                       exact_weights=False)
    torch.cuda.synchronize()
    ctx, lay = aux["ctx"], aux["ctx"].layout
    assert ctx.dims.flags == _lib.DIMS_Q16_TILES
    stats = aux["stats"].cpu().numpy()
    assert stats[6] == 0
    dmax_gpu = float(stats.view(np.float32)[0])
    dbound_half = float(stats.view(np.float32)[10])
    bj = R.pack_joints(j1[:, :, :2], j2[:, :, :2])
    dmax, dmin = R.c_minmax(bj)
    dmax = float(dmax)
    assert abs(dmax_gpu - dmax) <= 4e-7 * dmax
    dbound = 2.0 * dbound_half
    assert dbound >= dmax                                         # nothing can overflow the 16-bit range
    assert dbound <= 2.0 * dmax * 1.001
    qscale = np.float32(L.Q16_LEVELS) / np.float32(dbound)
    step = 1.0 / float(qscale)

    h, tiles, tasks, strips = L.parse_plan(ctx.plan_host.numpy())
    raw = aux["ws"][int(lay.off_dist):int(lay.off_dist) + int(lay.n_stored_tiles) * L.TILE_FLOATS * 2].cpu().numpy()
    q_all = raw.view(np.uint16).reshape(int(lay.n_stored_tiles), L.TILE_FLOATS)
    rr, cc = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    idx = L.distq_index(rr, cc)
    rng = np.random.default_rng(0)
    pick = rng.choice(len(tiles), min(len(tiles), 24), replace=False)
    # always include the tiles holding the largest distances
    d_cache = {}
    worst_d = worst_w = 0.0
    top_q = 0
    for t in pick:
        I, J = int(tiles[t][0]), int(tiles[t][1])
        if I not in d_cache:
            r1 = min(m, (I + 1) * 128)
            d_cache[I] = R.c_mpjpe_rows(bj, I * 128, r1)
        want = d_cache[I][:, J * 128:min(m, (J + 1) * 128)].astype(np.float64)
        got_q = q_all[t][idx][:want.shape[0], :want.shape[1]].astype(np.float64)
        top_q = max(top_q, int(got_q.max()))
        got = got_q * step
        err = np.abs(got - want)
        tol = 0.5 * step * 1.001 + 2e-6 * want
        assert (err <= tol).all(), (name, I, J, float(err.max()), step)
        worst_d = max(worst_d, float(err.max()))
        worst_w = max(worst_w, float(err.max()) / dmax)
    assert worst_w <= 1.6e-5, worst_w
    assert top_q <= 65000
    print(f"[q16 {name}] Dmax {dmax:.4f} Dbound {dbound:.4f} step {step:.2e} max|dD| {worst_d:.2e} "
          f"max|dW| {worst_w:.2e} top q {top_q}")

    # and the step on top of that image against the oracle
    ref = R.c_step(z1, z2, j1[:, :, :2], j2[:, :, :2])
    assert abs(float(loss) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    cos, mx = R.grad_metrics(torch.cat([dz1, dz2]).cpu().numpy(), np.concatenate([ref["dz1"], ref["dz2"]]))
    assert cos >= 0.9999 and mx <= 1e-3, (cos, mx)


@pytest.mark.parametrize("engine", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("exact", [False, True])
def test_weight_modes_match_reference_goldens(golden, engine, exact):
    """Both distance images under every tensor-core engine against the reference's own outputs (tests/golden)."""
    dev = _dev()
    z1, z2 = torch.from_numpy(golden["z1"]).to(dev), torch.from_numpy(golden["z2"]).to(dev)
    a = torch.from_numpy(golden["joints1"]).to(dev)[:, :, :2]
    b = torch.from_numpy(golden["joints2"]).to(dev)[:, :, :2]
    if z1.shape[0] < 8:
        pytest.skip("tensor-core logits over < 16 samples do not average to 1e-5")
    assert ops.step_flags(engine, exact_weights=exact) == (0 if exact else _lib.DIMS_Q16_TILES)
    loss, g1, g2, aux = ops.run_step(z1, z2, a, b, 0.5, engine, True, return_aux=True, exact_weights=exact)
    assert aux["ctx"].dims.flags == (0 if exact else _lib.DIMS_Q16_TILES)
    ref = float(golden["loss_f64"])
    assert abs(float(loss) - ref) <= (1e-3 if engine == "bf16" else 1e-5) * abs(ref)
    for got, key in ((g1, "dz1_f64"), (g2, "dz2_f64")):
        cos, mx = R.grad_metrics(got.cpu().numpy(), golden[key])
        assert cos >= 0.9999 and mx <= 1e-3
    dmax_ref = float(golden["neg_dmax"]) if "neg_dmax" in golden else None
    if exact and dmax_ref is not None:
        assert float(aux["stats"].cpu().numpy().view(np.float32)[0]) == dmax_ref
