"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(simhand_b200.ops -> libsimhand_b200.so) and is checked against the golden vectors produced by the
reference's own functions (tests/golden, oracle/gen_golden.py) and against the CPU oracle (oracle/).

Tolerances are the ones BASELINE.json states:
  weights           <= 1 ulp of fp32 (we assert 0 ulp: the kernels mirror torch-CPU's operation order)
  fp32 / tf32 mode  loss within 1e-5 relative
  gradients         cosine >= 0.9999 and max|err| <= 1e-3 max|grad| (stated for bf16; the tf32 and fp32
                    engines are held to it as an upper bound and to tighter figures below)
"""
import numpy as np
import pytest
import torch

from oracle import restate as R
from simhand_b200 import _lib, ops, synth

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_COS = 0.9999
GRAD_MAXABS = 1e-3


def _dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _to_dev(g):
    dev = _dev()
    j1 = torch.from_numpy(g["joints1"]).to(dev)
    j2 = torch.from_numpy(g["joints2"]).to(dev)
    return (torch.from_numpy(g["z1"]).to(dev), torch.from_numpy(g["z2"]).to(dev), j1[:, :, :2], j2[:, :, :2])


@pytest.mark.parametrize("which,name", [(0, "sqrt"), (1, "sqrt2"), (2, "div21"), (3, "divw")])
def test_exact_math_selftests(which, name):
    """The branch-free exact sqrt / division forms equal the IEEE intrinsics on their whole domain."""
    lib = _lib.load()
    out = torch.zeros(8, dtype=torch.int64, device=_dev())
    _lib.check(lib.smh_selftest(which, out.data_ptr(), 8, torch.cuda.current_stream().cuda_stream), name)
    tested, bad, first, max_ulp = out.cpu().tolist()[:4]
    assert tested > 1_000_000
    assert bad == 0, f"{name}: {bad} mismatches of {tested}, first input bits {first:#x}, max {max_ulp} ulp"


def test_fma_pipe_sqrt_selftest():
    """The MUFU-free square root of the 16-bit tile image (sqrt2_fma_pipe) stays within 7.5e-7 relative of the true square
    root for x = 0 and every float of the kernel's domain; packed and scalar forms agree bit for bit."""
    lib = _lib.load()
    out = torch.zeros(8, dtype=torch.int64, device=_dev())
    _lib.check(lib.smh_selftest(4, out.data_ptr(), 8, torch.cuda.current_stream().cuda_stream), "sqrt_fma_pipe")
    tested, bad, first, max_rel_e9 = out.cpu().tolist()[:4]
    print(f"sqrt_fma_pipe: {tested} values, max relative error {max_rel_e9 * 1e-9:.3e}")
    assert tested > 1_000_000_000
    assert bad == 0, f"{bad} of {tested} outside the bound, first input bits {first:#x}, max rel {max_rel_e9 * 1e-9:.3e}"


def test_weights_match_reference_bitwise(golden):
    z1, z2, a, b = _to_dev(golden)
    pos_w, neg_w = ops.mpjpe_weights(a, b)
    assert R.ulp_distance(pos_w.cpu().numpy(), golden["pos_w"]).max() == 0
    assert R.ulp_distance(neg_w.cpu().numpy(), golden["neg_w"]).max() == 0


@pytest.mark.parametrize("engine", ["fp32", "tf32", "auto", "bf16", "fp16"])
def test_step_matches_reference(golden, engine):
    z1, z2, a, b = _to_dev(golden)
    if engine in ("tf32", "fp16") and z1.shape[0] < 8:
        pytest.skip("tf32 logits over < 16 samples do not average to 1e-5; 'auto' picks the fp32 engine there")
    loss_rtol = 1e-3 if engine == "bf16" else LOSS_RTOL          # BASELINE.json: bf16 mode within 1e-3
    loss, dz1, dz2, aux = ops.run_step(z1, z2, a, b, 0.5, engine, True, return_aux=True)
    stats = aux["stats"].cpu().numpy()
    assert stats[6] == 0, f"pipeline wait timed out at site {stats[6]}"
    ref = float(golden["loss_f64"])
    assert abs(float(loss) - ref) <= loss_rtol * abs(ref), (float(loss), ref)
    for got, key in ((dz1, "dz1_f64"), (dz2, "dz2_f64")):
        cos, mx = R.grad_metrics(got.cpu().numpy(), golden[key])
        assert cos >= GRAD_COS and mx <= GRAD_MAXABS, (engine, key, cos, mx)
        if ops.resolve_engine(engine, z1.shape[0]) == "fp32":
            assert cos >= 1 - 1e-9 and mx <= 2e-5, (key, cos, mx)
    # row sums against the closed form on the reference weights
    _, _, _, neg = R.closed_form_fp64(torch.from_numpy(golden["z1"]), torch.from_numpy(golden["z2"]),
                                      torch.from_numpy(golden["pos_w"]), torch.from_numpy(golden["neg_w"]))
    rel = (aux["neg"].cpu().double() - neg).abs() / neg
    assert rel.max() < {"fp32": 2e-6, "tf32": 2e-4, "fp16": 2e-4, "bf16": 4e-3}[ops.resolve_engine(engine, z1.shape[0])]


def test_drop_in_api_and_autograd(golden):
    z1, z2, a, b = _to_dev(golden)
    z1 = z1.clone().requires_grad_(True)
    z2 = z2.clone().requires_grad_(True)
    pos_w, neg_w = ops.get_weights_linear(a, b, "mpjpe")
    loss = ops.vanila_weights_contrastive_loss(z1, z2, pos_w, neg_w)
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.requires_grad
    (3.0 * loss).backward()
    cos, mx = R.grad_metrics(z1.grad.cpu().numpy() / 3.0, golden["dz1_f64"])
    assert cos >= GRAD_COS and mx <= GRAD_MAXABS
    assert abs(float(loss) - float(golden["loss_f64"])) <= LOSS_RTOL * abs(float(golden["loss_f64"]))
    # the handles materialise to the reference's tensors
    assert tuple(neg_w.shape) == golden["neg_w"].shape
    assert R.ulp_distance(neg_w.materialize().cpu().numpy(), golden["neg_w"]).max() == 0
    # no-grad call skips the backward sweep and gives the same loss
    with torch.no_grad():
        l2 = ops.weighted_ntxent(z1, z2, a, b)
    assert abs(float(l2) - float(loss)) < 1e-6


@pytest.mark.parametrize("n,jset", [(1024, "hand"), (777, "uniform")])
def test_step_matches_c_oracle_mid_size(n, jset):
    """Sizes above the golden set, against the plain-C oracle on the same seeded inputs."""
    z1, z2, j1, j2 = synth.make_batch(n, 128, 17, jset)
    a, b = j1[:, :, :2], j2[:, :, :2]
    ref = R.c_step(z1, z2, a, b)
    dev = _dev()
    for engine in ("tf32", "fp32", "bf16", "fp16"):
        loss, dz1, dz2, aux = ops.run_step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2],
                                           0.5, engine, True, return_aux=True)
        assert aux["stats"].cpu().numpy()[6] == 0
        assert abs(float(loss) - ref["loss"]) <= (1e-3 if engine == "bf16" else LOSS_RTOL) * abs(ref["loss"])
        cos, mx = R.grad_metrics(torch.cat([dz1, dz2]).cpu().numpy(), np.concatenate([ref["dz1"], ref["dz2"]]))
        assert cos >= GRAD_COS and mx <= GRAD_MAXABS, (engine, cos, mx)
        stats = aux["stats"].cpu().numpy().view(np.float32)
        assert stats[1] == ref["stats"]["pmax"]
        if aux["ctx"].dims.flags & _lib.DIMS_Q16_TILES:
            # relaxed-weights mode (the tensor-core engines' default): Dmax comes from the approximate square roots
            assert abs(float(stats[0]) - float(ref["stats"]["dmax"])) <= 4e-7 * float(ref["stats"]["dmax"])
        else:
            assert stats[0] == ref["stats"]["dmax"]
        # the exact-weights mode reproduces Dmax bit for bit in every engine
        if engine != "fp32":
            _, _, _, aux_x = ops.run_step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2],
                                          0.5, engine, True, return_aux=True, exact_weights=True)
            assert aux_x["stats"].cpu().numpy().view(np.float32)[0] == ref["stats"]["dmax"]
        rel = np.abs(aux["neg"].cpu().numpy().astype(np.float64) - ref["neg"]) / ref["neg"]
        assert rel.max() < (4e-3 if engine == "bf16" else 2e-4)


@pytest.mark.parametrize("variant", ["pos", "neg", "plain"])
def test_other_weightings_match_reference(golden, variant):
    """pos-only, neg-only and unweighted NT-Xent (utils.py:430, :468, :157) through the same kernels."""
    z1, z2, a, b = _to_dev(golden)
    z1 = z1.clone().requires_grad_(True)
    z2 = z2.clone().requires_grad_(True)
    pw, nw = ops.get_weights_linear(a, b, "mpjpe")
    if variant == "pos":
        loss = ops.vanila_pos_weights_contrastive_loss(z1, z2, pw)
    elif variant == "neg":
        loss = ops.vanila_neg_weights_contrastive_loss(z1, z2, nw)
    else:
        loss = ops.vanila_contrastive_loss(z1, z2)
    loss.backward()
    ref = float(golden[f"loss_{variant}_f64"])
    assert abs(float(loss.detach()) - ref) <= LOSS_RTOL * abs(ref), (variant, float(loss.detach()), ref)
    cos, mx = R.grad_metrics(z1.grad.cpu().numpy(), golden[f"dz1_{variant}_f64"])
    assert cos >= GRAD_COS and mx <= GRAD_MAXABS, (variant, cos, mx)


def test_non_contiguous_and_half_inputs():
    dev = _dev()
    z1, z2, j1, j2 = synth.make_batch(96, 128, 23, "hand")
    base, _, _ = ops.run_step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2])
    # strided z rows (a column slice of a wider tensor) and joints given as contiguous copies
    wide1 = torch.zeros(96, 160, device=dev)
    wide2 = torch.zeros(96, 160, device=dev)
    wide1[:, 16:144] = z1.to(dev)
    wide2[:, 16:144] = z2.to(dev)
    l2, _, _ = ops.run_step(wide1[:, 16:144], wide2[:, 16:144], j1[:, :, :2].contiguous().to(dev),
                            j2[:, :, :2].contiguous().to(dev))
    assert float(l2) == float(base)
    # fp16 projections under autocast are promoted to fp32 before the op (reference trains with precision=16)
    with torch.autocast("cuda", dtype=torch.float16):
        l3 = ops.weighted_ntxent(z1.to(dev).half(), z2.to(dev).half(), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2])
    assert l3.dtype == torch.float32 and abs(float(l3) - float(base)) < 2e-3 * abs(float(base))


def test_reference_nan_edges():
    """N = 1 (pos max == pos min) and identical joints (Dmax == 0) give 0/0 in the reference
    (utils.py:235, :259); the CUDA path returns NaN as well."""
    dev = _dev()
    z1, z2, j1, j2 = synth.make_batch(1, 128, 3, "uniform")
    loss, _, _ = ops.run_step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2])
    assert torch.isnan(loss)
    z1, z2, j1, _ = synth.make_batch(40, 128, 3, "uniform")
    same = j1[:1].expand(40, 21, 3).contiguous().to(dev)
    loss, _, _ = ops.run_step(z1.to(dev), z2.to(dev), same[:, :, :2], same[:, :, :2])
    assert torch.isnan(loss)


def test_cpu_tensors_are_rejected():
    z1, z2, j1, j2 = synth.make_batch(8, 128, 3, "hand")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.weighted_ntxent(z1, z2, j1[:, :, :2], j2[:, :, :2])


def test_slow_domain_inputs_use_ieee_path():
    """Joints outside the fast exact-sqrt domain (tiny / huge magnitudes) take the IEEE path and still match
    the oracle bit for bit."""
    dev = _dev()
    z1, z2, j1, j2 = synth.make_batch(70, 128, 29, "uniform")
    j1 = j1 * 1e-9
    j2 = j2 * 1e-9
    a, b = j1[:, :, :2], j2[:, :, :2]
    pos_w, neg_w = ops.mpjpe_weights(a.to(dev), b.to(dev))
    bj = R.pack_joints(a, b)
    dmax, dmin = R.c_minmax(bj)
    want = R.c_neg_weights_rows(bj, 0, 140, dmax, dmin)
    assert R.ulp_distance(neg_w.cpu().numpy(), want).max() == 0


def test_coincident_joints_take_the_guarded_redo():
    """The hot MPJPE loop has no zero guard: a joint shared by two different samples (zero distance in an
    off-diagonal tile) makes it produce NaN, which the kernel notices through the integer max and repairs by
    redoing the tile with the guarded form.  Result must still be bit-exact."""
    dev = _dev()
    z1, z2, j1, j2 = synth.make_batch(200, 128, 31, "uniform")
    j1[150] = j1[5]                       # duplicate sample: all 21 distances zero, tiles (0, 1)
    j2[77, 3] = j1[190, 3]                # one shared joint across the two views, tiles (1, 2)
    a, b = j1[:, :, :2], j2[:, :, :2]
    pos_w, neg_w = ops.mpjpe_weights(a.to(dev), b.to(dev))
    bj = R.pack_joints(a, b)
    dmax, dmin = R.c_minmax(bj)
    want = R.c_neg_weights_rows(bj, 0, 400, dmax, dmin)
    got = neg_w.cpu().numpy()
    assert np.isfinite(got).all()
    assert R.ulp_distance(got, want).max() == 0
    assert got[5, 150] == 1.0 and got[150, 5] == 1.0
    ref = R.c_step(z1, z2, a, b)
    loss, dz1, dz2 = ops.run_step(z1.to(dev), z2.to(dev), a.to(dev), b.to(dev), 0.5, "tf32", True)
    assert abs(float(loss) - ref["loss"]) <= LOSS_RTOL * abs(ref["loss"])
    cos, mx = R.grad_metrics(torch.cat([dz1, dz2]).cpu().numpy(), np.concatenate([ref["dz1"], ref["dz2"]]))
    assert cos >= GRAD_COS and mx <= GRAD_MAXABS


def test_l2_normalize_matches_torch():
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    for rows, d in ((8192, 128), (100, 96), (33, 130)):
        x = torch.randn(rows, d, generator=g).to(dev)
        x[0] = 0                                                   # zero row: eps branch
        x1 = x.clone().requires_grad_(True)
        x2 = x.clone().requires_grad_(True)
        w = torch.randn(rows, d, generator=g).to(dev)
        y1 = ops.l2_normalize(x1)
        y2 = torch.nn.functional.normalize(x2, dim=1)
        assert torch.allclose(y1, y2, rtol=2e-6, atol=1e-7)
        (y1 * w).sum().backward()
        (y2 * w).sum().backward()
        assert torch.allclose(x1.grad[1:], x2.grad[1:], rtol=1e-4, atol=1e-6)

