"""The sharded step (fused exchange, smh_shard.cu) proven on ONE GPU: every rank's plan, kernels and exchange protocol run
on the same device (simhand_b200.dist.EmulatedGroup: `world` workspaces and signal blocks, the ranks' launches issued stage
by stage), and the result is held against the CPU oracle and the single-GPU step.  The real multi-GPU run differs only in
that the peer pointers cross NVLink (tests/test_gpu_dist.py, tools/dist_check.py on a multi-GPU box)."""
import numpy as np
import pytest
import torch

from oracle import restate as R
from simhand_b200 import _lib, ops, synth
from simhand_b200.dist import EmulatedGroup

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _check_against_oracle(losses, g1s, g2s, ref, loss_rtol=1e-5):
    for r, loss in enumerate(losses):
        assert abs(float(loss) - ref["loss"]) <= loss_rtol * abs(ref["loss"]), (r, float(loss), ref["loss"])
    assert len({float(x) for x in losses}) == 1                 # every rank evaluates the same loss bits
    got = torch.cat([torch.cat(g1s), torch.cat(g2s)]).cpu().numpy()
    cos, mx = R.grad_metrics(got, np.concatenate([ref["dz1"], ref["dz2"]]))
    assert cos >= 0.9999 and mx <= 1e-3, (cos, mx)
    return cos, mx


@pytest.mark.parametrize("n,world,joints,exact", [(1024, 8, "hand", False), (1024, 8, "hand", True), (776, 8, "uniform", False),
                                                  (640, 4, "peclr", False), (512, 2, "hand", False), (64, 8, "hand", False)])
def test_emulated_ranks_match_oracle(n, world, joints, exact):
    dev = _dev()
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, joints)
    ref = R.c_step(z1, z2, j1[:, :, :2], j2[:, :, :2])
    grp = EmulatedGroup(n, 128, world, dev, "fp16", exact_weights=exact)
    a, b = j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    losses, g1s, g2s = grp.step(z1.to(dev), z2.to(dev), a, b)
    torch.cuda.synchronize()
    assert grp.poisoned() == [0] * world
    cos, mx = _check_against_oracle(losses, g1s, g2s, ref)
    # and against the single-GPU step: same kernels, same scale of the 16-bit image
    l1, s1, s2 = ops.run_step(z1.to(dev), z2.to(dev), a, b, 0.5, "fp16", True, exact_weights=exact)
    assert abs(float(losses[0]) - float(l1)) <= 2e-6 * abs(float(l1))
    c2, m2 = R.grad_metrics(torch.cat(g1s).cpu().numpy(), s1.cpu().numpy())
    assert c2 >= 0.999999 and m2 <= 2e-4, (c2, m2)


def test_emulated_full_size_against_oracle():
    """BASELINE.json configs[2]: global batch 8192 (2N = 16384) over 8 ranks, default engine and distance image."""
    dev = _dev()
    n, world = 8192, 8
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    ref = R.c_step(z1, z2, j1[:, :, :2], j2[:, :, :2])
    grp = EmulatedGroup(n, 128, world, dev, "fp16")
    a, b = j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    losses, g1s, g2s = grp.step(z1.to(dev), z2.to(dev), a, b)
    torch.cuda.synchronize()
    assert grp.poisoned() == [0] * world
    cos, mx = _check_against_oracle(losses, g1s, g2s, ref)
    print(f"[emulated 8 ranks, 2N = 16384] loss rel {abs(float(losses[0]) - ref['loss']) / abs(ref['loss']):.2e} "
          f"grad cos {cos:.9f} max err {mx:.2e}")


def test_back_to_back_steps_alternate_batches_and_loss_only():
    """Consecutive steps on different batches (the per-step scalars and positives are double-buffered by step parity; the
    accumulators are re-zeroed by the next step's first launch), then a loss-only step."""
    dev = _dev()
    n, world = 512, 4
    grp = EmulatedGroup(n, 128, world, dev, "fp16")
    batches = [synth.make_batch(n, 128, seed, kind) for seed, kind in ((5, "hand"), (6, "uniform"), (7, "peclr"))]
    refs = [R.c_step(z1, z2, j1[:, :, :2], j2[:, :, :2]) for z1, z2, j1, j2 in batches]
    for it in range(7):
        z1, z2, j1, j2 = batches[it % 3]
        losses, g1s, g2s = grp.step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2])
        _check_against_oracle(losses, g1s, g2s, refs[it % 3])
    z1, z2, j1, j2 = batches[1]
    losses, g1s, g2s = grp.step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2], want_grad=False)
    assert g1s[0] is None
    for loss in losses:
        assert abs(float(loss) - refs[1]["loss"]) <= 1e-5 * abs(refs[1]["loss"])
    assert grp.poisoned() == [0] * world


@pytest.mark.parametrize("weighting,pos,neg", [(("non_linear", "mpjpe", 2.5, 0.05), True, True),
                                               (("linear", "w_abs", 0.0, 0.0), True, True),
                                               (("non_linear", "w_o_abs", 1.0, 0.1), True, True),
                                               (("linear", "mpjpe", 0.0, 0.0), True, False),
                                               (("linear", "mpjpe", 0.0, 0.0), False, True),
                                               (("linear", "mpjpe", 0.0, 0.0), False, False)])
def test_emulated_variants_match_reference_restatement(weighting, pos, neg):
    """The other weightings on several ranks (utils.py:219-227, :241-249, :304-346, :430-501): non_linear needs the global
    mean distance (rank-ordered sum of the ranks' partial sums, delivered with stage 2)."""
    dev = _dev()
    n, world = 384, 4
    wt = ops.make_weighting(*weighting)
    z1, z2, j1, j2 = synth.make_batch(n, 128, 21, "hand")
    a_c, b_c = j1[:, :, :2], j2[:, :, :2]
    pos_w, neg_w = R.port_get_weights(a_c, b_c, weighting[0], weighting[1], weighting[2], weighting[3])
    if not pos:
        pos_w = torch.ones_like(pos_w)
    if not neg:
        neg_w = torch.ones_like(neg_w)
    want, w1, w2, _ = R.closed_form_fp64(z1, z2, pos_w, neg_w)
    grp = EmulatedGroup(n, 128, world, dev, "fp16", weighting=wt, neg_weighted=neg)
    losses, g1s, g2s = grp.step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2], pos_weighted=pos,
                                neg_weighted=neg)
    torch.cuda.synchronize()
    assert grp.poisoned() == [0] * world
    for loss in losses:
        assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want)), (float(loss), float(want))
    cos, mx = R.grad_metrics(torch.cat([torch.cat(g1s), torch.cat(g2s)]).cpu().numpy(),
                             torch.cat([w1, w2]).numpy())
    assert cos >= 0.9999 and mx <= 1e-3, (cos, mx)


def test_missing_rank_poisons_the_group_instead_of_returning_garbage():
    """A rank that never delivers (here: rank 1's first launch is withheld) makes the waits of the others time out: the
    group is poisoned on every rank, the losses of this and of later steps are NaN, nothing hangs."""
    dev = _dev()
    n, world = 256, 2
    grp = EmulatedGroup(n, 128, world, dev, "fp16")
    for ex in grp.structs:
        ex.timeout_ms = 200
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    lib = _lib.load()
    from simhand_b200.dist import _fused_launches
    from simhand_b200.ops import _stream_ptr, make_inputs
    n_local = n // world
    a, b = j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    zz1, zz2 = z1.to(dev), z2.to(dev)
    li, keep = make_inputs(zz1[:n_local], zz2[:n_local], a[:n_local], b[:n_local])
    loss = torch.zeros((), device=dev)
    g1, g2 = torch.empty(n_local, 128, device=dev), torch.empty(n_local, 128, device=dev)
    _fused_launches(lib, grp.ctxs[0], grp.structs[0], grp.ws[0].data_ptr(), li, 0.5, _lib.ENGINES["fp16"], True, 1.0,
                    (loss, g1, g2), _stream_ptr(dev))            # rank 0 alone: rank 1 never shows up
    torch.cuda.synchronize()
    assert torch.isnan(loss)
    assert all(p != 0 for p in grp.poisoned())
    losses, _, _ = grp.step(zz1, zz2, a, b)
    torch.cuda.synchronize()
    assert all(torch.isnan(x) for x in losses)                   # sticky until the exchange is rebuilt
