"""simhand_b200.HostPipeline (the host-buffer front end bench.py's `e2e` times): every batch is copied once, the losses come
back in order, with the current step (lag=0) or one step in flight (lag=1), and they equal the direct call's."""
import pytest
import torch

from simhand_b200 import ops, synth
from simhand_b200.pipeline import HostPipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lag,use_graph", [(0, True), (1, True), (1, False)])
def test_pipeline_returns_each_steps_loss(lag, use_graph):
    dev = torch.device("cuda:0")
    n = 512
    batches = [synth.make_batch(n, 128, seed, kind) for seed, kind in ((5, "hand"), (6, "uniform"), (7, "peclr"))]
    host = [tuple(t.contiguous().pin_memory() for t in b) for b in batches]
    want = []
    for z1, z2, j1, j2 in batches:
        loss, _, _ = ops.run_step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2])
        want.append(float(loss))

    def step_fn(a, b, c, e):                      # the reference's call pattern
        a, b = a.detach().requires_grad_(True), b.detach().requires_grad_(True)
        pw, nw = ops.get_weights_linear(c[:, :, :2], e[:, :, :2], "mpjpe")
        loss = ops.vanila_weights_contrastive_loss(a, b, pw, nw)
        g1, g2 = torch.autograd.grad(loss, (a, b))
        return loss, g1, g2

    pipe = HostPipeline(step_fn, host[0], dev, depth=2, use_graph=use_graph, lag=lag)
    got = []
    steps = 7
    pipe.prefetch(*host[0])
    for k in range(steps):
        if k + 1 < steps:
            pipe.prefetch(*host[(k + 1) % 3])
        loss_h, g1, g2 = pipe.step()
        if loss_h is not None:
            got.append(float(loss_h))
            assert g1.shape == (n, 128) and torch.isfinite(g1).all()
    if lag:
        loss_h, g1, g2 = pipe.drain()
        got.append(float(loss_h))
    assert len(got) == steps
    for k, val in enumerate(got):
        assert abs(val - want[k % 3]) <= 2e-6 * abs(want[k % 3]), (k, val, want[k % 3])
    with pytest.raises(RuntimeError):
        pipe.step()                                # nothing prefetched
