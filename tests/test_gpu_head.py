"""The projection head fused with the first normalisation (reference: simclr_model.py:22-39 +
F.normalize, simhand_w_model.py:45-58) against plain PyTorch references of the same op:
  * fp32 reference with the autocast roundings made explicit (16-bit Linear-1 output, 16-bit post-ReLU activation): tight;
  * pure fp32 reference and the eager autocast composition the kernels replace: 16-bit-activation tolerance."""
import numpy as np
import pytest
import torch
from torch import nn

from simhand_b200.head import FusedProjectionHead, reference_head_forward

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _make(rows, in_dim, hidden, out_dim, seed, dev):
    torch.manual_seed(seed)
    seq = nn.Sequential(nn.Linear(in_dim, hidden, bias=True), nn.BatchNorm1d(hidden), nn.ReLU(),
                        nn.Linear(hidden, out_dim, bias=False)).to(dev)
    with torch.no_grad():
        seq[1].weight.uniform_(0.5, 1.5)
        seq[1].bias.uniform_(-0.3, 0.3)
        seq[0].bias.uniform_(-0.5, 0.5)
    x = torch.relu(torch.randn(rows, in_dim, device=dev)) * 0.7          # ResNet encodings are post-ReLU averages
    return seq, x


def _explicit_reference(seq, x, act):
    """fp32 math with the two roundings autocast makes (Linear outputs in the 16-bit type)."""
    lin1, bn, _, lin2 = seq[0], seq[1], seq[2], seq[3]
    x16, w1, w2 = x.to(act).double(), lin1.weight.detach().to(act).double(), lin2.weight.detach().to(act).double()
    h = (x16 @ w1.t() + lin1.bias.detach().double()).to(act).double()
    mean, var = h.mean(0), h.var(0, unbiased=False)
    hn = (h - mean) / torch.sqrt(var + bn.eps) * bn.weight.detach().double() + bn.bias.detach().double()
    a = torch.relu(hn).to(act).double()
    p = a @ w2.t()
    return (p / p.norm(dim=1, keepdim=True).clamp_min(1e-12)).float(), mean.float(), var.float()


@pytest.mark.parametrize("rows,in_dim,hidden,act", [(16384, 2048, 512, torch.bfloat16), (2048, 2048, 512, torch.float16),
                                                    (300, 512, 256, torch.bfloat16), (1000, 1024, 1024, torch.bfloat16)])
def test_forward_matches_references(rows, in_dim, hidden, act):
    dev = _dev()
    seq, x = _make(rows, in_dim, hidden, 128, 3, dev)
    rm0, rv0 = seq[1].running_mean.clone(), seq[1].running_var.clone()
    fused = FusedProjectionHead(seq, act_dtype=act).train()
    y, norm = fused(x, return_norm=True)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(norm).all()
    want, mean, var = _explicit_reference(seq, x, act)
    scale = float(want.abs().max())
    err = float((y - want).abs().max()) / scale
    assert err <= 2e-3, err
    cos = torch.nn.functional.cosine_similarity(y, want, dim=1)
    assert float(cos.min()) >= 0.99999
    assert torch.allclose(y.norm(dim=1), torch.ones(rows, device=dev), atol=1e-5)
    # running estimates as nn.BatchNorm1d updates them (momentum 0.1, unbiased variance)
    n = rows
    assert torch.allclose(seq[1].running_mean, 0.9 * rm0 + 0.1 * mean, rtol=1e-3, atol=1e-4)
    assert torch.allclose(seq[1].running_var, 0.9 * rv0 + 0.1 * var * n / (n - 1), rtol=2e-3, atol=1e-4)
    assert int(seq[1].num_batches_tracked) == 1
    # pure fp32 reference of the same op and the eager autocast composition: 16-bit activation noise
    seq32 = nn.Sequential(nn.Linear(in_dim, hidden), nn.BatchNorm1d(hidden), nn.ReLU(), nn.Linear(hidden, 128, bias=False)).to(dev)
    seq32.load_state_dict(seq.state_dict())
    seq32.train()
    y32 = torch.nn.functional.normalize(seq32(x), dim=1)
    assert float((y - y32).abs().max()) / scale <= 3e-2
    assert float(torch.nn.functional.cosine_similarity(y, y32, dim=1).min()) >= 0.9995
    print(f"[head fwd rows={rows} in={in_dim} hidden={hidden} {act}] vs explicit-rounding fp32 {err:.2e}, "
          f"vs pure fp32 {float((y - y32).abs().max()) / scale:.2e} (of max|y|)")


def test_eval_mode_uses_running_estimates():
    dev = _dev()
    seq, x = _make(512, 2048, 512, 128, 5, dev)
    with torch.no_grad():
        seq[1].running_mean.uniform_(-0.2, 0.2)
        seq[1].running_var.uniform_(0.5, 2.0)
    fused = FusedProjectionHead(seq, act_dtype=torch.bfloat16).eval()
    seq.eval()
    y = fused(x)
    want = reference_head_forward(seq, x, torch.bfloat16)
    assert float((y - want).abs().max()) / float(want.abs().max()) <= 3e-2
    assert float(torch.nn.functional.cosine_similarity(y, want, dim=1).min()) >= 0.9995


def _ste(t, act):
    """round to the 16-bit activation type in the value, identity in the gradient"""
    return t + (t.to(act).to(t.dtype) - t).detach()


@pytest.mark.parametrize("rows,in_dim,hidden", [(4096, 2048, 512), (300, 512, 256)])
def test_backward_matches_autograd(rows, in_dim, hidden):
    """Gradients of a random linear functional of the normalised projections w.r.t. the encodings and every parameter.
    Tight: fp64 autograd of the composition with the SAME two activation roundings (so the ReLU masks coincide; what is left
    is the 16-bit storage of dP / dH and of the library GEMMs' outputs).  Loose: pure fp32 autograd, where ~0.3 % of the ReLU
    masks differ because Linear-1's output is rounded to 16 bits -- the eager autocast step the kernels replace differs from
    fp32 by the same amount, which the test shows beside it."""
    dev = _dev()
    act = torch.bfloat16
    seq, x = _make(rows, in_dim, hidden, 128, 7, dev)
    cot = torch.randn(rows, 128, device=dev)
    fused = FusedProjectionHead(seq, act_dtype=act).train()
    xf = x.clone().requires_grad_(True)
    y = fused(xf)
    (y * cot).sum().backward()
    got = dict(x=xf.grad.float(), w1=seq[0].weight.grad.float(), b1=seq[0].bias.grad.float(), g=seq[1].weight.grad.float(),
               b=seq[1].bias.grad.float(), w2=seq[3].weight.grad.float())
    seq.zero_grad()

    # (1) same roundings, fp64
    lin1, bn, _, lin2 = seq[0], seq[1], seq[2], seq[3]
    xd = x.to(act).double().requires_grad_(True)
    w1 = lin1.weight.detach().to(act).double().requires_grad_(True)
    w2 = lin2.weight.detach().to(act).double().requires_grad_(True)
    b1 = lin1.bias.detach().double().requires_grad_(True)
    gam, bet = bn.weight.detach().double().requires_grad_(True), bn.bias.detach().double().requires_grad_(True)
    h = _ste(xd @ w1.t() + b1, act)
    hn = (h - h.mean(0)) / torch.sqrt(h.var(0, unbiased=False) + bn.eps) * gam + bet
    a = _ste(torch.relu(hn), act)
    p = a @ w2.t()
    yd = p / p.norm(dim=1, keepdim=True).clamp_min(1e-12)
    (yd * cot.double()).sum().backward()
    tight = dict(x=xd.grad, w1=w1.grad, g=gam.grad, b=bet.grad, w2=w2.grad)

    # (2) pure fp32, and the eager autocast composition against it
    seq32 = nn.Sequential(nn.Linear(in_dim, hidden), nn.BatchNorm1d(hidden), nn.ReLU(), nn.Linear(hidden, 128, bias=False)).to(dev)
    seq32.load_state_dict(seq.state_dict())
    seq32.train()
    xr = x.clone().requires_grad_(True)
    (torch.nn.functional.normalize(seq32(xr), dim=1) * cot).sum().backward()
    ref = dict(x=xr.grad, w1=seq32[0].weight.grad, b1=seq32[0].bias.grad, g=seq32[1].weight.grad, b=seq32[1].bias.grad,
               w2=seq32[3].weight.grad)
    xe = x.clone().requires_grad_(True)
    (reference_head_forward(seq, xe, act) * cot).sum().backward()
    eager = dict(x=xe.grad.float(), w1=seq[0].weight.grad.float(), g=seq[1].weight.grad.float(), b=seq[1].bias.grad.float(),
                 w2=seq[3].weight.grad.float())

    def metrics(u, v):
        u, v = u.flatten().double(), v.flatten().double()
        return float(u @ v / (u.norm() * v.norm())), float((u - v).abs().max() / v.abs().max())

    for k in ("x", "w1", "g", "b", "w2"):
        cos_t, mx_t = metrics(got[k], tight[k])
        cos_l, mx_l = metrics(got[k], ref[k])
        cos_e, mx_e = metrics(eager[k], ref[k])
        print(f"[head bwd rows={rows}] d{k}: vs same-rounding fp64 cos {cos_t:.6f} max {mx_t:.2e} | vs fp32 cos {cos_l:.6f} "
              f"max {mx_l:.2e} (eager autocast vs fp32: cos {cos_e:.6f} max {mx_e:.2e})")
        assert cos_t >= 0.9995 and mx_t <= 3e-2, (k, cos_t, mx_t)
        assert cos_l >= min(0.995, cos_e - 2e-3), (k, cos_l, cos_e)
    # BatchNorm in training mode cancels the bias: the reference's own gradient is rounding noise around zero
    assert float(got["b1"].abs().max()) == 0.0
    assert float(ref["b1"].abs().max()) <= 1e-4 * float(ref["b"].abs().max()) + 1e-6
