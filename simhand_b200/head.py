"""Host mirror of the reference's projection head fused with the first normalisation.

Reference: `src/models/unsupervised/simclr_model.py:22-39` builds
    nn.Sequential(Linear(in, hidden, bias=True), BatchNorm1d(hidden), ReLU(), Linear(hidden, out, bias=False))
and the w-models apply `F.normalize` to its output first thing (`simhand_w_model.py:45-58`).  `FusedProjectionHead` wraps
that very Sequential (sharing its Parameters and buffers, so optimizers, checkpoints and `state_dict` keys are untouched) and
runs `normalize(head(x))` through libsimhand_b200.so: two tcgen05 kernels forward, two fused kernels plus three library
GEMMs backward (csrc/smh_head.cu).  16-bit operands as under the reference's autocast; CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib
from ._lib import check


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, gamma, beta, w2, running_mean, running_var, training, bn_eps, momentum, norm_eps, act_dtype):
        if not x.is_cuda:
            raise RuntimeError(f"simhand_b200: encodings are on {x.device}; the head runs only on a CUDA sm_100 device "
                               "(there is no CPU fallback)")
        lib = _lib.load()
        dev = x.device
        rows, in_dim = x.shape
        hidden, out_dim = w1.shape[0], w2.shape[0]
        x16 = x.detach().to(act_dtype)
        if x16.stride(1) != 1 or (x16.stride(0) * 2) % 16 or x16.data_ptr() % 16:
            x16 = x16.contiguous()
        w1_16, w2_16 = w1.detach().to(act_dtype).contiguous(), w2.detach().to(act_dtype).contiguous()
        f32 = lambda t: t.detach().float().contiguous()                                     # noqa: E731
        b1f, gf, bf = f32(b1), f32(gamma), f32(beta)
        h = torch.empty((rows, hidden), dtype=act_dtype, device=dev)
        colsum = torch.empty((2, hidden), dtype=torch.float32, device=dev)
        save_mean = torch.empty(hidden, dtype=torch.float32, device=dev)
        save_rstd = torch.empty(hidden, dtype=torch.float32, device=dev)
        y = torch.empty((rows, out_dim), dtype=torch.float32, device=dev)
        norm = torch.empty(rows, dtype=torch.float32, device=dev)
        hd = _lib.Head()
        hd.rows, hd.in_dim, hd.hidden, hd.out_dim = rows, in_dim, hidden, out_dim
        hd.fp16 = 1 if act_dtype == torch.float16 else 0
        hd.training = 1 if training else 0
        hd.x, hd.x_row_stride = x16.data_ptr(), x16.stride(0)
        hd.w1, hd.b1, hd.gamma, hd.beta = w1_16.data_ptr(), b1f.data_ptr(), gf.data_ptr(), bf.data_ptr()
        hd.running_mean, hd.running_var = _ptr(running_mean), _ptr(running_var)
        hd.bn_eps, hd.bn_momentum = bn_eps, momentum
        hd.w2, hd.h, hd.colsum = w2_16.data_ptr(), h.data_ptr(), colsum.data_ptr()
        hd.save_mean, hd.save_rstd, hd.y, hd.norm, hd.norm_eps = (save_mean.data_ptr(), save_rstd.data_ptr(), y.data_ptr(),
                                                                  norm.data_ptr(), norm_eps)
        with torch.cuda.device(dev):
            check(lib.smh_head_forward(ctypes.byref(hd), _stream_ptr(dev)), "smh_head_forward")
        ctx.save_for_backward(x16, w1_16, w2_16, b1f, gf, bf, h, save_mean, save_rstd, y, norm)
        ctx.cfg = (rows, in_dim, hidden, out_dim, hd.fp16, hd.training, bn_eps, momentum, norm_eps, x.dtype, w1.dtype, w2.dtype,
                   b1.dtype, gamma.dtype)
        ctx.mark_non_differentiable(norm)
        return y, norm

    @staticmethod
    def backward(ctx, dy, _dnorm):
        x16, w1_16, w2_16, b1f, gf, bf, h, save_mean, save_rstd, y, norm = ctx.saved_tensors
        rows, in_dim, hidden, out_dim, fp16, training, bn_eps, momentum, norm_eps, xdt, w1dt, w2dt, b1dt, gdt = ctx.cfg
        if not training:
            raise RuntimeError("simhand_b200: the fused head differentiates the training-mode BatchNorm only")
        lib = _lib.load()
        dev = x16.device
        act = x16.dtype
        dyc = dy.detach().float().contiguous()
        w2t = w2_16.t().contiguous()
        dp = torch.empty((rows, out_dim), dtype=act, device=dev)
        a = torch.empty((rows, hidden), dtype=act, device=dev)
        dhn = torch.empty((rows, hidden), dtype=act, device=dev)
        dh = torch.empty((rows, hidden), dtype=act, device=dev)
        colsum = torch.empty((2, hidden), dtype=torch.float32, device=dev)
        dgamma = torch.empty(hidden, dtype=torch.float32, device=dev)
        dbeta = torch.empty(hidden, dtype=torch.float32, device=dev)
        hd = _lib.Head()
        hd.rows, hd.in_dim, hd.hidden, hd.out_dim, hd.fp16, hd.training = rows, in_dim, hidden, out_dim, fp16, training
        hd.x, hd.x_row_stride = x16.data_ptr(), x16.stride(0)
        hd.w1, hd.b1, hd.gamma, hd.beta = w1_16.data_ptr(), b1f.data_ptr(), gf.data_ptr(), bf.data_ptr()
        hd.running_mean = hd.running_var = None
        hd.bn_eps, hd.bn_momentum = bn_eps, momentum
        hd.w2, hd.h, hd.colsum = w2_16.data_ptr(), h.data_ptr(), colsum.data_ptr()
        hd.save_mean, hd.save_rstd, hd.y, hd.norm, hd.norm_eps = (save_mean.data_ptr(), save_rstd.data_ptr(), y.data_ptr(),
                                                                  norm.data_ptr(), norm_eps)
        bw = _lib.HeadBwd()
        bw.dy, bw.w2t, bw.dp, bw.a, bw.dhn, bw.dh = (dyc.data_ptr(), w2t.data_ptr(), dp.data_ptr(), a.data_ptr(),
                                                      dhn.data_ptr(), dh.data_ptr())
        bw.colsum, bw.dgamma, bw.dbeta = colsum.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr()
        with torch.cuda.device(dev):
            check(lib.smh_head_backward(ctypes.byref(hd), ctypes.byref(bw), _stream_ptr(dev)), "smh_head_backward")
        need = ctx.needs_input_grad
        # the three plain GEMMs (library): K = rows for the weight gradients, K = hidden for the input gradient
        dw2 = torch.matmul(dp.t(), a).to(w2dt) if need[5] else None                  # [out, hidden]
        dw1 = torch.matmul(dh.t(), x16).to(w1dt) if need[1] else None                # [hidden, in]
        dx = torch.matmul(dh, w1_16).to(xdt) if need[0] else None                    # [rows, in]
        # BatchNorm in training mode removes any per-column constant: d loss / d b1 == 0 identically
        db1 = torch.zeros(hidden, dtype=b1dt, device=dev) if need[2] else None
        return (dx, dw1, db1, dgamma.to(gdt) if need[3] else None, dbeta.to(gdt) if need[4] else None, dw2,
                None, None, None, None, None, None, None)


class FusedProjectionHead(nn.Module):
    """`F.normalize(projection_head(x))` of the reference in fused kernels.  Wraps the reference's own `nn.Sequential`
    (`simclr_model.py:22-39`): parameters, buffers and `state_dict` stay where they are.

        model.projection_head = simhand_b200.FusedProjectionHead(model.projection_head)      # returns normalised rows

    forward(x) -> `[rows, out]` fp32, rows L2-normalised (what `simhand_w_model.py:56-58` computes next anyway; a second
    `F.normalize` on it is the identity up to rounding)."""

    def __init__(self, sequential: nn.Sequential, act_dtype: torch.dtype | None = None, norm_eps: float = 1e-12):
        super().__init__()
        lin1, bn, relu, lin2 = sequential[0], sequential[1], sequential[2], sequential[3]
        if not (isinstance(lin1, nn.Linear) and isinstance(bn, nn.BatchNorm1d) and isinstance(relu, nn.ReLU) and
                isinstance(lin2, nn.Linear) and lin1.bias is not None and lin2.bias is None and bn.affine):
            raise ValueError("expected Sequential(Linear(bias=True), BatchNorm1d, ReLU, Linear(bias=False))")
        self.seq = sequential
        self.act_dtype, self.norm_eps = act_dtype, norm_eps

    def forward(self, x: torch.Tensor, return_norm: bool = False):
        lin1, bn, _, lin2 = self.seq[0], self.seq[1], self.seq[2], self.seq[3]
        act = self.act_dtype
        if act is None:
            act = x.dtype if x.dtype in (torch.bfloat16, torch.float16) else (
                torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.bfloat16)
        momentum = bn.momentum if bn.momentum is not None else 0.1
        training = self.training or not bn.track_running_stats
        if training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        y, norm = _HeadFn.apply(x, lin1.weight, lin1.bias, bn.weight, bn.bias, lin2.weight,
                                bn.running_mean if bn.track_running_stats else None,
                                bn.running_var if bn.track_running_stats else None, training, float(bn.eps), float(momentum),
                                float(self.norm_eps), act)
        return (y, norm) if return_norm else y


def reference_head_forward(sequential: nn.Sequential, x: torch.Tensor, act_dtype=torch.bfloat16) -> torch.Tensor:
    """The eager composition the fused head replaces (library GEMMs + elementwise kernels under autocast), for tests and
    for the timing comparison of bench_head.py."""
    with torch.autocast("cuda", dtype=act_dtype):
        p = sequential(x)
    return torch.nn.functional.normalize(p.float(), dim=1)
