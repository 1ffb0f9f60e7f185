"""Host mirror of the reference's loss operator (`src/models/utils.py`), backed by libsimhand_b200.so.

Drop-in names, argument meaning and reduction follow the reference:

    get_weights_linear(joints1, joints2, diff_type)                       utils.py:218-261
    vanila_weights_contrastive_loss(z1, z2, pos_w, neg_w, temperature)    utils.py:391-427

`get_weights_linear` returns two lazy handles (they only remember the joints); handing them to
`vanila_weights_contrastive_loss` runs the fused CUDA path, in which the [2N, 2N] weight and logit
matrices are never materialised.  `handle.materialize()` gives the real tensors with the reference's
shapes (pos_w [N], neg_w [2N, 2N]) computed by the same kernels.  `weighted_ntxent` is the one-call
fused form used by the benchmark and by the sharded path (simhand_b200/dist.py).

There is no CPU or eager-PyTorch fallback: inputs must live on a CUDA (sm_100) device.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import Dims, Inputs, Layout, check

_DEFAULT_ENGINE = "auto"
_AUTO_FP32_MAX_M = 256     # below this many samples the exact-fp32 engine is used: tf32 rounding of the logits
                           # does not average out over so few negatives, and the sweep is launch-bound anyway


def resolve_engine(engine: str, n: int) -> str:
    """'auto' -> 'fp32' for tiny batches (2N <= 256), else 'fp16' (tcgen05; fp16 logit operands carry the same 11-bit
    significand as tf32 for L2-normalised z and measure bit-identical losses to the tf32 engine at half the staging
    traffic)."""
    if engine == "auto":
        return "fp32" if 2 * n <= _AUTO_FP32_MAX_M else "fp16"
    if engine not in _lib.ENGINES:
        raise ValueError(f"unknown engine {engine!r}; choose from auto, fp16, tf32, bf16, fp32")
    return engine
_ctx_lock = threading.Lock()
_ctx_cache = {}


class _Context:
    """Layout + device copy of the task plan for one (n, d, world, rank, strip_len, device)."""

    def __init__(self, n: int, d: int, world: int, rank: int, strip_len: int, device: torch.device, flags: int = 0,
                 weighting=None):
        lib = _lib.load()
        diff, wtype, lam_p, lam_n = weighting or DEFAULT_WEIGHTING
        self.dims = Dims(n, d, world, rank, strip_len, flags, _lib.DIFF_TYPES[diff], _lib.WEIGHT_TYPES[wtype],
                         float(lam_p), float(lam_n))
        self.layout = Layout()
        check(lib.smh_layout(ctypes.byref(self.dims), ctypes.byref(self.layout)), "smh_layout")
        host = torch.empty(int(self.layout.plan_bytes), dtype=torch.uint8).pin_memory() \
            if device.type == "cuda" else torch.empty(int(self.layout.plan_bytes), dtype=torch.uint8)
        check(lib.smh_plan_build(ctypes.byref(self.dims), host.data_ptr(), host.numel()), "smh_plan_build")
        self.plan_host = host
        self.plan_dev = host.to(device) if device.type == "cuda" else None
        self.device = device

    def view(self, ws: torch.Tensor, off: int, count: int, dtype=torch.float32) -> torch.Tensor:
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return ws[off:off + nbytes].view(dtype)


def make_weighting(weight_type: str = "linear", diff_type: str = "mpjpe", lambda_pos: float = 0.0,
                   lambda_neg: float = 0.0):
    """(diff_type, weight_type, lambda_pos, lambda_neg) as the reference's config names them
    (`config.weight_type`, `config.diff_type`, `config.non_linear_lambda_pos/neg`, simhand_w_model.py:106-118)."""
    if diff_type not in _lib.DIFF_TYPES:
        raise ValueError(f"diff_type must be one of {sorted(_lib.DIFF_TYPES)}, got {diff_type!r}")
    if weight_type not in _lib.WEIGHT_TYPES:
        raise ValueError(f"weight_type must be one of {sorted(_lib.WEIGHT_TYPES)}, got {weight_type!r}")
    if weight_type == "linear":
        lambda_pos = lambda_neg = 0.0
    return (diff_type, weight_type, float(lambda_pos), float(lambda_neg))


DEFAULT_WEIGHTING = ("mpjpe", "linear", 0.0, 0.0)


def get_context(n: int, d: int, world: int, rank: int, device, strip_len: int = 0, flags: int = 0,
                weighting=None) -> _Context:
    device = torch.device(device)
    weighting = tuple(weighting or DEFAULT_WEIGHTING)
    key = (n, d, world, rank, strip_len, device.type, device.index, flags, weighting)
    with _ctx_lock:
        ctx = _ctx_cache.get(key)
        if ctx is None:
            ctx = _Context(n, d, world, rank, strip_len, device, flags, weighting)
            _ctx_cache[key] = ctx
    return ctx


def exact_weights_default() -> bool:
    """Process-wide default of `exact_weights` (see step_flags): SMH_EXACT_WEIGHTS=1 or SMH_Q16=0 in the environment."""
    return os.environ.get("SMH_EXACT_WEIGHTS", "0") == "1" or os.environ.get("SMH_Q16", "1") == "0"


def step_flags(engine_name: str, weighting=None, neg_weighted: bool = True, exact_weights: Optional[bool] = None) -> int:
    """smh_dims_t.flags of the fused step: which image of the joint distances the sweeps read.

    exact_weights=True: fp32 tiles holding the bit-exact MPJPE of the reference (0 ulp; Dmax bit-exact), i.e. the
    weights inside the fused loss are the ones the weights API materialises.
    exact_weights=False (default for the tensor-core engines with linear / mpjpe weights): SMH_DIMS_Q16_TILES, a 16-bit
    fixed-point image built from approximate square roots: |D error| <= Dbound / 130000 + ~2e-6 D, |W error| <= 1.6e-5
    (hundreds of ulp -- NOT the 1-ulp weight contract; it rides inside the 2^-11 operand rounding of the logits), Dmax
    within 4e-7 relative.  Half the workspace and tile bytes, twice the prefetch depth of the sweeps.
    None: exact_weights_default() (environment)."""
    if exact_weights is None:
        exact_weights = exact_weights_default()
    if exact_weights or engine_name == "fp32" or not neg_weighted:
        return 0
    if tuple(weighting or DEFAULT_WEIGHTING) != DEFAULT_WEIGHTING:
        return 0
    return _lib.DIMS_Q16_TILES


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"simhand_b200: `{name}` is on {t.device}; the op runs only on a CUDA sm_100 device "
            "(there is no CPU fallback)")


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float32 else t.float()


def make_inputs(z1, z2, joints1, joints2, n_local: Optional[int] = None, z_rank_stride: int = 0,
                j_rank_stride: int = 0) -> Tuple[Inputs, tuple]:
    """Describes the caller's tensors to the library without copying them: z rows may be strided,
    joints may be any strided `[N, 21, 2]` view (the reference passes `joints[:, :, :2]`)."""
    z1, z2 = _as_f32(z1), _as_f32(z2)
    joints1, joints2 = _as_f32(joints1), _as_f32(joints2)
    if z1.dim() != 2 or z1.shape != z2.shape:
        raise ValueError(f"z1/z2 must be [N, d] with equal shapes, got {tuple(z1.shape)} / {tuple(z2.shape)}")
    if joints1.shape != joints2.shape or joints1.dim() != 3 or joints1.shape[1:] != (21, 2):
        raise ValueError(f"joints must be [N, 21, 2], got {tuple(joints1.shape)} / {tuple(joints2.shape)}")
    if z1.stride(1) != 1 or z2.stride(1) != 1 or z1.stride(0) != z2.stride(0):
        z1, z2 = z1.contiguous(), z2.contiguous()
    if joints1.stride() != joints2.stride():
        joints1, joints2 = joints1.contiguous(), joints2.contiguous()
    n = z1.shape[0]
    inp = Inputs(z1.data_ptr(), z2.data_ptr(), z1.stride(0) if n > 1 else z1.shape[1],
                 joints1.data_ptr(), joints2.data_ptr(),
                 joints1.stride(0), joints1.stride(1), joints1.stride(2),
                 n if n_local is None else n_local, z_rank_stride, j_rank_stride)
    return inp, (z1, z2, joints1, joints2)       # keep the (possibly new) tensors alive


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def run_step(z1, z2, joints1, joints2, temperature: float = 0.5, engine: str = _DEFAULT_ENGINE,
             want_grad: bool = True, grad_scale: float = 1.0, strip_len: int = 0, return_aux: bool = False,
             pos_weighted: bool = True, neg_weighted: bool = True, weighting=None,
             exact_weights: Optional[bool] = None):
    """One fused fwd(+bwd) step on a single GPU.  Returns (loss[()], dz1, dz2[, aux]).
    pos_weighted / neg_weighted = False give the reference's neg-only / pos-only / unweighted losses
    (utils.py:468, :430, :157): the corresponding weight is 1.  weighting = make_weighting(...) selects
    weight_type linear / non_linear and diff_type mpjpe / w_abs / w_o_abs (utils.py:218-261, :304-346).
    exact_weights: see step_flags (True = bit-exact distances inside the fused step)."""
    for t, nm in ((z1, "z1"), (z2, "z2"), (joints1, "joints1"), (joints2, "joints2")):
        _require_cuda(t, nm)
    lib = _lib.load()
    dev = z1.device
    n, d = z1.shape
    engine_name = resolve_engine(engine, n)
    eng = _lib.ENGINES[engine_name]
    with torch.cuda.device(dev):
        ctx = get_context(n, d, 1, 0, dev, strip_len, step_flags(engine_name, weighting, neg_weighted, exact_weights),
                          weighting)
        lay, dims = ctx.layout, ctx.dims
        inp, keep = make_inputs(z1, z2, joints1, joints2)
        ws = torch.empty(int(lay.ws_bytes), dtype=torch.uint8, device=dev)
        st = _stream_ptr(dev)
        pd, pi = ctypes.byref(dims), ctypes.byref(inp)
        sweep_eng = eng | (0 if neg_weighted else _lib.UNIT_NEG_WEIGHTS)
        fin_flags = 0 if pos_weighted else _lib.UNIT_POS_WEIGHTS
        check(lib.smh_prep(pd, pi, ws.data_ptr(), eng, st), "smh_prep")
        if neg_weighted:                  # unweighted negatives need no distance matrix at all
            check(lib.smh_mpjpe(pd, ctx.plan_dev.data_ptr(), ws.data_ptr(), None, st), "smh_mpjpe")
        check(lib.smh_forward(pd, ctx.plan_dev.data_ptr(), ws.data_ptr(), temperature, sweep_eng, None, st), "smh_forward")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = dz2 = None
        if want_grad:
            check(lib.smh_backward(pd, ctx.plan_dev.data_ptr(), ws.data_ptr(), temperature, sweep_eng, None, st), "smh_backward")
            dz1 = torch.empty((n, d), dtype=torch.float32, device=dev)
            dz2 = torch.empty((n, d), dtype=torch.float32, device=dev)
        check(lib.smh_finalize(pd, pi, ws.data_ptr(), None, temperature, grad_scale, loss.data_ptr(),
                               dz1.data_ptr() if want_grad else None, dz2.data_ptr() if want_grad else None,
                               d, fin_flags, None, st), "smh_finalize")
        del keep
        if return_aux:
            aux = dict(ws=ws, ctx=ctx, neg=ctx.view(ws, lay.off_neg, lay.m),
                       stats=ctx.view(ws, lay.off_stats, 12, torch.int32),
                       posd=ctx.view(ws, lay.off_posd, n))
            return loss, dz1, dz2, aux
    return loss, dz1, dz2


def run_step_dense(z1, z2, pos_weights, neg_weights, temperature: float = 0.5, engine: str = _DEFAULT_ENGINE,
                   want_grad: bool = True, grad_scale: float = 1.0):
    """The reference's two-call API with MATERIALISED weights (`src/models/utils.py:391-427`, and :430 / :468 when one
    of the tensors is None = unit weights): `neg_weights` is any fp32 `[2N, 2N]` tensor (not assumed symmetric),
    `pos_weights` any `[N]` tensor.  Same sweeps as the fused path; the weight tiles are copied from the dense matrix
    instead of being computed from the joints (HBM-bound: one read of the matrix per sweep).  Single GPU.
    Returns (loss[()], dz1, dz2)."""
    _require_cuda(z1, "z1")
    _require_cuda(z2, "z2")
    lib = _lib.load()
    dev = z1.device
    n, d = z1.shape
    m = 2 * n
    if neg_weights is not None:
        _require_cuda(neg_weights, "neg_weights")
        if tuple(neg_weights.shape) != (m, m):
            raise ValueError(f"neg_weights must be [{m}, {m}], got {tuple(neg_weights.shape)}")
        neg_weights = _as_f32(neg_weights)
        if neg_weights.stride(1) != 1:
            neg_weights = neg_weights.contiguous()
    if pos_weights is not None:
        _require_cuda(pos_weights, "pos_weights")
        if pos_weights.numel() != n:
            raise ValueError(f"pos_weights must have {n} elements, got {tuple(pos_weights.shape)}")
        pos_weights = _as_f32(pos_weights).reshape(n).contiguous()
    eng = _lib.ENGINES[resolve_engine(engine, n)]
    with torch.cuda.device(dev):
        dense = neg_weights is not None
        fwd = get_context(n, d, 1, 0, dev, 0, _lib.DIMS_DENSE_WEIGHTS if dense else 0)
        bwd = get_context(n, d, 1, 0, dev, 0, _lib.DIMS_DENSE_WEIGHTS | _lib.DIMS_DENSE_BACKWARD) if dense else fwd
        lay = fwd.layout
        if dense and int(bwd.layout.ws_bytes) != int(lay.ws_bytes):
            raise RuntimeError("simhand_b200: forward/backward dense layouts differ")
        z1c, z2c = _as_f32(z1), _as_f32(z2)
        if z1c.stride(1) != 1 or z2c.stride(1) != 1 or z1c.stride(0) != z2c.stride(0):
            z1c, z2c = z1c.contiguous(), z2c.contiguous()
        inp = Inputs(z1c.data_ptr(), z2c.data_ptr(), z1c.stride(0) if n > 1 else d, None, None, 0, 0, 0, n, 0, 0)
        if not dense:
            # unit negatives on the fused plan: the sweeps never read the tiles, joints are not needed either
            zero = torch.zeros((n, 21, 2), dtype=torch.float32, device=dev)
            inp, _keep = make_inputs(z1c, z2c, zero, zero)
        ws = torch.empty(int(lay.ws_bytes), dtype=torch.uint8, device=dev)
        st = _stream_ptr(dev)
        pf, pb, pi = ctypes.byref(fwd.dims), ctypes.byref(bwd.dims), ctypes.byref(inp)
        sweep_eng = eng | (_lib.DENSE_WEIGHTS if dense else _lib.UNIT_NEG_WEIGHTS)
        fin_flags = _lib.DENSE_WEIGHTS if pos_weights is not None else _lib.UNIT_POS_WEIGHTS
        check(lib.smh_prep(pf, pi, ws.data_ptr(), eng, st), "smh_prep")
        if dense or pos_weights is not None:
            check(lib.smh_import_weights(pf, fwd.plan_dev.data_ptr(), ws.data_ptr(),
                                         neg_weights.data_ptr() if dense else None,
                                         neg_weights.stride(0) if dense else 0,
                                         pos_weights.data_ptr() if pos_weights is not None else None, st),
                  "smh_import_weights")
        check(lib.smh_forward(pf, fwd.plan_dev.data_ptr(), ws.data_ptr(), temperature, sweep_eng, None, st), "smh_forward")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = dz2 = None
        if want_grad:
            check(lib.smh_backward(pb, bwd.plan_dev.data_ptr(), ws.data_ptr(), temperature, sweep_eng, None, st),
                  "smh_backward")
            dz1 = torch.empty((n, d), dtype=torch.float32, device=dev)
            dz2 = torch.empty((n, d), dtype=torch.float32, device=dev)
        check(lib.smh_finalize(pf, pi, ws.data_ptr(), None, temperature, grad_scale, loss.data_ptr(),
                               dz1.data_ptr() if want_grad else None, dz2.data_ptr() if want_grad else None,
                               d, fin_flags, None, st), "smh_finalize")
    return loss, dz1, dz2


def _scale_saved_grads(ctx, grad_out):
    """backward() of the fused losses: the gradients were computed with the forward; one launch scales both by the
    upstream gradient of the 0-dim loss (which stays on the device)."""
    dz1, dz2 = ctx.saved_tensors
    if not (ctx.needs_input_grad[0] and ctx.needs_input_grad[1]) or not dz1.is_contiguous() or not dz2.is_contiguous():
        return (grad_out * dz1 if ctx.needs_input_grad[0] else None,
                grad_out * dz2 if ctx.needs_input_grad[1] else None)
    lib = _lib.load()
    g = grad_out.reshape(1).to(dtype=torch.float32)
    o1, o2 = torch.empty_like(dz1), torch.empty_like(dz2)
    with torch.cuda.device(dz1.device):
        check(lib.smh_scale_grads(dz1.data_ptr(), dz2.data_ptr(), g.data_ptr(), o1.data_ptr(), o2.data_ptr(),
                                  dz1.numel(), _stream_ptr(dz1.device)), "smh_scale_grads")
    return o1, o2


class _DenseWeightedNTXentFn(torch.autograd.Function):
    """Weighted NT-Xent with materialised weight tensors (no gradient flows to the weights, as in the reference,
    where they are built from the joints outside the graph)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, z1, z2, pos_weights, neg_weights, temperature, engine):
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        loss, dz1, dz2 = run_step_dense(z1, z2, pos_weights, neg_weights, temperature, engine, want)
        if want:
            ctx.save_for_backward(dz1, dz2)
        return loss

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        g1, g2 = _scale_saved_grads(ctx, grad_out)
        return g1, g2, None, None, None, None


class _WeightedNTXentFn(torch.autograd.Function):
    """loss = weighted NT-Xent(z1, z2 | joints1, joints2); the backward sweep runs inside forward (the
    gradient w.r.t. z is what a training step needs), backward() only scales the saved gradients."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, z1, z2, joints1, joints2, temperature, engine, group, pos_weighted=True, neg_weighted=True,
                weighting=None, exact_weights=None, grad_scale=1.0):
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if group is None:
            loss, dz1, dz2 = run_step(z1, z2, joints1, joints2, temperature, engine, want, grad_scale,
                                      pos_weighted=pos_weighted, neg_weighted=neg_weighted, weighting=weighting,
                                      exact_weights=exact_weights)
        else:
            from .dist import run_step_sharded
            loss, dz1, dz2 = run_step_sharded(z1, z2, joints1, joints2, temperature, engine, want, group, grad_scale,
                                              pos_weighted=pos_weighted, neg_weighted=neg_weighted, weighting=weighting,
                                              exact_weights=exact_weights)
        if want:
            ctx.save_for_backward(dz1, dz2)
        return loss

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        g1, g2 = _scale_saved_grads(ctx, grad_out)
        return g1, g2, None, None, None, None, None, None, None, None, None, None


def weighted_ntxent(z1: torch.Tensor, z2: torch.Tensor, joints1: torch.Tensor, joints2: torch.Tensor,
                    temperature: float = 0.5, group=None, engine: str = _DEFAULT_ENGINE,
                    pos_weighted: bool = True, neg_weighted: bool = True, weighting=None,
                    exact_weights: Optional[bool] = None, grad_scale: float = 1.0) -> torch.Tensor:
    """Fused similarity-weighted NT-Xent (weight_type linear, diff_type mpjpe, pos_neg): equals
    `vanila_weights_contrastive_loss(z1, z2, *get_weights_linear(joints1, joints2, 'mpjpe'), temperature)`
    of the reference.  With `group` (a torch.distributed process group) the batch is the concatenation of
    every rank's local batch and the work is sharded over the ranks: every rank gets the GLOBAL loss and
    d(global loss)/d(its local z).  `grad_scale` multiplies the gradients only: DistributedDataParallel averages
    parameter gradients over the ranks, so pass grad_scale=world_size under DDP to obtain the gradient of the
    global-batch loss.  exact_weights: see ops.step_flags."""
    return _WeightedNTXentFn.apply(z1, z2, joints1, joints2, float(temperature), engine, group,
                                   bool(pos_weighted), bool(neg_weighted), weighting, exact_weights, float(grad_scale))


# ----------------------------------------------------------------------------------------------------
# the reference's two-call API
# ----------------------------------------------------------------------------------------------------
class _WeightSource:
    def __init__(self, joints1, joints2, weighting=None):
        self.joints1, self.joints2 = joints1, joints2
        self.weighting = tuple(weighting or DEFAULT_WEIGHTING)
        self._dense = None

    def dense(self):
        if self._dense is None:
            self._dense = mpjpe_weights(self.joints1, self.joints2, weighting=self.weighting)
        return self._dense


class LazyWeights:
    """Stand-in for one of the two tensors `get_weights_linear` returns.  Behaves like the tensor on
    demand (`materialize()`, attribute access), but the fused loss never needs the values."""

    def __init__(self, source: _WeightSource, kind: str):
        self._source, self.kind = source, kind

    def materialize(self) -> torch.Tensor:
        pos_w, neg_w = self._source.dense()
        return pos_w if self.kind == "pos" else neg_w

    @property
    def shape(self):
        n = self._source.joints1.shape[0]
        return torch.Size([n]) if self.kind == "pos" else torch.Size([2 * n, 2 * n])

    @property
    def device(self):
        return self._source.joints1.device

    @property
    def dtype(self):
        return torch.float32

    def dim(self):
        return len(self.shape)

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def __getattr__(self, name):
        # A handle is not a tensor: building the [2N, 2N] matrix (1 GiB at 2N = 16384) is never done behind the
        # caller's back.  Tensor code that needs the values calls .materialize().
        if name.startswith("__"):
            raise AttributeError(name)
        raise AttributeError(f"LazyWeights has no attribute {name!r}: it stands for the {self.kind} weights of "
                             f"get_weights_*; call .materialize() for the real tensor of shape {tuple(self.shape)}")

    def __repr__(self):
        return f"LazyWeights(kind={self.kind!r}, shape={tuple(self.shape)})"


def mpjpe_weights(joints1: torch.Tensor, joints2: torch.Tensor, strip_len: int = 0, weighting=None):
    """Materialised (pos_w [N], neg_w [2N, 2N]) exactly as `get_weights_linear(j1, j2, 'mpjpe')` returns
    them, from the CUDA kernels (prep -> MPJPE tiles -> dense expansion); `weighting` selects the other
    diff_type / weight_type combinations of the reference."""
    _require_cuda(joints1, "joints1")
    _require_cuda(joints2, "joints2")
    lib = _lib.load()
    dev = joints1.device
    n = joints1.shape[0]
    with torch.cuda.device(dev):
        ctx = get_context(n, 1, 1, 0, dev, strip_len, 0, weighting)
        zero = torch.zeros((n, 1), dtype=torch.float32, device=dev)
        inp, keep = make_inputs(zero, zero, joints1, joints2)
        ws = torch.empty(int(ctx.layout.ws_bytes), dtype=torch.uint8, device=dev)
        st = _stream_ptr(dev)
        pd = ctypes.byref(ctx.dims)
        check(lib.smh_prep(pd, ctypes.byref(inp), ws.data_ptr(), _lib.ENGINE_FP32, st), "smh_prep")
        check(lib.smh_mpjpe(pd, ctx.plan_dev.data_ptr(), ws.data_ptr(), None, st), "smh_mpjpe")
        pos_w = torch.empty(n, dtype=torch.float32, device=dev)
        neg_w = torch.empty((2 * n, 2 * n), dtype=torch.float32, device=dev)
        check(lib.smh_weights_dense(pd, ctx.plan_dev.data_ptr(), ws.data_ptr(), pos_w.data_ptr(),
                                    neg_w.data_ptr(), st), "smh_weights_dense")
        del keep
    return pos_w, neg_w


def get_weights_linear(joints1: torch.Tensor, joints2: torch.Tensor, diff_type: str):
    """Drop-in for `src/models/utils.py:218`: `diff_type` 'mpjpe' (the hot path; bit-exact weights), 'w_abs' or
    'w_o_abs'.  Returns `(pos_weights, neg_weights)` as lazy handles."""
    src = _WeightSource(joints1, joints2, make_weighting("linear", diff_type))
    return LazyWeights(src, "pos"), LazyWeights(src, "neg")


def get_weights_nonlinear(joints1: torch.Tensor, joints2: torch.Tensor, lambda_pos: float, lambda_neg: float,
                          diff_type: str):
    """Drop-in for `src/models/utils.py:304` (`weight_type == 'non_linear'`): W = 1 / (1 + exp(lambda (D - mean D))).
    Returns `(pos_weights, neg_weights)` as lazy handles."""
    src = _WeightSource(joints1, joints2, make_weighting("non_linear", diff_type, lambda_pos, lambda_neg))
    return LazyWeights(src, "pos"), LazyWeights(src, "neg")


def apply_pca(joints: torch.Tensor, target_dim: int = 14) -> torch.Tensor:
    """Drop-in for `src/models/utils.py:192-215`: `[B, 21, 2]` joints -> `[B, target_dim]` coordinates in the basis of
    `torch.pca_lowrank` (a randomised library routine, as in the reference).  Unlike the reference it stays on the joints'
    device (the reference moves the batch to the CPU and back every step)."""
    if joints.dim() != 3 or tuple(joints.shape[1:]) != (21, 2):
        raise ValueError(f"Expected joints to have shape (batch, 21, 2), but got {tuple(joints.shape)}")
    flat = joints.contiguous().view(joints.shape[0], -1).float()
    _, _, v = torch.pca_lowrank(flat, q=target_dim)
    return torch.matmul(flat, v[:, :target_dim])


def _pca_source(joints1: torch.Tensor, joints2: torch.Tensor, weight_type: str, lambda_pos: float, lambda_neg: float,
                diff_type: str) -> "_WeightSource":
    if diff_type not in ("mpjpe", "w_abs", "w_o_abs"):
        raise ValueError(f"diff_type must be mpjpe, w_abs or w_o_abs, got {diff_type!r}")
    if joints1.dim() != 2 or joints1.shape != joints2.shape or joints1.shape[1] > 42:
        raise ValueError(f"PCA coordinates must be [N, K <= 42] with equal shapes, got {tuple(joints1.shape)} / "
                         f"{tuple(joints2.shape)}")
    k = joints1.shape[1]
    pad = lambda t: torch.nn.functional.pad(_as_f32(t), (0, 42 - k)).view(t.shape[0], 21, 2)      # noqa: E731
    # all three diff_types are the Euclidean distance between the coordinate vectors there (utils.py:265-293):
    # || |a - b| || == || a - b ||
    return _WeightSource(pad(joints1), pad(joints2), make_weighting(weight_type, "pca", lambda_pos, lambda_neg))


def get_weights_linear_with_pca(joints1: torch.Tensor, joints2: torch.Tensor, diff_type: str):
    """Drop-in for `src/models/utils.py:264-301` (`config.use_pca`): linear weights from the Euclidean distance between
    `[N, K]` PCA coordinates (`apply_pca`).  Returns lazy handles, as `get_weights_linear`."""
    src = _pca_source(joints1, joints2, "linear", 0.0, 0.0, diff_type)
    return LazyWeights(src, "pos"), LazyWeights(src, "neg")


def get_weights_nonlinear_with_pca(joints1: torch.Tensor, joints2: torch.Tensor, lambda_pos: float, lambda_neg: float,
                                   diff_type: str):
    """Drop-in for `src/models/utils.py:349-388`: sigmoid weights from the same distance."""
    src = _pca_source(joints1, joints2, "non_linear", lambda_pos, lambda_neg, diff_type)
    return LazyWeights(src, "pos"), LazyWeights(src, "neg")


def vanila_weights_contrastive_loss(z1: torch.Tensor, z2: torch.Tensor, pos_weights, neg_weights,
                                    temperature: float = 0.5, engine: str = _DEFAULT_ENGINE,
                                    exact_weights: Optional[bool] = None) -> torch.Tensor:
    """Drop-in for `src/models/utils.py:391`: mean-reduced weighted NT-Xent, differentiable in z1, z2."""
    if isinstance(pos_weights, LazyWeights) and isinstance(neg_weights, LazyWeights):
        if pos_weights._source is not neg_weights._source or pos_weights.kind != "pos" or neg_weights.kind != "neg":
            raise ValueError("pos_weights / neg_weights must come from the same get_weights_linear call")
        src = pos_weights._source
        if src.joints1.shape[0] != z1.shape[0]:
            raise ValueError(f"weights were built for batch {src.joints1.shape[0]}, z1 has {z1.shape[0]}")
        return weighted_ntxent(z1, z2, src.joints1, src.joints2, temperature, None, engine, weighting=src.weighting,
                               exact_weights=exact_weights)
    if isinstance(pos_weights, LazyWeights):
        pos_weights = pos_weights.materialize()
    if isinstance(neg_weights, LazyWeights):
        neg_weights = neg_weights.materialize()
    # real tensors (any values): the materialised-weights path
    return _DenseWeightedNTXentFn.apply(z1, z2, pos_weights, neg_weights, float(temperature), engine)


def _source_of(weights, kind: str):
    """The joints behind a lazy handle of the right kind, or None for a real tensor."""
    if isinstance(weights, LazyWeights):
        if weights.kind != kind:
            raise ValueError(f"simhand_b200: expected the {kind}-weights handle, got the {weights.kind} one")
        return weights._source
    if not isinstance(weights, torch.Tensor):
        raise TypeError(f"simhand_b200: {kind}_weights must be a tensor or a handle from get_weights_linear")
    return None


def vanila_pos_weights_contrastive_loss(z1, z2, pos_weights, temperature: float = 0.5,
                                        engine: str = _DEFAULT_ENGINE) -> torch.Tensor:
    """Drop-in for `src/models/utils.py:430` (`pos_neg == "pos"`): only the positive logits are weighted."""
    src = _source_of(pos_weights, "pos")
    if src is None:
        return _DenseWeightedNTXentFn.apply(z1, z2, pos_weights, None, float(temperature), engine)
    return weighted_ntxent(z1, z2, src.joints1, src.joints2, temperature, None, engine, True, False, src.weighting)


def vanila_neg_weights_contrastive_loss(z1, z2, neg_weights, temperature: float = 0.5,
                                        engine: str = _DEFAULT_ENGINE) -> torch.Tensor:
    """Drop-in for `src/models/utils.py:468` (`pos_neg == "neg"`): only the negative logits are weighted."""
    src = _source_of(neg_weights, "neg")
    if src is None:
        return _DenseWeightedNTXentFn.apply(z1, z2, None, neg_weights, float(temperature), engine)
    return weighted_ntxent(z1, z2, src.joints1, src.joints2, temperature, None, engine, False, True, src.weighting)


def vanila_contrastive_loss(z1, z2, temperature: float = 0.5, engine: str = _DEFAULT_ENGINE) -> torch.Tensor:
    """Drop-in for `src/models/utils.py:157`: plain NT-Xent (both weights 1), same fused sweeps."""
    zero = torch.zeros((z1.shape[0], 21, 2), dtype=torch.float32, device=z1.device)
    return weighted_ntxent(z1, z2, zero, zero, temperature, None, engine, False, False)


_DROP_INS = ("get_weights_linear", "get_weights_nonlinear", "vanila_weights_contrastive_loss", "vanila_pos_weights_contrastive_loss",
             "vanila_neg_weights_contrastive_loss", "vanila_contrastive_loss", "apply_pca", "get_weights_linear_with_pca",
             "get_weights_nonlinear_with_pca")


def install(*modules) -> None:
    """Rebinds `get_weights_linear` / `vanila_weights_contrastive_loss` in the given modules (the reference's
    `src.models.utils` and the model modules that imported the names: simhand_w_model, peclr_w_model,
    simclr_w_model).  See INTEGRATION.md."""
    for mod in modules:
        for name in _DROP_INS:
            if hasattr(mod, name):
                setattr(mod, name, globals()[name])


# ----------------------------------------------------------------------------------------------------
# K3: L2 normalisation
# ----------------------------------------------------------------------------------------------------
class _L2NormFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, eps):
        _require_cuda(x, "x")
        lib = _lib.load()
        x = x.contiguous()
        rows, d = x.shape
        y = torch.empty_like(x)
        norm = torch.empty(rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.smh_l2norm_fwd(x.data_ptr(), y.data_ptr(), norm.data_ptr(), rows, d, eps,
                                     _stream_ptr(x.device)), "smh_l2norm_fwd")
        ctx.save_for_backward(y, norm)
        ctx.eps = eps
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        y, norm = ctx.saved_tensors
        lib = _lib.load()
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        with torch.cuda.device(y.device):
            check(lib.smh_l2norm_bwd(y.data_ptr(), norm.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.shape[0],
                                     y.shape[1], ctx.eps, _stream_ptr(y.device)), "smh_l2norm_bwd")
        return dx, None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """`torch.nn.functional.normalize(x, dim=1)` for `[B, d]` fp32 (simhand_w_model.py:56-58, 91-93)."""
    return _L2NormFn.apply(x, float(eps))


# ----------------------------------------------------------------------------------------------------
# K4: fused projection-space transform
# ----------------------------------------------------------------------------------------------------
class _TransformFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, tx, ty, angle, eps):
        _require_cuda(x, "projections")
        lib = _lib.load()
        rows, d = x.shape
        xc = x.contiguous()
        vec = lambda t, nm: None if t is None else _vec_on(t, rows, x.device, nm)        # noqa: E731
        txc, tyc, anc = vec(tx, "translate_x"), vec(ty, "translate_y"), vec(angle, "angle")
        out = torch.empty_like(xc)
        save = torch.empty((rows, 4), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.smh_transform_fwd(xc.data_ptr(), d, txc.data_ptr() if txc is not None else None,
                                        tyc.data_ptr() if tyc is not None else None,
                                        anc.data_ptr() if anc is not None else None, out.data_ptr(), d,
                                        save.data_ptr(), rows, d, eps, _stream_ptr(x.device)), "smh_transform_fwd")
        ctx.save_for_backward(xc, out, save)
        ctx.eps = eps
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        xc, out, save = ctx.saved_tensors
        lib = _lib.load()
        rows, d = xc.shape
        gc = g.contiguous().float()
        dx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            check(lib.smh_transform_bwd(xc.data_ptr(), d, out.data_ptr(), d, save.data_ptr(), gc.data_ptr(), d,
                                        dx.data_ptr(), d, rows, d, ctx.eps, _stream_ptr(xc.device)),
                  "smh_transform_bwd")
        return dx, None, None, None, None


def _vec_on(t: torch.Tensor, rows: int, device, name: str) -> torch.Tensor:
    t = torch.as_tensor(t)
    if t.numel() != rows:
        raise ValueError(f"{name} must have one value per projection row ({rows}), got {tuple(t.shape)}")
    return t.detach().to(device=device, dtype=torch.float32).reshape(rows).contiguous()


def get_transformed_projections(projections: torch.Tensor, translate_x=None, translate_y=None, angle=None,
                                eps: float = 1e-12) -> torch.Tensor:
    """The projection-space equivariance step of HandCLR_W / PeCLR_W (`simhand_w_model.py:55-94`,
    `peclr_w_model.py:52-91`) in one kernel each way:
        normalize -> translate_encodings(., translate_x, translate_y) -> rotate_encoding(., angle) -> normalize
    on `[2B, d]` raw projections seen as d/2 2-D points per row (`utils.py:636-684`).  `translate_*` / `angle` are the
    values the reference hands to its helpers (the models pass `-jitter / image_size` and `-angles`); None skips that
    stage ("crop" / "rotate" not in `config.augmentation`).  The per-row extent and centroid are detached, as in the
    reference.  Returns `[2B, d]`; split it in halves for (projection1, projection2)."""
    if projections.dim() != 2 or projections.shape[1] % 2 or projections.shape[1] > 128:
        raise ValueError(f"projections must be [rows, d] with even d <= 128, got {tuple(projections.shape)}")
    if (translate_x is None) != (translate_y is None):
        raise ValueError("translate_x and translate_y go together")
    return _TransformFn.apply(projections, translate_x, translate_y, angle, float(eps))
