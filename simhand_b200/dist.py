"""Sharded weighted NT-Xent over the GPUs of one box (BASELINE.json configs[2]).

One process per GPU (`torch.distributed`, NCCL over NVLink).  Every rank holds `n_local` samples of each
view; the loss is the reference's loss on the concatenation of all local batches (global 2N samples).

Three transports (run_step_sharded): "fused" (default: run_step_fused, the exchange rides in the kernels' heads and tails,
csrc/smh_shard.cu), "peer" (the same exchange as separate push / barrier kernels) and "nccl".  EmulatedGroup runs all ranks
of the fused step on one device (tests).  The collective pattern, spelled out for the NCCL transport --

Per step and rank:
  1. all-gather of one packed buffer [z1 | z2 | joints1 | joints2] (the autograd transpose of step 5)
  2. smh_prep on the gathered batch, smh_mpjpe on this rank's share of the upper-triangular tiles
  3. all-reduce(MAX) of three integers (Dmax, Pmax, -Pmin images)                 utils.py:233-234, :255
  4. forward sweep over this rank's tasks -> partial row sums; all-reduce(SUM) of neg [2N]
  5. backward sweep -> partial dz for all rows; reduce-scatter(SUM) -> dz of the local samples
  6. smh_finalize: loss (identical on every rank) and the local gradients
The MPJPE work is split by unordered tile pairs (each rank evaluates 1/P of the upper triangle), which is
why the row sums and the gradient contributions of a rank touch all rows and the exchange steps 4 and 5 are
real collectives rather than bookkeeping.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .ops import DEFAULT_WEIGHTING, _as_f32, _stream_ptr, get_context, make_inputs, resolve_engine, step_flags


def pack_local(z1, z2, joints1, joints2) -> torch.Tensor:
    """[z1 | z2 | joints1 | joints2] of this rank as one flat fp32 buffer (joints as contiguous [n,21,2])."""
    return torch.cat([_as_f32(z1).reshape(-1), _as_f32(z2).reshape(-1),
                      _as_f32(joints1).contiguous().reshape(-1), _as_f32(joints2).contiguous().reshape(-1)])


def gathered_views(gathered: torch.Tensor, world: int, n_local: int, d: int):
    """Strided views of the all-gathered buffer with the reference's shapes: z1, z2 [N, d] and joints
    [N, 21, 2] are not contiguous across ranks, so they are described by (rank_stride, row_stride);
    returns the four base offsets (elements) and the chunk length."""
    chunk = 2 * n_local * d + 2 * n_local * 42
    assert gathered.numel() == world * chunk
    off_z1, off_z2 = 0, n_local * d
    off_j1, off_j2 = 2 * n_local * d, 2 * n_local * d + n_local * 42
    return (off_z1, off_z2, off_j1, off_j2), chunk


def sample_offset(k: int, n_local: int, rank_stride: int, row_stride: int) -> int:
    """Element offset of global sample k inside a gathered segment (mirrors smh_inputs_t)."""
    return (k // n_local) * rank_stride + (k % n_local) * row_stride


def dz_out_row(i: int, n: int, n_local: int) -> int:
    """Row of global sample row i (= v*N + k) in the rank-major gradient accumulator (smh_common.cuh)."""
    v = 1 if i >= n else 0
    k = i - v * n
    return (k // n_local) * 2 * n_local + v * n_local + k % n_local


DEFAULT_TIMEOUT_MS = 30000


def exchange_timeout_ms() -> int:
    """Bound of every cross-rank wait on the device (SMH_EXCHANGE_TIMEOUT_MS; a rank that stays away longer -- a
    checkpoint, a stalled dataloader -- poisons the group: the loss becomes NaN instead of a silently wrong value)."""
    return int(os.environ.get("SMH_EXCHANGE_TIMEOUT_MS", DEFAULT_TIMEOUT_MS))


class PeerExchange:
    """Symmetric buffers of one sharded problem shape, mapped into every rank (torch.distributed._symmetric_memory):
    the workspace blob (peers store operand images / row sums / gradient rows / Dmax into it over NVLink), for the
    unfused transport the gathered-input buffer (peers push their packed inputs into it), and the signal words.
    Built once per (shape, group) and reused."""

    def __init__(self, ctx, group, chunk_floats: int, fused: bool):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = ctx.device
        lay = ctx.layout
        # all ranks must allocate the same size: the MPJPE tile region differs by a few tiles between ranks
        ws_bytes = torch.tensor([int(lay.ws_bytes)], dtype=torch.int64, device=dev)
        dist.all_reduce(ws_bytes, op=dist.ReduceOp.MAX, group=group)
        self.ws = symm.empty(int(ws_bytes.item()), dtype=torch.uint8, device=dev)
        self.ws.zero_()                    # fused: padding rows of the images and both slots of the per-step scalars
        self.xin = None if fused else symm.empty(self.world * chunk_floats, dtype=torch.float32, device=dev)
        self.signal = symm.empty(_lib.SIGNAL_WORDS, dtype=torch.int32, device=dev)
        self.signal.zero_()
        self.h_ws = symm.rendezvous(self.ws, group)
        self.h_xin = None if fused else symm.rendezvous(self.xin, group)
        self.h_sig = symm.rendezvous(self.signal, group)
        ex = _lib.Exchange()
        ex.world, ex.rank = self.world, self.rank
        for p in range(self.world):
            ex.ws_peer[p] = self.h_ws.buffer_ptrs[p]
            ex.xin_peer[p] = None if fused else self.h_xin.buffer_ptrs[p]
            ex.signal_peer[p] = self.h_sig.buffer_ptrs[p]
        ex.fused = 1 if fused else 0
        ex.timeout_ms = exchange_timeout_ms()
        assert ex.ws_peer[self.rank] == self.ws.data_ptr()
        self.struct = ex
        self.fused = fused
        torch.cuda.synchronize(dev)
        dist.barrier(group)                # every rank's signal words are zero before anyone signals

    def poisoned(self) -> int:
        """Site of the first cross-rank wait that timed out anywhere in the group (0 = healthy).  Synchronises."""
        return int(self.signal[_lib.SIG_POISON].item())


_exchanges = {}


def get_exchange(ctx, group, chunk_floats: int, fused: bool = False) -> PeerExchange:
    key = (id(ctx), id(group), chunk_floats, fused)
    ex = _exchanges.get(key)
    if ex is None:
        ex = PeerExchange(ctx, group, chunk_floats, fused)
        _exchanges[key] = ex
    return ex


def peer_exchange_available() -> bool:
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401
        return True
    except Exception:
        return False


def run_step_peer(z1, z2, joints1, joints2, temperature: float, engine: str, want_grad: bool, group,
                  grad_scale: float = 1.0, strip_len: int = 0, exact_weights: Optional[bool] = None):
    """Sharded step with the collectives done by the library over peer memory (NVLink): push-gather of the inputs,
    Dmax pushed by the MPJPE kernel's last CTA, partial row sums / gradient rows stored into the peers' partial
    buffers and reduced in rank order by their consumers; device-side barriers separate the phases.  No NCCL call
    on the data path; the whole step is a fixed kernel sequence (CUDA-graph capturable).  The cross-rank sums are taken in
    rank order (every rank returns the same loss bits); inside a rank the sweeps' strip flushes use floating-point atomics,
    so a step is not bitwise reproducible from run to run (differences at the 1e-7 level)."""
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = z1.device
    n_local, d = z1.shape
    n = n_local * world
    engine_name = resolve_engine(engine, n)
    eng = _lib.ENGINES[engine_name]
    with torch.cuda.device(dev):
        ctx = get_context(n, d, world, rank, dev, strip_len, step_flags(engine_name, exact_weights=exact_weights))
        lay, dims = ctx.layout, ctx.dims
        chunk = 2 * n_local * (d + 42)
        ex = get_exchange(ctx, group, chunk)
        local_in, keep = make_inputs(z1, z2, joints1, joints2)
        px = ctypes.byref(ex.struct)
        ws = ex.ws
        (o1, o2, oj1, oj2), _ = gathered_views(ex.xin, world, n_local, d)
        base = ex.xin.data_ptr()
        inp = _lib.Inputs(base + 4 * o1, base + 4 * o2, d, base + 4 * oj1, base + 4 * oj2, 42, 2, 1,
                          n_local, chunk, chunk)
        st = _stream_ptr(dev)
        pd, pi = ctypes.byref(dims), ctypes.byref(inp)
        plan = ctx.plan_dev.data_ptr()
        # The accumulators this rank owns are zeroed by smh_prep; peers may only add to them after the first barrier.
        check(lib.smh_push_inputs(px, ctypes.byref(local_in), n_local, d, st), "smh_push_inputs")
        del keep
        check(lib.smh_prep_zero(pd, ws.data_ptr(), st), "smh_prep_zero")
        check(lib.smh_barrier(px, st), "smh_barrier")
        check(lib.smh_prep(pd, pi, ws.data_ptr(), eng | _lib.PREP_NO_ZERO, st), "smh_prep")
        check(lib.smh_mpjpe(pd, plan, ws.data_ptr(), px, st), "smh_mpjpe")
        check(lib.smh_barrier(px, st), "smh_barrier")
        check(lib.smh_forward(pd, plan, ws.data_ptr(), temperature, eng, px, st), "smh_forward")
        check(lib.smh_exchange_neg(pd, ws.data_ptr(), px, st), "smh_exchange_neg")
        check(lib.smh_barrier(px, st), "smh_barrier")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = dz2 = None
        fin = lambda flags: check(lib.smh_finalize(pd, pi, ws.data_ptr(), None, temperature, grad_scale,      # noqa: E731
                                                   loss.data_ptr(), dz1.data_ptr() if want_grad else None,
                                                   dz2.data_ptr() if want_grad else None, d, flags, px, st),
                                  "smh_finalize")
        if want_grad:
            dz1 = torch.empty((n_local, d), dtype=torch.float32, device=dev)
            dz2 = torch.empty((n_local, d), dtype=torch.float32, device=dev)
            check(lib.smh_backward(pd, plan, ws.data_ptr(), temperature, eng, px, st), "smh_backward")
            # sharded finalize, first half: loss terms of the own rows, partial sum stored into every peer
            fin(_lib.FINALIZE_LOSS_PART)
            check(lib.smh_exchange_dz(pd, ws.data_ptr(), px, st), "smh_exchange_dz")
        else:
            # the loss needs the summed row sums: same reduction the backward's first kernel would do
            check(lib.smh_backward(pd, plan, ws.data_ptr(), temperature, eng | _lib.BACKWARD_RN_ONLY, px, st), "smh_rn")
            fin(_lib.FINALIZE_LOSS_PART)
        check(lib.smh_barrier(px, st), "smh_barrier")
        # second half: gradients of the own rows, loss = rank-ordered sum of the partials.  No closing barrier: after
        # smh_prep nothing reads another rank's slot of the gathered inputs, and every cross-rank store of this step
        # precedes the barrier above, so the next step's push cannot race with this one.
        fin(_lib.FINALIZE_GRAD)
    return loss, dz1, dz2


_side_streams = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device)
    return _side_streams[key]


def _fused_launches(lib, ctx, exch_struct, ws_ptr, local, temperature, eng, want_grad, grad_scale, outs, st,
                    pos_weighted=True, neg_weighted=True, stages=None, overlap_push: bool = True):
    """The launches of one rank's fused step (smh_shard.cu).  `stages`: subset to issue (the single-GPU emulation runs the
    ranks stage by stage, on one stream).  overlap_push: the z images travel on a second stream next to the MPJPE kernel
    (a parallel branch when the step is captured into a CUDA graph)."""
    dims = ctx.dims
    pd, pl, px = ctypes.byref(dims), ctypes.byref(local), ctypes.byref(exch_struct)
    plan = ctx.plan_dev.data_ptr()
    loss, dz1, dz2 = outs
    d = dims.d
    sweep_eng = eng | (0 if neg_weighted else _lib.UNIT_NEG_WEIGHTS)
    fin_flags = 0 if pos_weighted else _lib.UNIT_POS_WEIGHTS
    table = {
        "prep": lambda: check(lib.smh_shard_prep(pd, pl, ws_ptr, px, eng | _lib.SHARD_PREP_NO_IMAGES, st), "smh_shard_prep"),
        "zpush": lambda: check(lib.smh_shard_push_z(pd, pl, ws_ptr, px, eng, st), "smh_shard_push_z"),
        "mpjpe": lambda: check(lib.smh_mpjpe(pd, plan, ws_ptr, px, st), "smh_mpjpe"),
        "fwd": lambda: check(lib.smh_forward(pd, plan, ws_ptr, temperature, sweep_eng, px, st), "smh_forward"),
        "bwd": lambda: check(lib.smh_backward(pd, plan, ws_ptr, temperature,
                                              sweep_eng | (0 if want_grad else _lib.BACKWARD_RN_ONLY), px, st), "smh_backward"),
        "fin": lambda: check(lib.smh_finalize(pd, pl, ws_ptr, None, temperature, grad_scale, loss.data_ptr(),
                                              dz1.data_ptr() if want_grad else None, dz2.data_ptr() if want_grad else None,
                                              d, fin_flags, px, st), "smh_finalize"),
    }
    if stages is None and overlap_push:
        # prep -> { z push on the side stream | MPJPE on the main stream } -> join -> forward sweep ...
        dev = ctx.device
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        table["prep"]()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            check(lib.smh_shard_push_z(pd, pl, ws_ptr, px, eng, side.cuda_stream), "smh_shard_push_z")
        table["mpjpe"]()
        main.wait_stream(side)
        for name in ("fwd", "bwd", "fin"):
            table[name]()
        return
    for name in (stages or FUSED_STAGES):
        table[name]()


FUSED_STAGES = ("prep", "zpush", "mpjpe", "fwd", "bwd", "fin")


def run_step_fused(z1, z2, joints1, joints2, temperature: float, engine: str, want_grad: bool, group,
                   grad_scale: float = 1.0, strip_len: int = 0, pos_weighted: bool = True, neg_weighted: bool = True,
                   weighting=None, exact_weights: Optional[bool] = None):
    """Sharded step with the exchange fused into the kernels (smh_exchange_t.fused; smh_shard.cu): 6 launches per rank
    and step, no barrier / copy kernels, no NCCL call on the data path; CUDA-graph capturable.  Every weighting of the
    single-GPU path is available (the non_linear mean distance and the positive terms travel with the stage signals)."""
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = z1.device
    n_local, d = z1.shape
    n = n_local * world
    engine_name = resolve_engine(engine, n)
    if engine_name == "fp32":
        raise ValueError("the fused exchange runs the tensor-core engines (fp16 / tf32 / bf16)")
    eng = _lib.ENGINES[engine_name]
    with torch.cuda.device(dev):
        ctx = get_context(n, d, world, rank, dev, strip_len, step_flags(engine_name, weighting, neg_weighted, exact_weights),
                          weighting)
        ex = get_exchange(ctx, group, 0, fused=True)
        local_in, keep = make_inputs(z1, z2, joints1, joints2)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = dz2 = None
        if want_grad:
            dz1 = torch.empty((n_local, d), dtype=torch.float32, device=dev)
            dz2 = torch.empty((n_local, d), dtype=torch.float32, device=dev)
        _fused_launches(lib, ctx, ex.struct, ex.ws.data_ptr(), local_in, temperature, eng, want_grad, grad_scale,
                        (loss, dz1, dz2), _stream_ptr(dev), pos_weighted, neg_weighted)
        del keep
    return loss, dz1, dz2


class EmulatedGroup:
    """All ranks of a fused sharded step on ONE device: `world` workspaces and signal blocks in the same memory, every
    rank's kernels launched stage by stage on one stream (rank 0 first, so no wait ever blocks).  The kernels, plans,
    layouts and the exchange protocol are exactly those of the multi-GPU run; only NVLink is missing.  Used by the
    single-GPU tests (tests/test_gpu_shard_emulation.py) and for timing one rank's kernels without an 8-GPU box."""

    def __init__(self, n: int, d: int, world: int, device, engine: str = "fp16", weighting=None, neg_weighted: bool = True,
                 exact_weights: Optional[bool] = None, strip_len: int = 0):
        if n % world:
            raise ValueError("n must be a multiple of world")
        self.n, self.d, self.world, self.device = n, d, world, torch.device(device)
        self.engine_name = resolve_engine(engine, n)
        flags = step_flags(self.engine_name, weighting, neg_weighted, exact_weights)
        self.ctxs = [get_context(n, d, world, r, self.device, strip_len, flags, weighting) for r in range(world)]
        ws_bytes = max(int(c.layout.ws_bytes) for c in self.ctxs)
        self.ws = [torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device) for _ in range(world)]
        self.signal = [torch.zeros(_lib.SIGNAL_WORDS, dtype=torch.int32, device=self.device) for _ in range(world)]
        self.structs = []
        for r in range(world):
            ex = _lib.Exchange()
            ex.world, ex.rank = world, r
            for p in range(world):
                ex.ws_peer[p] = self.ws[p].data_ptr()
                ex.xin_peer[p] = None
                ex.signal_peer[p] = self.signal[p].data_ptr()
            ex.fused = 1
            ex.timeout_ms = 2000               # a protocol bug must not hold the device for long in a test
            self.structs.append(ex)

    def step(self, z1, z2, joints1, joints2, temperature: float = 0.5, want_grad: bool = True, grad_scale: float = 1.0,
             pos_weighted: bool = True, neg_weighted: bool = True, timing: Optional[dict] = None):
        """z1 / z2 / joints: the GLOBAL batch ([n, d], [n, 21, 2] views); rank r takes samples [r n_local, (r+1) n_local).
        Returns per-rank lists (loss, dz1, dz2).  timing: dict filled with CUDA-event milliseconds per (stage, rank)."""
        lib = _lib.load()
        n_local = self.n // self.world
        eng = _lib.ENGINES[self.engine_name]
        st = _stream_ptr(self.device)
        outs, locals_, keeps = [], [], []
        for r in range(self.world):
            sl = slice(r * n_local, (r + 1) * n_local)
            li, keep = make_inputs(z1[sl], z2[sl], joints1[sl], joints2[sl])
            locals_.append(li)
            keeps.append(keep)
            loss = torch.empty((), dtype=torch.float32, device=self.device)
            g1 = torch.empty((n_local, self.d), dtype=torch.float32, device=self.device) if want_grad else None
            g2 = torch.empty((n_local, self.d), dtype=torch.float32, device=self.device) if want_grad else None
            outs.append((loss, g1, g2))
        with torch.cuda.device(self.device):
            for stage in FUSED_STAGES:
                for r in range(self.world):
                    if timing is not None:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                    _fused_launches(lib, self.ctxs[r], self.structs[r], self.ws[r].data_ptr(), locals_[r], temperature, eng,
                                    want_grad, grad_scale, outs[r], st, pos_weighted, neg_weighted, stages=(stage,))
                    if timing is not None:
                        e1.record()
                        timing.setdefault((stage, r), []).append((e0, e1))
        del keeps
        return [o[0] for o in outs], [o[1] for o in outs], [o[2] for o in outs]

    def poisoned(self):
        return [int(s[_lib.SIG_POISON].item()) for s in self.signal]


def run_step_sharded(z1, z2, joints1, joints2, temperature: float, engine: str, want_grad: bool,
                     group: Optional[dist.ProcessGroup], grad_scale: float = 1.0, strip_len: int = 0,
                     transport: str = "auto", pos_weighted: bool = True, neg_weighted: bool = True, weighting=None,
                     exact_weights: Optional[bool] = None):
    """transport: "fused" (6 launches, the exchange rides in the kernels' heads and tails over peer memory; every
    weighting), "peer" (the same exchange as separate push / barrier kernels: 14 launches; linear / mpjpe / pos_neg only),
    "nccl" (NCCL calls between the kernels) or "auto".  Both peer-memory forms were measured on 8 B200s (DESIGN.md section 7,
    profiles/r02_bench_n8_*.json): inside a CUDA graph a launch boundary costs ~1 us, both forms pay the same cross-rank
    skew at four synchronisation points and the same NVLink-bound payloads, and the unfused form came out 3-5 % ahead
    (3670 vs 3496 steps/s at 8 GPUs) -- so "auto" picks "peer" for the default weighting and "fused" for everything
    only it implements (other weightings, pos / neg only); "nccl" when symmetric memory is unavailable or the group is
    larger than SMH_MAX_PEERS.  SMH_TRANSPORT overrides "auto"."""
    n_glob = z1.shape[0] * dist.get_world_size(group)
    plain = pos_weighted and neg_weighted and tuple(weighting or DEFAULT_WEIGHTING) == DEFAULT_WEIGHTING
    if transport == "auto":
        can_peer = peer_exchange_available() and dist.get_world_size(group) <= _lib.MAX_PEERS
        tc = resolve_engine(engine, n_glob) != "fp32"
        if not can_peer:
            transport = "nccl"
        elif plain or not tc:
            transport = "peer"
        else:
            transport = "fused"
        transport = os.environ.get("SMH_TRANSPORT", transport)
    if transport == "fused":
        return run_step_fused(z1, z2, joints1, joints2, temperature, engine, want_grad, group, grad_scale, strip_len,
                              pos_weighted, neg_weighted, weighting, exact_weights)
    if not plain:
        raise NotImplementedError("only the fused transport implements the other weightings on several ranks")
    if transport == "peer":
        return run_step_peer(z1, z2, joints1, joints2, temperature, engine, want_grad, group, grad_scale, strip_len,
                             exact_weights)
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = z1.device
    n_local, d = z1.shape
    n = n_local * world
    engine_name = resolve_engine(engine, n)
    eng = _lib.ENGINES[engine_name]
    with torch.cuda.device(dev):
        ctx = get_context(n, d, world, rank, dev, strip_len, step_flags(engine_name, exact_weights=exact_weights))
        lay, dims = ctx.layout, ctx.dims
        local = pack_local(z1, z2, joints1, joints2)
        gathered = torch.empty(world * local.numel(), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, local, group=group)
        (o1, o2, oj1, oj2), chunk = gathered_views(gathered, world, n_local, d)
        base = gathered.data_ptr()
        inp = _lib.Inputs(base + 4 * o1, base + 4 * o2, d, base + 4 * oj1, base + 4 * oj2, 42, 2, 1,
                          n_local, chunk, chunk)
        ws = torch.empty(int(lay.ws_bytes), dtype=torch.uint8, device=dev)
        st = _stream_ptr(dev)
        pd, pi = ctypes.byref(dims), ctypes.byref(inp)
        plan = ctx.plan_dev.data_ptr()
        check(lib.smh_prep(pd, pi, ws.data_ptr(), eng, st), "smh_prep")
        check(lib.smh_mpjpe(pd, plan, ws.data_ptr(), None, st), "smh_mpjpe")
        stats3 = ctx.view(ws, lay.off_stats, 3, torch.int32)
        dist.all_reduce(stats3, op=dist.ReduceOp.MAX, group=group)
        check(lib.smh_forward(pd, plan, ws.data_ptr(), temperature, eng, None, st), "smh_forward")
        neg = ctx.view(ws, lay.off_neg, lay.m)
        dist.all_reduce(neg, op=dist.ReduceOp.SUM, group=group)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dz1 = dz2 = None
        dz_local = None
        if want_grad:
            check(lib.smh_backward(pd, plan, ws.data_ptr(), temperature, eng, None, st), "smh_backward")
            dzacc = ctx.view(ws, lay.off_dzacc, lay.m * 128)
            dz_local = torch.empty(2 * n_local * 128, dtype=torch.float32, device=dev)
            dist.reduce_scatter_tensor(dz_local, dzacc, op=dist.ReduceOp.SUM, group=group)
            dz1 = torch.empty((n_local, d), dtype=torch.float32, device=dev)
            dz2 = torch.empty((n_local, d), dtype=torch.float32, device=dev)
        check(lib.smh_finalize(pd, pi, ws.data_ptr(), dz_local.data_ptr() if want_grad else None, temperature,
                               grad_scale, loss.data_ptr(), dz1.data_ptr() if want_grad else None,
                               dz2.data_ptr() if want_grad else None, d, 0, None, st), "smh_finalize")
        # keep the gathered inputs alive until the stream has consumed them
        gathered.record_stream(torch.cuda.current_stream(dev))
    return loss, dz1, dz2
