#!/usr/bin/env python
"""Benchmark of the similarity-weighted NT-Xent hot path (BASELINE.json metric):

    weighted NT-Xent fwd+bwd steps/sec @ 2N = 16384, d = 128 on N B200s, with % of roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine tf32|fp32]

One "step" = everything between (z1, z2, joints1, joints2) and (loss, dz1, dz2): prep, MPJPE tiles, forward
sweep, backward sweep, finalize (+ the collectives when N > 1).  Prints ONE JSON line on rank 0.
`--impl reference` times the reference algorithm's CPU port (oracle/restate.py: the same torch ops as
src/models/utils.py:218-261, :391-427) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_VIEW = 8192
DIM = 128
TAU = 0.5
METRIC = "weighted NT-Xent fwd+bwd steps/sec @2N=16384,d=128"
UNIT = "steps/s"
MUFU_LANES_PER_SM = 16
NUM_SMS = 148


def _ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` per launch, from the newest committed ncu summary under
    profiles/ (tools/ncu_summary.py output of a `ncu --set full` capture of this same command); None if absent."""
    import glob
    import re
    # by name, newest round first (file times mean nothing after a checkout); `kernel` carries the template argument that
    # tells the 16-bit-image instantiation ("mpjpe_kernel<1,") from the exact one ("mpjpe_kernel<0,")
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_summary*.txt")))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reversed(files):
        total, inside, found = 0.0, False, 0
        for ln in open(path, errors="replace"):
            if ln.startswith("----"):
                inside = kernel in ln
            m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", ln)
            if inside and m and m.group(3) in scale:
                total += float(m.group(2)) * scale[m.group(3)]
                found += 1
        if found >= 2:
            return dict(bytes=total, source=os.path.relpath(path, ROOT))
    return None


def _mufu_peak():
    """Measured special-function throughput (tools/mufu_peak.cu, run on this pool's B200: profiles/r02_mufu_peak.json)."""
    p = os.path.join(ROOT, "profiles", "r02_mufu_peak.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return json.load(fh)
    return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)",
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0))
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)", sm_max_mhz=1965.0)


class ClockSampler:
    """Streams nvidia-smi clocks / throttle reasons of one GPU (-lms 50) while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [v for v in (num(s[0]) for s in self.samples) if v is not None]
        mx = [v for v in (num(s[1]) for s in self.samples) if v is not None]
        pw = [v for v in (num(s[2]) for s in self.samples) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[3 + i].lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, reasons=reasons, samples=len(self.samples))


def _make_batch(n):
    from simhand_b200 import synth
    return synth.make_batch(n, DIM, 5, "hand")


# --------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_port_sample(rows: int = 128, repeats: int = 3, warmup: int = 1, min_seconds: float = 0.0):
    """Times the reference's torch ops on a row block of the 2N = 16384 problem and extrapolates to a step.
    The reference itself cannot run this size on a host (63 GiB of temporaries)."""
    from oracle import restate as R
    torch.set_num_threads(os.cpu_count() or 1)
    z1, z2, j1, j2 = _make_batch(N_PER_VIEW)
    z = torch.cat([z1, z2], 0)
    bj = torch.cat((j1[:, :, :2], j2[:, :, :2]), dim=0)
    m = z.shape[0]
    dmax = 70.0   # value only scales the weights; timing is data independent
    times = []
    it = 0
    # at least `repeats` samples and, for the cpu_baseline leg, about min_seconds of CPU work (bounded at 400 samples)
    while len(times) < repeats or (sum(times) < min_seconds and len(times) < 400):
        r0 = (it * rows) % (m - rows)
        t0 = time.perf_counter()
        R.port_step_rows(z, bj, r0, r0 + rows, dmax, TAU)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
    repeats = len(times)
    t_sample = statistics.median(times)
    steps_per_s = 1.0 / (t_sample * (m / rows))
    return dict(value=steps_per_s, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{rows} of {m} rows of one fwd+bwd step (weights, logits, exp, row sums, autograd) "
                       f"x{repeats}, extrapolated x{m // rows}; median {t_sample * 1e3:.0f} ms per sample",
                pairs_per_s=rows * m / t_sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_port_sample(rows=128, repeats=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / base["value"], higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, global_batch=N_PER_VIEW, proj_dim=DIM),
                cpu_baseline=dict(value=base["value"], unit=UNIT, cores=base["cores"], kind=base["kind"],
                                  sample=base["sample"]),
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    _emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
WORKLOAD = ("handclr_w loss fwd+bwd, global batch 8192 (2N=16384), d=128, 21 joints, mpjpe/linear/pos_neg, tau 0.5")


def run_ours(args):
    import torch.distributed as dist
    from simhand_b200 import _lib, ops
    from simhand_b200.dist import run_step_sharded
    from simhand_b200.pipeline import HostPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    n_local = N_PER_VIEW // world
    z1, z2, j1, j2 = _make_batch(N_PER_VIEW)
    sl = slice(rank * n_local, (rank + 1) * n_local)
    hz1, hz2 = z1[sl].contiguous().pin_memory(), z2[sl].contiguous().pin_memory()
    hj1, hj2 = j1[sl].contiguous().pin_memory(), j2[sl].contiguous().pin_memory()
    dz1_, dz2_, dj1, dj2 = hz1.to(dev), hz2.to(dev), hj1.to(dev), hj2.to(dev)
    engine = args.engine
    transport = args.transport
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier(group)
        torch.cuda.synchronize(dev)

    def make_step(exact):
        """The library's step on device-resident inputs (what `value` times)."""
        def fn(a, b, c, e):
            if world == 1:
                return ops.run_step(a, b, c[:, :, :2], e[:, :, :2], TAU, engine, True, exact_weights=exact)
            return run_step_sharded(a, b, c[:, :, :2], e[:, :, :2], TAU, engine, True, group, transport=transport,
                                    exact_weights=exact)
        return fn

    def make_dropin(exact):
        """The reference's call pattern through the public API (simhand_w_model.py:122-136): get_weights_linear ->
        vanila_weights_contrastive_loss -> backward (what `e2e` times); on several ranks weighted_ntxent(group=...)."""
        def fn(a, b, c, e):
            a = a.detach().requires_grad_(True)
            b = b.detach().requires_grad_(True)
            if world == 1:
                pw, nw = ops.get_weights_linear(c[:, :, :2], e[:, :, :2], "mpjpe")
                loss = ops.vanila_weights_contrastive_loss(a, b, pw, nw, TAU, engine=engine, exact_weights=exact)
            else:
                loss = ops.weighted_ntxent(a, b, c[:, :, :2], e[:, :, :2], TAU, group=group, engine=engine,
                                           exact_weights=exact)
            g1, g2 = torch.autograd.grad(loss, (a, b))
            return loss, g1, g2
        return fn

    # The step is a fixed sequence of kernel launches on one stream (no host decisions, no NCCL with the peer
    # transports): capture it once into a CUDA graph and replay it, as a training loop would.
    use_graph = args.graph and (world == 1 or transport != "nccl")

    def time_device(step_fn):
        graph, static_out = None, None
        if use_graph:
            for _ in range(3):
                step_fn(dz1_, dz2_, dj1, dj2)
            barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = step_fn(dz1_, dz2_, dj1, dj2)
            barrier()

        def step():
            if graph is None:
                return step_fn(dz1_, dz2_, dj1, dj2)
            graph.replay()
            return static_out

        for _ in range(warmup):
            out = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = step()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), float(out[0]), graph is not None

    def time_e2e(step_fn):
        """Host buffers in, loss out, copies inside the timed region: simhand_b200.HostPipeline copies batch k+1 H2D on a
        copy stream while batch k computes, replays the captured drop-in step, reads every step's loss back to pinned
        host memory and hands it to the caller one step later (one step in flight); the last loss is drained inside the
        timed region too."""
        pipe = HostPipeline(step_fn, (hz1, hz2, hj1, hj2), dev, depth=2, use_graph=use_graph, sync_all=barrier, lag=1)
        pipe.prefetch(hz1, hz2, hj1, hj2)
        for _ in range(3):
            pipe.prefetch(hz1, hz2, hj1, hj2)
            pipe.step()
        pipe.step()
        pipe.drain()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        pipe.prefetch(hz1, hz2, hj1, hj2)
        host_loss = None
        for k in range(steps):
            if k + 1 < steps:
                pipe.prefetch(hz1, hz2, hj1, hj2)
            got, _, _ = pipe.step()                    # the previous step's loss, already on the host
            if got is not None:
                host_loss = float(got)
        got, _, _ = pipe.drain()
        host_loss = float(got)
        f1.record()
        barrier()
        return f0.elapsed_time(f1), host_loss

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, loss_dev, graphed = time_device(make_step(False))
    ms_exact, loss_exact, _ = time_device(make_step(True))
    ms_e2e, loss_e2e = time_e2e(make_dropin(False))
    ms_e2e_exact, loss_e2e_exact = time_e2e(make_dropin(True))
    if sampler:
        sampler.stop()

    t = torch.tensor([ms_total, ms_exact, ms_e2e, ms_e2e_exact], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms_total, ms_exact, ms_e2e, ms_e2e_exact = t.tolist()
    ms_step = ms_total / steps
    value = 1e3 / ms_step

    # ---- per-launch durations of the same step (instrumented eager pass), both distance images
    iters = max(3, min(steps, 10))
    if world == 1:
        kernels = {False: time_kernels(ops, _lib, dz1_, dz2_, dj1, dj2, engine, iters, False),
                   True: time_kernels(ops, _lib, dz1_, dz2_, dj1, dj2, engine, iters, True)}
    elif world > 1:
        # per-launch times of the fused form of the exchange (whatever transport the headline ran with): its six launches
        # map one to one onto the step's phases, and each launch's time includes its wait for the other ranks
        kernels = {False: time_kernels_sharded(ops, _lib, dz1_, dz2_, dj1, dj2, engine, iters, False, group, barrier),
                   True: time_kernels_sharded(ops, _lib, dz1_, dz2_, dj1, dj2, engine, iters, True, group, barrier)}
    else:
        kernels = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = _peaks()
    clocks = sampler.summary() if sampler else None
    m = 2 * N_PER_VIEW
    h2d = int(hz1.numel() * 8 + hj1.numel() * 8) * world
    resolved_transport = None
    if world > 1:
        resolved_transport = os.environ.get("SMH_TRANSPORT", "peer") if transport == "auto" else transport
    launches = 6 if world == 1 else (6 if resolved_transport == "fused" else 14)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=ms_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype={"tf32": "tf32 logits / bf16 value operands, fp32 accumulate", "bf16": "bf16",
                       "fp16": "f16 logits (11-bit significand, as tf32) / bf16 value operands, fp32 accumulate"
                       }.get(engine, "f32"), data="synthetic",
                config=dict(workload=WORKLOAD, global_batch=N_PER_VIEW, proj_dim=DIM,
                            engine=engine, parallelism=f"tile-pair sharded x{world}" if world > 1 else "single GPU",
                            transport=resolved_transport, cuda_graph=bool(graphed),
                            weights="value / e2e: relaxed-weights mode (16-bit image of the joint distances, |dW| <= 1.6e-5, "
                                    "loss and gradients inside BASELINE.json's tolerances at this size: "
                                    "tests/test_gpu_fullsize.py); exact_weights: bit-exact distances (0 ulp, the 1-ulp "
                                    "weight contract) in the fused step",
                            l2="per-step working set (MPJPE tile workspace, 0.27 / 0.54 GiB at 1 GPU) exceeds the 126 MB L2; "
                               "no explicit flush"),
                clocks=clocks,
                e2e=dict(value=1e3 / (ms_e2e / steps), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4 * world,
                         bytes_scope="whole job (sum over ranks; every rank copies its own shard and reads the loss)",
                         loss=loss_e2e,
                         api="simhand_b200.HostPipeline(lag=1) over the reference's call pattern: get_weights_linear -> "
                             "vanila_weights_contrastive_loss -> autograd backward (simhand_w_model.py:122-136)"
                             if world == 1 else
                             "simhand_b200.HostPipeline(lag=1) over weighted_ntxent(group=WORLD) -> autograd backward"),
                exact_weights=dict(value=1e3 / (ms_exact / steps), unit=UNIT, ms_per_step=ms_exact / steps, loss=loss_exact,
                                   e2e=dict(value=1e3 / (ms_e2e_exact / steps), unit=UNIT, loss=loss_e2e_exact,
                                            h2d_bytes_per_step=h2d, d2h_bytes_per_step=4 * world),
                                   note="same step with exact_weights=True: fp32 tiles of the bit-exact MPJPE (weights 0 ulp "
                                        "from the reference), correctly rounded sqrt in the distance kernel"),
                loss=loss_dev,
                gpu_launches=launches * steps)
    if kernels is not None:
        f_clk = (clocks or {}).get("sm_mhz") or peaks["sm_max_mhz"]
        xu_nominal = NUM_SMS * MUFU_LANES_PER_SM * f_clk * 1e6 / 1e9       # G special-function ops / s
        mufu = _mufu_peak()
        tiles_total = (m // 128) * (m // 128 + 1) // 2                     # stored (upper-triangular) MPJPE tiles
        tiles = tiles_total / world                                        # per rank (balanced to +-1 tile)

        def roofline_of(kern, exact):
            # denominator: the MUFU rate of the instruction the kernel issues, measured on this pool (scaled to the SM
            # clock seen during this run); the nominal 16 lanes / clk / SM when the measurement file is absent
            if mufu:
                xu_peak = mufu["mufu_rsq_gops" if exact else "mufu_sqrt_gops"] * f_clk / mufu["clock_rate_mhz"]
                src = (f"measured {'MUFU.RSQ' if exact else 'MUFU.SQRT'} throughput (tools/mufu_peak.cu -> profiles/r02_mufu_peak.json, "
                       f"{mufu['clock_rate_mhz']:.0f} MHz) scaled to the {f_clk:.0f} MHz seen under load; nominal 148 x 16 lanes/clk = "
                       f"{xu_nominal:.0f} Gop/s")
            else:
                xu_peak = xu_nominal
                src = f"nominal: 148 SMs x 16 MUFU lanes/clk x {f_clk:.0f} MHz (median SM clock under load)"
            t_mpjpe = kern["mpjpe_kernel"] * 1e-3
            traffic = _ncu_traffic("mpjpe_kernel<0," if exact else "mpjpe_kernel<1,")
            executed = 21.0 * tiles * 128 * 128 / t_mpjpe / 1e9            # sqrt one rank's launch evaluates / its duration
            step_ms = (ms_exact if exact else ms_total) / steps
            return dict(
                kernel="mpjpe_kernel", bound="xu (MUFU pipe; neither HBM nor tensor binds this path, SURVEY.md 8d)",
                achieved=executed, peak=xu_peak,
                unit=("Gop/s per GPU: correctly rounded sqrt (MUFU.RSQ + Newton step on the FMA pipe), 21 per pair the launch evaluates"
                      if exact else
                      "Gop/s per GPU: approximate sqrt (relaxed-weights mode), 21 per pair the launch evaluates -- 19 of them one "
                      "MUFU.SQRT each, 2 on the FMA pipe (sqrt2_fma_pipe)"),
                frac=executed / xu_peak,
                xu_pipe_frac=(executed if exact else executed * 19.0 / 21.0) / xu_peak,
                traffic=(traffic or {}).get("bytes") if world == 1 else None,
                traffic_source=(traffic or {}).get("source") if world == 1 else None,
                peak_source=src,
                units_per_launch=f"{tiles:.0f} tiles x 16384 unordered pairs per rank (symmetry: D_ij == D_ji bitwise)",
                algorithmic_frac=(21.0 * m * m / world / t_mpjpe / 1e9) / xu_peak,
                step_frac=(22.0 * m * m / world / (step_ms * 1e-3) / 1e9) / xu_peak,
                note="frac counts the sqrt the kernel executes over the MUFU rate; xu_pipe_frac counts only those that are MUFU "
                     "instructions (what ncu's sm__inst_executed_pipe_xu shows).  algorithmic_frac counts 21 per ORDERED pair (M^2 / P per rank, "
                     "as the reference evaluates them) over the same launch time and step_frac is SURVEY 8d's figure, "
                     "(21 sqrt + 1 exp) M^2 / (P x whole step time) over the MUFU peak: both exceed frac because symmetry "
                     "halves the executed count.")
        line["roofline"] = roofline_of(kernels[False], False)
        line["exact_weights"]["roofline"] = roofline_of(kernels[True], True)
        for exact in (False, True):
            tile_bytes = 65536.0 if exact else 32768.0       # fp32 tiles, or their 16-bit image
            hbm_bytes = tiles * tile_bytes * 2               # each stored tile is read direct + transposed
            for k in ("sweep_fwd", "sweep_bwd"):
                kernels[exact][k + "_hbm_frac"] = hbm_bytes / (kernels[exact][k] * 1e-3) / 1e9 / peaks["hbm_gbs"]
        line["kernels_ms"] = kernels[False]
        line["exact_weights"]["kernels_ms"] = kernels[True]
        line["peaks"] = peaks
    if world == 1:
        line["cpu_baseline"] = cpu_port_sample(rows=128, repeats=3, warmup=1, min_seconds=10.0)
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def time_kernels(ops, _lib, z1, z2, j1, j2, engine, iters, exact):
    """CUDA-event duration of every launch of one step (same stream, same inputs), averaged."""
    import ctypes
    lib = _lib.load()
    eng = _lib.ENGINES[engine]
    dev = z1.device
    n, d = z1.shape
    ctx = ops.get_context(n, d, 1, 0, dev, 0, ops.step_flags(engine, exact_weights=exact))
    inp, keep = ops.make_inputs(z1, z2, j1[:, :, :2], j2[:, :, :2])
    ws = torch.empty(int(ctx.layout.ws_bytes), dtype=torch.uint8, device=dev)
    loss = torch.empty((), device=dev)
    g1, g2 = torch.empty((n, d), device=dev), torch.empty((n, d), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    pd, pi, plan = ctypes.byref(ctx.dims), ctypes.byref(inp), ctx.plan_dev.data_ptr()
    calls = [
        ("prep", lambda: lib.smh_prep(pd, pi, ws.data_ptr(), eng, st)),
        ("mpjpe_kernel", lambda: lib.smh_mpjpe(pd, plan, ws.data_ptr(), None, st)),
        ("sweep_fwd", lambda: lib.smh_forward(pd, plan, ws.data_ptr(), TAU, eng, None, st)),
        ("sweep_bwd", lambda: lib.smh_backward(pd, plan, ws.data_ptr(), TAU, eng, None, st)),
        ("finalize", lambda: lib.smh_finalize(pd, pi, ws.data_ptr(), None, TAU, 1.0, loss.data_ptr(), g1.data_ptr(),
                                              g2.data_ptr(), d, 0, None, st)),
    ]
    acc = {k: 0.0 for k, _ in calls}
    for it in range(iters + 1):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        evs[0].record()
        for i, (_, fn) in enumerate(calls):
            _lib.check(fn())
            evs[i + 1].record()
        torch.cuda.synchronize(dev)
        if it == 0:
            continue
        for i, (k, _) in enumerate(calls):
            acc[k] += evs[i].elapsed_time(evs[i + 1])
    return {k: v / iters for k, v in acc.items()}


def time_kernels_sharded(ops, _lib, z1, z2, j1, j2, engine, iters, exact, group, barrier):
    """Fused transport: CUDA-event duration of each of the six launches of a rank (a launch's time includes its wait for
    the other ranks' stage signal), max over ranks."""
    import torch.distributed as dist
    from simhand_b200 import dist as sd
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = z1.device
    n_local, d = z1.shape
    n = n_local * world
    eng = _lib.ENGINES[engine]
    ctx = ops.get_context(n, d, world, rank, dev, 0, ops.step_flags(engine, exact_weights=exact))
    ex = sd.get_exchange(ctx, group, 0, fused=True)
    local_in, keep = ops.make_inputs(z1, z2, j1[:, :, :2], j2[:, :, :2])
    loss = torch.empty((), device=dev)
    g1, g2 = torch.empty((n_local, d), device=dev), torch.empty((n_local, d), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    names = {"prep": "shard_prep", "zpush": "shard_push_z (overlaps the MPJPE kernel in the real step)", "mpjpe": "mpjpe_kernel",
             "fwd": "sweep_fwd", "bwd": "sweep_bwd", "fin": "finalize"}
    acc = {v: 0.0 for v in names.values()}
    for it in range(iters + 1):
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(sd.FUSED_STAGES) + 1)]
        evs[0].record()
        for i, stage in enumerate(sd.FUSED_STAGES):
            sd._fused_launches(lib, ctx, ex.struct, ex.ws.data_ptr(), local_in, TAU, eng, True, 1.0, (loss, g1, g2), st,
                               stages=(stage,))
            evs[i + 1].record()
        torch.cuda.synchronize(dev)
        if it == 0:
            continue
        for i, stage in enumerate(sd.FUSED_STAGES):
            acc[names[stage]] += evs[i].elapsed_time(evs[i + 1])
    t = torch.tensor([acc[v] / iters for v in names.values()], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return dict(zip(names.values(), t.tolist()))


_JSON_FD = None


def _claim_stdout():
    """Native libraries (NCCL's version banner) write to fd 1: route everything but the one JSON line to stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    sys.stdout.flush()
    payload = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, payload)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="fp16", choices=["tf32", "fp32", "auto", "bf16", "fp16"])
    ap.add_argument("--transport", default="auto", choices=["auto", "fused", "peer", "nccl"],
                    help="multi-GPU exchange: fused into the kernels over peer memory (default), the same as separate "
                         "push / barrier kernels, or NCCL calls")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch the step eagerly")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 30)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
