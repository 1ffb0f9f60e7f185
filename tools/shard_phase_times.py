#!/usr/bin/env python
"""Phase clocks of the fused sharded step on real GPUs (torchrun, one rank per GPU): block 0 of every fused kernel stamps
its phases into the rank's signal block (smh_common.cuh: PhaseClock); this prints them next to the CUDA-event time of each
launch.   torchrun --nproc-per-node 2 tools/shard_phase_times.py [n_per_view=8192] [exact=0]"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import _lib, ops, synth  # noqa: E402
from simhand_b200 import dist as sd  # noqa: E402

PHASES = {
    0: ("prep", ["issue", "pivot", "bound", "fence", "end"]),
    2: ("fwd", ["wait", "sweep", "gridbar", "copy", "fence"]),
    4: ("bwd", ["head", "sweep", "gridbar", "copy", "fence"]),
    1: ("mpjpe", ["wait"]),
    5: ("fin", ["wait", "loss", "grads"]),
}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    exact = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    n_local = n // world
    sl = slice(rank * n_local, (rank + 1) * n_local)
    a, b, c, e = z1[sl].to(dev), z2[sl].to(dev), j1[sl].to(dev)[:, :, :2], j2[sl].to(dev)[:, :, :2]
    lib = _lib.load()
    eng = _lib.ENGINES["fp16"]
    ctx = ops.get_context(n, 128, world, rank, dev, 0, ops.step_flags("fp16", exact_weights=exact))
    ex = sd.get_exchange(ctx, group, 0, fused=True)
    local_in, keep = ops.make_inputs(a, b, c, e)
    loss = torch.empty((), device=dev)
    g1, g2 = torch.empty((n_local, 128), device=dev), torch.empty((n_local, 128), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    graphs = {}
    for _ in range(3):
        sd._fused_launches(lib, ctx, ex.struct, ex.ws.data_ptr(), local_in, 0.5, eng, True, 1.0, (loss, g1, g2), st)
    torch.cuda.synchronize()
    dist.barrier()
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for stage in sd.FUSED_STAGES:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                sd._fused_launches(lib, ctx, ex.struct, ex.ws.data_ptr(), local_in, 0.5, eng, True, 1.0, (loss, g1, g2),
                                   torch.cuda.current_stream(dev).cuda_stream, stages=(stage,))
            graphs[stage] = g
    torch.cuda.synchronize()
    dist.barrier()
    iters = 20
    acc = {s: 0.0 for s in sd.FUSED_STAGES}
    clocks = torch.zeros(6, 16, dtype=torch.float64)
    for it in range(iters + 2):
        dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(sd.FUSED_STAGES) + 1)]
        evs[0].record()
        for i, stage in enumerate(sd.FUSED_STAGES):
            graphs[stage].replay()
            evs[i + 1].record()
        torch.cuda.synchronize()
        if it >= 2:
            for i, stage in enumerate(sd.FUSED_STAGES):
                acc[stage] += evs[i].elapsed_time(evs[i + 1])
            clocks += ex.signal[160:160 + 96].cpu().view(torch.int32).to(torch.int64).bitwise_and(0xffffffff).double().view(6, 16)
    clocks /= iters
    # the real step: one CUDA graph (z push on its parallel branch); intervals between the kernels' absolute start stamps
    with torch.cuda.stream(side):
        for _ in range(3):
            sd._fused_launches(lib, ctx, ex.struct, ex.ws.data_ptr(), local_in, 0.5, eng, True, 1.0, (loss, g1, g2),
                               torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
        gstep = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gstep, stream=side):
            sd._fused_launches(lib, ctx, ex.struct, ex.ws.data_ptr(), local_in, 0.5, eng, True, 1.0, (loss, g1, g2),
                               torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    order = [0, 1, 2, 4, 5]                     # prep, mpjpe, fwd, bwd, finalize
    gaps = torch.zeros(len(order), dtype=torch.float64)
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(reps + 3):
        if it == 3:
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
        gstep.replay()
        gstep.replay()
        torch.cuda.synchronize()
        if it >= 3:
            stamps = ex.signal[160:160 + 96].cpu().view(torch.int32).to(torch.int64).bitwise_and(0xffffffff).view(6, 16)[:, 15]
            for i, k in enumerate(order[:-1]):
                gaps[i] += float((int(stamps[order[i + 1]]) - int(stamps[k])) % (1 << 32))
    e1.record()
    torch.cuda.synchronize()
    gaps /= reps * 1e3
    step_us = e0.elapsed_time(e1) / (2 * reps) * 1e3
    names5 = ["prep", "mpjpe", "fwd", "bwd"]
    graph_line = " ".join(f"{n}->next {gaps[i]:.1f}us" for i, n in enumerate(names5))
    line = f"rank {rank}/{world} exact={int(exact)} poisoned={ex.poisoned()} | " + " ".join(
        f"{s} {acc[s] / iters * 1e3:.1f}us" for s in sd.FUSED_STAGES) + f" | total {sum(acc.values()) / iters * 1e3:.1f}us"
    line += f"\n    whole-step graph (includes host sync every 2 steps): ~{step_us:.1f}us/step; start-to-start: {graph_line} "
    for k, (name, phases) in PHASES.items():
        line += f"\\n    {name}: " + " ".join(f"{p}@{clocks[k][i] / 1e3:.1f}us" for i, p in enumerate(phases))
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(line, flush=True)
    del keep
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
