#!/usr/bin/env python
"""Per-launch CUDA-event times of the fused sharded step with all ranks emulated on ONE GPU (simhand_b200.dist.EmulatedGroup):
what each of a rank's six launches costs at world = 2 / 4 / 8 without NVLink in the way.  Every (stage, rank) launch is
captured into its own CUDA graph, so the host's launch path is not in the measurement.
    python tools/shard_emulate_profile.py [world=8] [n=8192] [engine=fp16] [exact=0]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import _lib, synth  # noqa: E402
from simhand_b200.dist import FUSED_STAGES, EmulatedGroup, _fused_launches  # noqa: E402
from simhand_b200.ops import _stream_ptr, make_inputs  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    engine = sys.argv[3] if len(sys.argv) > 3 else "fp16"
    exact = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
    direct = len(sys.argv) > 5 and sys.argv[5] == "direct"      # under ncu: a few eager steps, the profiler times the kernels
    dev = torch.device("cuda:0")
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    z1, z2, a, b = z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    grp = EmulatedGroup(n, 128, world, dev, engine, exact_weights=exact)
    for _ in range(3):
        grp.step(z1, z2, a, b)
    torch.cuda.synchronize()
    if direct:
        assert grp.poisoned() == [0] * world, grp.poisoned()
        return
    lib = _lib.load()
    n_local = n // world
    eng = _lib.ENGINES[grp.engine_name]
    locals_, outs, keeps = [], [], []
    for r in range(world):
        sl = slice(r * n_local, (r + 1) * n_local)
        li, keep = make_inputs(z1[sl], z2[sl], a[sl], b[sl])
        locals_.append(li)
        keeps.append(keep)
        outs.append((torch.empty((), device=dev), torch.empty(n_local, 128, device=dev), torch.empty(n_local, 128, device=dev)))
    graphs = {}
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for stage in FUSED_STAGES:
            for r in range(world):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    _fused_launches(lib, grp.ctxs[r], grp.structs[r], grp.ws[r].data_ptr(), locals_[r], 0.5, eng, True, 1.0,
                                    outs[r], _stream_ptr(dev), stages=(stage,))
                graphs[(stage, r)] = g
    torch.cuda.synchronize()
    iters = 20
    acc = {k: 0.0 for k in graphs}
    for it in range(iters + 2):
        evs = {}
        for stage in FUSED_STAGES:
            for r in range(world):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graphs[(stage, r)].replay()
                e1.record()
                evs[(stage, r)] = (e0, e1)
        torch.cuda.synchronize()
        if it >= 2:
            for k, (e0, e1) in evs.items():
                acc[k] += e0.elapsed_time(e1)
    assert grp.poisoned() == [0] * world, grp.poisoned()
    line, total = [], 0.0
    for stage in FUSED_STAGES:
        per_rank = [acc[(stage, r)] / iters * 1e3 for r in range(world)]
        total += max(per_rank)
        line.append(f"{stage} {min(per_rank):.1f}-{max(per_rank):.1f}us")
    lay = grp.ctxs[0].layout
    print(f"world {world} n {n} {engine} exact={int(exact)}: tiles {lay.n_stored_tiles} tasks {lay.n_tasks} strips {lay.n_strips} | " +
          " ".join(line) + f" | sum of slowest ranks {total:.1f}us (graph-replayed launches, one GPU, no NVLink; a replay "
          "adds ~3-5us of launch latency to each figure)", flush=True)
    del keeps


if __name__ == "__main__":
    main()
