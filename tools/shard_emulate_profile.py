#!/usr/bin/env python
"""Per-stage CUDA-event times of the fused sharded step with all ranks emulated on ONE GPU (simhand_b200.dist.EmulatedGroup):
what one rank's six launches cost at world = 2 / 4 / 8 without NVLink in the way.
    python tools/shard_emulate_profile.py [world=8] [n=8192] [engine=fp16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import synth  # noqa: E402
from simhand_b200.dist import FUSED_STAGES, EmulatedGroup  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    engine = sys.argv[3] if len(sys.argv) > 3 else "fp16"
    dev = torch.device("cuda:0")
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    z1, z2, a, b = z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    grp = EmulatedGroup(n, 128, world, dev, engine)
    for _ in range(3):
        grp.step(z1, z2, a, b)
    timing = {}
    iters = 20
    for _ in range(iters):
        grp.step(z1, z2, a, b, timing=timing)
    torch.cuda.synchronize()
    assert grp.poisoned() == [0] * world, grp.poisoned()
    line = []
    total = 0.0
    for stage in FUSED_STAGES:
        per_rank = [sum(e0.elapsed_time(e1) for e0, e1 in timing[(stage, r)]) / iters * 1e3 for r in range(world)]
        total += max(per_rank)
        line.append(f"{stage} {min(per_rank):.1f}-{max(per_rank):.1f}us")
    lay = grp.ctxs[0].layout
    print(f"world {world} n {n} {engine}: tiles {lay.n_stored_tiles} tasks {lay.n_tasks} strips {lay.n_strips} | " + " ".join(line) +
          f" | sum of slowest ranks {total:.1f}us (eager launches, one GPU, no NVLink)", flush=True)


if __name__ == "__main__":
    main()
