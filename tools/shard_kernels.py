#!/usr/bin/env python
"""Times the kernels of ONE rank of a P-way sharded step on a single GPU (no exchange: the rank's plan over the full
gathered batch), to study the short-run behaviour of the sharded kernels without an 8-GPU box.
    python tools/shard_kernels.py [world=8] [rank=0]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import _lib, ops, synth  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda:0")
    n, d = 8192, 128
    lib = _lib.load()
    z1, z2, j1, j2 = synth.make_batch(n, d, 5, "hand")
    z1, z2, a, b = z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    ctx = ops.get_context(n, d, world, rank, dev, 0, _lib.DIMS_Q16_TILES if os.environ.get('SMH_Q16', '1') == '1' else 0)
    lay = ctx.layout
    # the full batch described as one "rank" of n samples (n_local = n): same addressing as a gathered buffer
    inp, keep = ops.make_inputs(z1, z2, a, b)
    ws = torch.empty(int(lay.ws_bytes), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    eng = _lib.ENGINES["fp16"]
    pd, pi, plan = ctypes.byref(ctx.dims), ctypes.byref(inp), ctx.plan_dev.data_ptr()
    calls = [("prep", lambda: lib.smh_prep(pd, pi, ws.data_ptr(), eng, st)),
             ("mpjpe", lambda: lib.smh_mpjpe(pd, plan, ws.data_ptr(), None, st)),
             ("fwd", lambda: lib.smh_forward(pd, plan, ws.data_ptr(), 0.5, eng, None, st)),
             ("bwd", lambda: lib.smh_backward(pd, plan, ws.data_ptr(), 0.5, eng, None, st))]
    iters = 30
    acc = {k: 0.0 for k, _ in calls}
    for it in range(iters + 3):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        evs[0].record()
        for i, (k, fn) in enumerate(calls):
            _lib.check(fn(), k)
            evs[i + 1].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, (k, _) in enumerate(calls):
                acc[k] += evs[i].elapsed_time(evs[i + 1])
    print(f"world {world} rank {rank}: tiles {lay.n_stored_tiles} tasks {lay.n_tasks} strips {lay.n_strips} | " +
          " ".join(f"{k} {v / iters * 1e3:.1f}us" for k, v in acc.items()), flush=True)
    del keep


if __name__ == "__main__":
    main()
