#!/usr/bin/env python
"""Per-phase CUDA-event timing of the sharded step (peer exchange), run under torchrun."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import _lib, synth  # noqa: E402
from simhand_b200.dist import gathered_views, get_exchange, pack_local  # noqa: E402
from simhand_b200.ops import get_context  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    lib = _lib.load()
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    n_local = n // world
    sl = slice(rank * n_local, (rank + 1) * n_local)
    a, b = z1[sl].to(dev), z2[sl].to(dev)
    c, e = j1[sl].to(dev)[:, :, :2], j2[sl].to(dev)[:, :, :2]
    ctx = get_context(n, 128, world, rank, dev, 0, _lib.DIMS_Q16_TILES)
    lay, dims = ctx.layout, ctx.dims
    from simhand_b200.ops import make_inputs
    chunk = 2 * n_local * (128 + 42)
    local_in, keep = make_inputs(a, b, c, e)
    ex = get_exchange(ctx, group, chunk)
    px = ctypes.byref(ex.struct)
    ws = ex.ws
    (o1, o2, oj1, oj2), _ = gathered_views(ex.xin, world, n_local, 128)
    base = ex.xin.data_ptr()
    inp = _lib.Inputs(base + 4 * o1, base + 4 * o2, 128, base + 4 * oj1, base + 4 * oj2, 42, 2, 1, n_local, chunk, chunk)
    st = torch.cuda.current_stream().cuda_stream
    pd, pi, plan = ctypes.byref(dims), ctypes.byref(inp), ctx.plan_dev.data_ptr()
    loss = torch.empty((), device=dev)
    g1, g2 = torch.empty((n_local, 128), device=dev), torch.empty((n_local, 128), device=dev)
    eng = _lib.ENGINES["fp16"]
    dz_src = ws.data_ptr() + int(lay.off_dzacc)
    calls = [
        ("push", lambda: lib.smh_push_inputs(px, ctypes.byref(local_in), n_local, 128, st)),
        ("zero", lambda: lib.smh_prep_zero(pd, ws.data_ptr(), st)),
        ("barrier1", lambda: lib.smh_barrier(px, st)),
        ("prep", lambda: lib.smh_prep(pd, pi, ws.data_ptr(), eng | _lib.PREP_NO_ZERO, st)),
        ("mpjpe", lambda: lib.smh_mpjpe(pd, plan, ws.data_ptr(), px, st)),
        ("barrier2", lambda: lib.smh_barrier(px, st)),
        ("fwd", lambda: lib.smh_forward(pd, plan, ws.data_ptr(), 0.5, eng, px, st)),
        ("xneg", lambda: lib.smh_exchange_neg(pd, ws.data_ptr(), px, st)),
        ("barrier3", lambda: lib.smh_barrier(px, st)),
        ("bwd", lambda: lib.smh_backward(pd, plan, ws.data_ptr(), 0.5, eng, px, st)),
        ("fin_loss", lambda: lib.smh_finalize(pd, pi, ws.data_ptr(), None, 0.5, 1.0, loss.data_ptr(), g1.data_ptr(),
                                              g2.data_ptr(), 128, _lib.FINALIZE_LOSS_PART, px, st)),
        ("xdz", lambda: lib.smh_exchange_dz(pd, ws.data_ptr(), px, st)),
        ("barrier4", lambda: lib.smh_barrier(px, st)),
        ("fin_grad", lambda: lib.smh_finalize(pd, pi, ws.data_ptr(), None, 0.5, 1.0, loss.data_ptr(), g1.data_ptr(),
                                              g2.data_ptr(), 128, _lib.FINALIZE_GRAD, px, st)),
    ]
    iters = 20
    acc = {k: 0.0 for k, _ in calls}
    for it in range(iters + 2):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)]
        evs[0].record()
        for i, (_, fn) in enumerate(calls):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        if it < 2:
            continue
        for i, (k, _) in enumerate(calls):
            acc[k] += evs[i].elapsed_time(evs[i + 1])
    tot = sum(acc.values()) / iters
    print(f"rank {rank}: total {tot:.3f} ms | " + " ".join(f"{k} {v / iters * 1e3:.0f}us" for k, v in acc.items()), flush=True)
    print(f"rank {rank}: stored tiles {lay.n_stored_tiles} tasks {lay.n_tasks} strips {lay.n_strips}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
