#!/usr/bin/env python
"""Where the sweep CTAs spend their time, per warp role (development build with cycle counters):
    python -m simhand_b200.build --trace && python tools/trace_sweeps.py [world=1] [rank=0]
Prints, averaged over the busy CTAs, the microseconds each role spent waiting on each barrier and working."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SMH_LIB"] = os.path.join(ROOT, "simhand_b200", "lib", "libsimhand_b200_trace.so")
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from simhand_b200 import _lib, ops, synth  # noqa: E402

ROLES = {
    0: ("tile producer", ["other", "wait free tile stage", "", "", "", "", ""]),
    1: ("operand producer", ["other", "wait A buffer", "wait free z stage", "", "", "", ""]),
    2: ("MMA issuer", ["other", "wait A | poll before value MMA", "wait z block | poll before logit MMA",
                       "wait free S buffer | issue value MMAs", "issue logit MMAs", "idle polls", ""]),
    3: ("epilogue warp 0", ["other", "wait tile", "wait S", "tcgen05.ld", "weights/exp/sums/G", "strip flush", "#tasks"]),
}


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda:0")
    n, d = 8192, 128
    lib = _lib.load()
    z1, z2, j1, j2 = synth.make_batch(n, d, 5, "hand")
    z1, z2, a, b = z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    ctx = ops.get_context(n, d, world, rank, dev, 0, _lib.DIMS_Q16_TILES)
    lay = ctx.layout
    inp, keep = ops.make_inputs(z1, z2, a, b)
    ws = torch.empty(int(lay.ws_bytes), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    eng = _lib.ENGINES["fp16"]
    pd, pi, plan = ctypes.byref(ctx.dims), ctypes.byref(inp), ctx.plan_dev.data_ptr()
    mhz = 1965.0          # SM clock of the B200 boxes under load (nvidia-smi during bench.py)
    for name, call in (("forward", lambda: lib.smh_forward(pd, plan, ws.data_ptr(), 0.5, eng, None, st)),
                       ("backward", lambda: lib.smh_backward(pd, plan, ws.data_ptr(), 0.5, eng, None, st))):
        for rep in range(3):
            _lib.check(lib.smh_prep(pd, pi, ws.data_ptr(), eng, st), "prep")
            _lib.check(lib.smh_mpjpe(pd, plan, ws.data_ptr(), None, st), "mpjpe")
            if name == "backward":
                _lib.check(lib.smh_forward(pd, plan, ws.data_ptr(), 0.5, eng, None, st), "fwd")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(call(), name)
            e1.record()
            torch.cuda.synchronize()
        tr = ws[lay.off_rowloss:lay.off_rowloss + 148 * 4 * 8 * 8].view(torch.int64).view(148, 4, 8).cpu().double()
        busy = tr[:, 3, 6] > 0
        print(f"== {name} sweep, world {world} rank {rank}: {e0.elapsed_time(e1) * 1e3:.1f} us (traced build), "
              f"{int(busy.sum())} busy CTAs, {tr[busy, 3, 6].mean() * 2:.1f} tasks per CTA, SM clock {mhz:.0f} MHz")
        for slot, (role, labels) in ROLES.items():
            t = tr[busy, slot] / mhz
            parts = ", ".join(f"{lab} {t[:, i].mean():.1f}" for i, lab in enumerate(labels) if lab and lab != "#tasks")
            print(f"   {role:18s} total {t[:, 7].mean():6.1f} us (max {t[:, 7].max():6.1f}) | {parts}")
    del keep


if __name__ == "__main__":
    main()
