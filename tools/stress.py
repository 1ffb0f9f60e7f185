#!/usr/bin/env python
"""Stress the single-GPU step: many back-to-back steps (eager and CUDA-graph replay), checking the loss bits, the
pipeline fail_site and CUDA errors.  python tools/stress.py [steps] [engine]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import ops, synth  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    engine = sys.argv[2] if len(sys.argv) > 2 else "tf32"
    dev = torch.device("cuda")
    z1, z2, j1, j2 = synth.make_batch(8192, 128, 5, "hand")
    a, b, c, e = z1.to(dev), z2.to(dev), j1.to(dev), j2.to(dev)
    ref = None
    for it in range(steps):
        loss, g1, g2, aux = ops.run_step(a, b, c[:, :, :2], e[:, :, :2], 0.5, engine, True, return_aux=True)
        if it % 20 == 0 or it == steps - 1:
            torch.cuda.synchronize()
            st = aux["stats"].cpu().numpy()
            val = float(loss)
            if ref is None:
                ref = val
            print(f"eager it {it}: loss {val:.7f} fail_site {st[6]} flags {st[3]} gsum {float(g1.sum()):.6e}", flush=True)
            assert st[6] == 0 and abs(val - ref) < 1e-5
    for _ in range(3):
        ops.run_step(a, b, c[:, :, :2], e[:, :, :2], 0.5, engine, True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = ops.run_step(a, b, c[:, :, :2], e[:, :, :2], 0.5, engine, True, return_aux=True)
    torch.cuda.synchronize()
    for it in range(steps):
        g.replay()
        if it % 20 == 0 or it == steps - 1:
            torch.cuda.synchronize()
            st = out[3]["stats"].cpu().numpy()
            print(f"graph it {it}: loss {float(out[0]):.7f} fail_site {st[6]} flags {st[3]}", flush=True)
            assert st[6] == 0 and abs(float(out[0]) - ref) < 1e-5
    print("stress ok")


if __name__ == "__main__":
    main()
