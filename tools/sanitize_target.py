#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer (tools/sanitize.sh): the fused step under each engine and
both distance images, the weights API, a variant weighting, the materialised-weights path, the transform, the fused sharded
step with 2 emulated ranks, and the projection head forward + backward.  Sizes are tiny: the tools serialise everything."""
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import ops, synth  # noqa: E402
from simhand_b200.dist import EmulatedGroup  # noqa: E402
from simhand_b200.head import FusedProjectionHead  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = torch.device("cuda:0")
    n = 192
    z1, z2, j1, j2 = synth.make_batch(n, 128, 5, "hand")
    z1, z2, a, b = z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2]
    if what in ("all", "step"):
        for engine, exact in (("fp16", False), ("fp16", True), ("tf32", True), ("bf16", False), ("fp32", True)):
            loss, g1, g2 = ops.run_step(z1, z2, a, b, 0.5, engine, True, exact_weights=exact)
            print(f"step {engine} exact={exact}: loss {float(loss):.6f}")
        pw, nw = ops.mpjpe_weights(a, b)
        wt = ops.make_weighting("non_linear", "w_abs", 2.5, 0.05)
        loss, _, _ = ops.run_step(z1, z2, a, b, 0.5, "fp16", True, weighting=wt)
        loss, _, _ = ops.run_step_dense(z1, z2, pw, nw, 0.5, "fp16", True)
        x = torch.randn(2 * n, 128, device=dev, requires_grad=True)
        y = ops.get_transformed_projections(x, torch.rand(2 * n, device=dev) * 0.1, torch.rand(2 * n, device=dev) * 0.1,
                                            torch.rand(2 * n, device=dev) * 30)
        y.sum().backward()
        print("weights / variant / dense / transform done")
    if what in ("all", "shard"):
        grp = EmulatedGroup(n, 128, 2, dev, "fp16")
        for _ in range(2):
            losses, _, _ = grp.step(z1, z2, a, b)
        torch.cuda.synchronize()
        print("emulated 2 ranks:", [float(x) for x in losses], grp.poisoned())
    if what in ("all", "head"):
        torch.manual_seed(0)
        seq = nn.Sequential(nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Linear(256, 128, bias=False)).to(dev)
        head = FusedProjectionHead(seq, act_dtype=torch.bfloat16).train()
        xin = torch.randn(300, 512, device=dev, requires_grad=True)
        out = head(xin)
        out.sum().backward()
        print("head:", float(out.norm()))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
