#!/usr/bin/env python
"""Fused projection head + first normalisation (smh_head.cu) against the eager composition it replaces (cuBLAS GEMMs +
ATen BatchNorm / ReLU / normalize under bf16 autocast) on the same GPU: forward and forward+backward, CUDA-event timed.
    python tools/bench_head.py [rows=16384] > gpurun_out/head_bench.json"""
import json
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200.head import FusedProjectionHead, reference_head_forward  # noqa: E402


def timed(fn, iters=50, warm=5, graph=True):
    """CUDA-event time per call.  graph=True: the call is captured once and replayed, so the host's launch path (a dozen
    Python-dispatched launches per call on either side) is not what is measured."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    run = fn
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
        for _ in range(3):
            run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    seq = nn.Sequential(nn.Linear(2048, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Linear(512, 128, bias=False)).to(dev).train()
    fused = FusedProjectionHead(seq, act_dtype=torch.bfloat16).train()
    x = (torch.relu(torch.randn(rows, 2048, device=dev)) * 0.7).to(torch.bfloat16)     # encoder output under autocast
    cot = torch.randn(rows, 128, device=dev)
    params = [p for p in seq.parameters()]

    def eager_fwd():
        with torch.no_grad():
            return reference_head_forward(seq, x)

    def fused_fwd():
        with torch.no_grad():
            return fused(x)

    def eager_fb():
        xr = x.detach().requires_grad_(True)
        y = reference_head_forward(seq, xr)
        return torch.autograd.grad((y * cot).sum(), [xr] + params)

    def fused_fb():
        xr = x.detach().requires_grad_(True)
        y = fused(xr)
        return torch.autograd.grad((y * cot).sum(), [xr] + params)

    res = dict(rows=rows, in_dim=2048, hidden=512, out_dim=128, dtype="bf16 operands, fp32 accumulate",
               timing="CUDA events over 50 replays of the captured call (CUDA graph)",
               eager_fwd_ms=timed(eager_fwd), fused_fwd_ms=timed(fused_fwd), eager_fwd_bwd_ms=timed(eager_fb),
               fused_fwd_bwd_ms=timed(fused_fb),
               eager_fwd_ms_launched_from_python=timed(eager_fwd, graph=False),
               fused_fwd_ms_launched_from_python=timed(fused_fwd, graph=False))
    res["fwd_speedup"] = res["eager_fwd_ms"] / res["fused_fwd_ms"]
    res["fwd_bwd_speedup"] = res["eager_fwd_bwd_ms"] / res["fused_fwd_bwd_ms"]
    flops = 2.0 * rows * 2048 * 512
    res["gemm1_tflops_if_fwd_were_only_gemm1"] = flops / (res["fused_fwd_ms"] * 1e-3) / 1e12
    print(json.dumps(res))


if __name__ == "__main__":
    main()
