#!/usr/bin/env python
"""End-to-end pre-training step with the fused loss (BASELINE.json configs[4]).

What the reference runs per step under `strategy="dp"` (src/experiments/main.py:152-163; HandCLR_W.training_step,
simhand_w_model.py:122-152) rebuilt as one process per GPU:

    images [2B_local, 3, 128, 128]  -> ResNet-50 features (2048) -> projection head (simclr_model.py:22-39)
      -> simhand_b200.get_transformed_projections   (normalise, translate by -jitter, rotate by -angle, normalise)
      -> simhand_b200.weighted_ntxent(..., group)   (the GLOBAL-batch loss: all ranks' samples are negatives)
      -> backward -> gradient all-reduce of the backbone (DistributedDataParallel) -> SGD step

The reference's DataParallel computes 8 independent local losses; here the loss is the reference
function applied to the concatenated global batch.  The loss returned by the op is identical on every rank and its
autograd hands each rank d(loss)/d(z_local), so the DDP average of the parameter gradients would be 1/world of the true
gradient: the op is called with `grad_scale=world`.

`--fused-head` swaps the head for simhand_b200.FusedProjectionHead (tcgen05 GEMMs with BatchNorm / ReLU / L2-normalise
fused; same parameters).  `--dump FILE` saves the first step's projections, joints, loss and d loss / d projections of every
rank (tests/test_gpu_e2e.py holds them against the CPU oracle).

Synthetic images and joints, random-init weights (no network in this environment).  Prints one line per rank-0 with the
step time and the share of the loss (transform + loss forward/backward) in it.

    python examples/e2e_step.py --batch 512                              # one GPU, 512 samples per view
    torchrun --nproc-per-node 8 examples/e2e_step.py --batch 8192        # global batch 8192 over 8 GPUs
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simhand_b200  # noqa: E402
from simhand_b200 import synth  # noqa: E402


def build_model(out_dim: int = 128, fused_head: bool = False):
    import torchvision
    backbone = torchvision.models.resnet50(weights=None)
    backbone.fc = nn.Identity()
    head = nn.Sequential(nn.Linear(2048, 512, bias=True), nn.BatchNorm1d(512), nn.ReLU(),
                         nn.Linear(512, out_dim, bias=False))          # simclr_model.py:22-39
    if fused_head:
        head = simhand_b200.FusedProjectionHead(head, act_dtype=torch.bfloat16)
    return nn.Sequential(backbone, head)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512, help="GLOBAL samples per view")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--image", type=int, default=128)
    ap.add_argument("--fused-head", action="store_true")
    ap.add_argument("--dump", default=None, help="save the first step's loss inputs / outputs of every rank to FILE.rank")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    b = args.batch // world
    torch.manual_seed(1234 + rank)
    model = build_model(fused_head=args.fused_head).to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index])
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-6)
    images = torch.randn(2 * b, 3, args.image, args.image, device=dev).contiguous(memory_format=torch.channels_last)
    _, _, j1, j2 = synth.make_batch(args.batch, 128, 7, "peclr")
    sl = slice(rank * b, (rank + 1) * b)
    joints1, joints2 = j1[sl].to(dev)[:, :, :2], j2[sl].to(dev)[:, :, :2]
    gen = torch.Generator().manual_seed(99 + rank)
    jitter = (torch.randint(0, 16, (2, 2 * b), generator=gen).float() / args.image).to(dev)
    angles = torch.randint(-45, 46, (2 * b,), generator=gen).float().to(dev)

    def step(timers=None, dump=None):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            proj = model(images)                                        # [2b, 128]
        if timers:
            timers[0].record()
        p = simhand_b200.get_transformed_projections(proj.float(), -jitter[0], -jitter[1], -angles)
        if dump:
            p.retain_grad()
        loss = simhand_b200.weighted_ntxent(p[:b], p[b:], joints1, joints2, 0.5, group, grad_scale=float(world))
        if timers:
            timers[1].record()
        loss.backward()
        if dump:
            torch.save(dict(p=p.detach().cpu(), dp=p.grad.cpu(), joints1=joints1.cpu(), joints2=joints2.cpu(),
                            loss=float(loss), world=world, rank=rank), f"{dump}.{rank}")
        opt.step()
        return loss.detach()

    losses = []
    for it in range(args.warmup):
        losses.append(float(step(dump=args.dump if it == 0 else None)))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev[2].record()
    t_loss = 0.0
    for _ in range(args.steps):
        losses.append(float(step(ev)))          # float(): the per-step host read the reference's logging does
        torch.cuda.synchronize(dev)
        t_loss += ev[0].elapsed_time(ev[1])
    ev[3].record()
    torch.cuda.synchronize(dev)
    ms = ev[2].elapsed_time(ev[3]) / args.steps
    finite = all(x == x and abs(x) < 1e9 for x in losses)
    if rank == 0:
        print(json.dumps(dict(example="e2e_step", world=world, global_batch=args.batch, local_batch=b, image=args.image,
                              ms_per_step=ms, samples_per_s=2 * args.batch / (ms * 1e-3),
                              loss_forward_ms=t_loss / args.steps, losses=[round(x, 5) for x in losses],
                              loss_decreasing=bool(losses[-1] < losses[0]), finite=finite)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if finite else 1)


if __name__ == "__main__":
    main()
