"""Host-buffer front end of the fused loss step: the call a training loop makes when its batch lives in host memory.

The reference computes the loss where the dataloader left the tensors (`src/models/simhand_w_model.py:43-60` moves
the batch to the device and calls `get_weights_linear` + `vanila_weights_contrastive_loss`).  `HostPipeline` is the
same step for a caller that hands over HOST tensors: the host->device copy of batch k+1 runs on a copy stream while
the kernels of batch k run, the step itself is a CUDA graph replay over static device buffers, and the loss comes
back to pinned host memory.  Every batch is still copied exactly once, inside the caller's loop.

    pipe = HostPipeline(step_fn, example_inputs, device)        # step_fn(z1, z2, j1, j2) -> (loss, dz1, dz2)
    pipe.prefetch(z1_h, z2_h, j1_h, j2_h)                       # batch 0
    for k in range(steps):
        if k + 1 < steps: pipe.prefetch(*host_batch[k + 1])     # overlaps with the step below
        loss_h, dz1, dz2 = pipe.step()                          # returns once the loss is on the host

`lag=1` keeps one step in flight: `step()` enqueues step k (copy wait, kernels, loss read-back) and returns the HOST loss
of step k - 1 (None for the first call), which is what a logging training loop needs; `drain()` returns the last one.
The host then never idles the device between steps (with lag=0 every step ends with a device->host round trip before
the next one can be enqueued: 6 % of a step on one GPU, 16 % on eight).

dz1 / dz2 are device tensors owned by the slot that ran: they are valid until that slot runs again (`depth` steps
later), which is when a training loop has long consumed them.  No CPU fallback: CUDA only.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class HostPipeline:
    def __init__(self, step_fn: Callable, example_inputs: Sequence[torch.Tensor], device: torch.device,
                 depth: int = 2, use_graph: bool = True, sync_all: Callable[[], None] | None = None, lag: int = 0):
        if device.type != "cuda":
            raise RuntimeError("simhand_b200.HostPipeline needs a CUDA device")
        if lag not in (0, 1):
            raise ValueError("lag must be 0 (every step returns its own loss) or 1 (one step in flight)")
        if lag and depth < 2:
            raise ValueError("lag=1 needs depth >= 2")
        self.device, self.depth, self.step_fn, self.lag = device, depth, step_fn, lag
        self.host_losses = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.inflight = None                                     # slot of the step whose loss has not been returned yet
        self.copy_stream = torch.cuda.Stream(device)
        self.slots = [[torch.empty_like(t, device=device) for t in example_inputs] for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.used = [False] * depth
        self.host_loss = torch.empty((), dtype=torch.float32).pin_memory()
        self.head = self.tail = self.pending = 0
        self.graphs, self.outs = [None] * depth, [None] * depth
        sync = sync_all or (lambda: torch.cuda.synchronize(device))
        if use_graph:
            for s in range(depth):
                for t, e in zip(self.slots[s], example_inputs):
                    t.copy_(e)
                for _ in range(2):
                    step_fn(*self.slots[s])
                sync()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.outs[s] = step_fn(*self.slots[s])
                sync()
                self.graphs[s] = g

    def prefetch(self, *host_inputs: torch.Tensor) -> None:
        """Enqueue the host->device copy of one batch (pinned host tensors copy asynchronously)."""
        if self.pending == self.depth:
            raise RuntimeError("HostPipeline: every slot holds a batch that has not been stepped yet")
        s = self.tail
        if self.used[s]:
            self.copy_stream.wait_event(self.free[s])           # the step that read this slot has finished
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.slots[s], host_inputs):
                dst.copy_(src, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self.tail = (s + 1) % self.depth
        self.pending += 1

    def step(self):
        """Run the step on the oldest prefetched batch; returns (loss on the host, dz1, dz2) after the loss landed."""
        if self.pending == 0:
            raise RuntimeError("HostPipeline.step() without a prefetched batch")
        s = self.head
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.ready[s])
        if self.graphs[s] is not None:
            self.graphs[s].replay()
            loss, dz1, dz2 = self.outs[s]
        else:
            loss, dz1, dz2 = self.step_fn(*self.slots[s])
        self.host_losses[s].copy_(loss, non_blocking=True)
        self.free[s].record(main)
        self.done[s].record(main)
        self.used[s] = True
        self.head = (s + 1) % self.depth
        self.pending -= 1
        if self.lag == 0:
            self.done[s].synchronize()
            self.host_loss = self.host_losses[s]
            return self.host_loss, dz1, dz2
        prev, self.inflight = self.inflight, (s, dz1, dz2)
        if prev is None:
            return None, None, None
        self.done[prev[0]].synchronize()                        # the PREVIOUS step's loss is on the host; this one runs on
        return self.host_losses[prev[0]], prev[1], prev[2]

    def drain(self):
        """lag=1: waits for the step still in flight and returns its (host loss, dz1, dz2)."""
        if self.inflight is None:
            return None, None, None
        s, dz1, dz2 = self.inflight
        self.inflight = None
        self.done[s].synchronize()
        return self.host_losses[s], dz1, dz2
