#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): checks that the sharded step equals the single-GPU step on the
concatenated global batch.   torchrun --nproc-per-node 2 tools/dist_check.py [n_per_view]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simhand_b200 import ops, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    transport = sys.argv[2] if len(sys.argv) > 2 else "auto"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    z1, z2, j1, j2 = synth.make_batch(n, 128, 13, "hand")
    n_local = n // world
    sl = slice(rank * n_local, (rank + 1) * n_local)
    a = z1[sl].to(dev).requires_grad_(True)
    b = z2[sl].to(dev).requires_grad_(True)
    from simhand_b200.dist import run_step_sharded
    ok_all = True
    for rep in range(3):          # repeated steps exercise the barrier counters and the buffer reuse
        loss, g1, g2 = run_step_sharded(a.detach(), b.detach(), j1[sl].to(dev)[:, :, :2], j2[sl].to(dev)[:, :, :2],
                                        0.5, "tf32", True, dist.group.WORLD, transport=transport)
    a.grad, b.grad = g1, g2
    full_loss, f1, f2 = ops.run_step(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2], 0.5, "tf32", True)
    torch.cuda.synchronize()
    e_loss = abs(float(loss) - float(full_loss)) / abs(float(full_loss))
    scale = float(f1.abs().max())
    e1 = float((a.grad - f1[sl]).abs().max()) / scale
    e2 = float((b.grad - f2[sl]).abs().max()) / scale
    ok = e_loss < 2e-6 and e1 < 1e-4 and e2 < 1e-4
    print(f"rank {rank}/{world}: n={n} transport={transport} sharded loss {float(loss):.7f} single {float(full_loss):.7f} rel {e_loss:.1e} "
          f"grad err {e1:.1e} {e2:.1e} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
