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
    # Back-to-back steps on alternating batches with rank-dependent delays between them: the step has no closing barrier,
    # so a fast rank starts pushing batch k+1 while a slow one still finishes batch k.  Every step must still equal the
    # single-GPU result of ITS batch, and the loss-only path must agree with the loss of the gradient path.
    if transport != "nccl":
        y1, y2, k1, k2 = synth.make_batch(n, 128, 29, "uniform")
        batches = [(z1, z2, j1, j2), (y1, y2, k1, k2)]
        refs = [ops.run_step(p.to(dev), q.to(dev), r.to(dev)[:, :, :2], s_.to(dev)[:, :, :2], 0.5, "tf32", True)
                for p, q, r, s_ in batches]
        local = [(p[sl].to(dev), q[sl].to(dev), r[sl].to(dev)[:, :, :2], s_[sl].to(dev)[:, :, :2]) for p, q, r, s_ in batches]
        outs = []
        for step in range(24):
            if (step + rank) % 3 == 0:
                torch.cuda._sleep(2_000_000 * (1 + (step * 7 + rank * 3) % 4))      # ~1-4 ms of skew on this rank
            outs.append(run_step_sharded(*local[step % 2], 0.5, "tf32", True, dist.group.WORLD, transport=transport))
        loss_only, _, _ = run_step_sharded(*local[1], 0.5, "tf32", False, dist.group.WORLD, transport=transport)
        torch.cuda.synchronize()
        worst = 0.0
        for step, (l, g1, g2) in enumerate(outs):
            rl, r1, r2 = refs[step % 2]
            sc = float(r1.abs().max())
            worst = max(worst, abs(float(l) - float(rl)) / abs(float(rl)), float((g1 - r1[sl]).abs().max()) / sc * 1e-2,
                        float((g2 - r2[sl]).abs().max()) / sc * 1e-2)
        worst = max(worst, abs(float(loss_only) - float(refs[1][0])) / abs(float(refs[1][0])))
        ok2 = worst < 2e-6
        print(f"rank {rank}/{world}: 24 skewed back-to-back steps + loss-only step: worst scaled error {worst:.1e} -> "
              f"{'PASS' if ok2 else 'MISMATCH'}", flush=True)
        ok = ok and ok2
    if transport in ("auto", "fused"):
        # a rank that stays away for 3 s (checkpoint, dataloader stall): the others wait (bounded at 30 s by default) and
        # the step is still exact -- a timeout would poison the group and give NaN, never a silently wrong loss
        if rank == world - 1:
            torch.cuda._sleep(int(3.0 * 1.9e9))
        l3, g31, g32 = run_step_sharded(*local[0], 0.5, "tf32", True, dist.group.WORLD, transport=transport)
        torch.cuda.synchronize()
        ok3 = abs(float(l3) - float(refs[0][0])) <= 2e-6 * abs(float(refs[0][0]))
        # the other weightings on several ranks (fused exchange only)
        ok4 = True
        for wt, pos, neg in ((ops.make_weighting("non_linear", "mpjpe", 2.5, 0.05), True, True),
                             (ops.make_weighting("linear", "w_abs"), True, True),
                             (ops.make_weighting("linear", "mpjpe"), True, False),
                             (ops.make_weighting("linear", "mpjpe"), False, True)):
            ls, s1, s2 = run_step_sharded(*local[0], 0.5, "fp16", True, dist.group.WORLD, transport=transport,
                                          pos_weighted=pos, neg_weighted=neg, weighting=wt)
            p, q, r, s_ = batches[0]
            lf, f1, f2 = ops.run_step(p.to(dev), q.to(dev), r.to(dev)[:, :, :2], s_.to(dev)[:, :, :2], 0.5, "fp16", True,
                                      pos_weighted=pos, neg_weighted=neg, weighting=wt)
            torch.cuda.synchronize()
            sc = float(f1.abs().max())
            e = max(abs(float(ls) - float(lf)) / abs(float(lf)), float((s1 - f1[sl]).abs().max()) / sc * 1e-2,
                    float((s2 - f2[sl]).abs().max()) / sc * 1e-2)
            ok4 = ok4 and e < 2e-6
        print(f"rank {rank}/{world}: step after a 3 s stall of one rank -> {'PASS' if ok3 else 'MISMATCH'}; "
              f"other weightings sharded -> {'PASS' if ok4 else 'MISMATCH'}", flush=True)
        ok = ok and ok3 and ok4
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
