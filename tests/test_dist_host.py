"""CPU tests (gloo, world_size 2) of the host-side logic of the sharded path (simhand_b200/dist.py): the packed
all-gather layout and its stride description, the rank-major gradient rows the reduce-scatter relies on, and the
order-preserving integer images combined with all_reduce(MAX).  No kernels are launched."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simhand_b200 import dist as sd
from simhand_b200 import layouts as L
from simhand_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_local, d, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = n_local * world
        z1, z2, j1, j2 = synth.make_batch(n, d, 7, "uniform")
        sl = slice(rank * n_local, (rank + 1) * n_local)
        local = sd.pack_local(z1[sl], z2[sl], j1[sl][:, :, :2], j2[sl][:, :, :2])
        gathered = torch.empty(world * local.numel())
        dist.all_gather_into_tensor(gathered, local)
        (o1, o2, oj1, oj2), chunk = sd.gathered_views(gathered, world, n_local, d)
        ok = True
        for k in range(n):
            a = o1 + sd.sample_offset(k, n_local, chunk, d)
            b = o2 + sd.sample_offset(k, n_local, chunk, d)
            ok &= torch.equal(gathered[a:a + d], z1[k]) and torch.equal(gathered[b:b + d], z2[k])
            ja = oj1 + sd.sample_offset(k, n_local, chunk, 42)
            jb = oj2 + sd.sample_offset(k, n_local, chunk, 42)
            ok &= torch.equal(gathered[ja:ja + 42], j1[k, :, :2].reshape(-1))
            ok &= torch.equal(gathered[jb:jb + 42], j2[k, :, :2].reshape(-1))
        # rank-major gradient rows: reduce-scatter chunk r == [view-1 rows of rank r; view-2 rows of rank r]
        full = torch.zeros(2 * n, 4)
        for i in range(2 * n):
            full[sd.dz_out_row(i, n, n_local)] = float(i)
        mine = torch.empty(2 * n_local, 4)
        dist.reduce_scatter_tensor(mine, full.clone(), op=dist.ReduceOp.SUM)
        want = torch.cat([torch.arange(rank * n_local, (rank + 1) * n_local),
                          n + torch.arange(rank * n_local, (rank + 1) * n_local)]).float() * world
        ok &= torch.equal(mine[:, 0], want)
        # order-preserving images: max over ranks of (Dmax bits, Pmax bits, 0x7fffffff - Pmin bits)
        vals = np.array([3.5 + rank, 2.0 - rank, 0.25 + 0.5 * rank], np.float32)
        img = vals.view(np.uint32).astype(np.int64)
        img[2] = 0x7FFFFFFF - img[2]
        t = torch.from_numpy(img.astype(np.int32))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        got = t.numpy().astype(np.int64)
        dmax = np.array([got[0]], np.uint32).view(np.float32)[0]
        pmin = np.array([0x7FFFFFFF - got[2]], np.uint32).view(np.float32)[0]
        ok &= (dmax == 3.5 + world - 1) and (pmin == 0.25)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_packed_gather_and_scatter_layout_world2():
    world, n_local, d = 2, 24, 16
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_local, d, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_dz_out_row_matches_library_convention():
    n, n_local = 12, 4
    rows = [sd.dz_out_row(i, n, n_local) for i in range(2 * n)]
    assert sorted(rows) == list(range(2 * n))
    # rank r owns rows [r * 2 n_local, (r + 1) * 2 n_local): first its view-1 samples, then its view-2 samples
    for r in range(n // n_local):
        blk = [i for i in range(2 * n) if r * 2 * n_local <= sd.dz_out_row(i, n, n_local) < (r + 1) * 2 * n_local]
        assert blk == list(range(r * n_local, (r + 1) * n_local)) + list(range(n + r * n_local, n + (r + 1) * n_local))


def test_rank_plans_partition_the_work():
    """Ranks' plans are disjoint and cover everything (complements tests/test_plan.py) and are balanced."""
    n, world = 8192, 8
    counts = [L.build_plan(n, 128, world, r)[0].n_stored_tiles for r in range(world)]
    assert sum(counts) == 128 * 129 // 2 and max(counts) == min(counts) == 1032


def test_host_pipeline_and_step_flags_without_gpu(monkeypatch):
    """Host-side behaviour that needs no device: the pipeline refuses a CPU device, and the tile-image flag follows the
    engine / weighting / environment."""
    import torch
    from simhand_b200 import _lib, ops
    from simhand_b200.pipeline import HostPipeline
    with pytest.raises(RuntimeError):
        HostPipeline(lambda *a: None, (torch.zeros(2, 2),), torch.device("cpu"))
    monkeypatch.delenv("SMH_Q16", raising=False)
    assert ops.step_flags("fp16") == _lib.DIMS_Q16_TILES and ops.step_flags("bf16") == _lib.DIMS_Q16_TILES
    assert ops.step_flags("fp32") == 0 and ops.step_flags("fp16", neg_weighted=False) == 0
    assert ops.step_flags("fp16", ops.make_weighting("non_linear", "mpjpe", 1.0, 0.05)) == 0
    assert ops.step_flags("fp16", ops.make_weighting("linear", "w_abs")) == 0
    monkeypatch.setenv("SMH_Q16", "0")
    assert ops.step_flags("fp16") == 0
    with pytest.raises(ValueError):
        ops.make_weighting("cubic", "mpjpe")
    with pytest.raises(ValueError):
        ops.make_weighting("linear", "l1")


def test_bench_reads_ncu_traffic_of_the_right_instantiation():
    """bench.py's roofline.traffic comes from the committed ncu summaries: the 16-bit-image kernel and the exact kernel are
    told apart by their template argument, whatever the files' times are after a checkout."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("smh_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    relaxed, exact = bench._ncu_traffic("mpjpe_kernel<1,"), bench._ncu_traffic("mpjpe_kernel<0,")
    assert relaxed and exact
    assert relaxed["source"].endswith("r02_ncu_summary.txt") and exact["source"].endswith("r02_ncu_summary_exact.txt")
    # 8256 tiles of 32 KiB resp. 64 KiB written once; part of it is still in the L2 when the kernel ends
    assert 0.6 * 8256 * 32768 < relaxed["bytes"] < 1.1 * 8256 * 32768
    assert 0.6 * 8256 * 65536 < exact["bytes"] < 1.1 * 8256 * 65536
