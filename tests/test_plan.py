"""CPU tests of the host side of the C ABI: library loads, exports every declared symbol, argument
validation, and the task plan / layout arithmetic (no kernel launches)."""
import ctypes
import re
import os

import numpy as np
import pytest

from simhand_b200 import _lib, layouts as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "simhand_b200.h")).read()
    declared = set(re.findall(r"\b(smh_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.smh_version() == 100


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_lib.Dims) == 40
    assert ctypes.sizeof(_lib.Stats) == 48
    assert ctypes.sizeof(_lib.Exchange) == 8 + 3 * 8 * 8 + 8
    assert ctypes.sizeof(_lib.Layout) == 16 * 8 + 8 * 4
    assert ctypes.sizeof(_lib.Inputs) == 88


@pytest.mark.parametrize("args,code", [((0, 128, 1, 0), -1), ((8, 129, 1, 0), -2), ((8, 0, 1, 0), -2),
                                       ((8, 128, 2, 2), -2), ((9, 128, 2, 0), -2)])
def test_layout_rejects_bad_dims(args, code):
    lib = _lib.load()
    dims, lay = _lib.Dims(*args, 0), _lib.Layout()
    assert lib.smh_layout(ctypes.byref(dims), ctypes.byref(lay)) == code
    assert lib.smh_last_error()


def test_compute_calls_reject_null_workspace_without_gpu():
    lib = _lib.load()
    dims = _lib.Dims(8, 128, 1, 0, 0)
    assert lib.smh_mpjpe(ctypes.byref(dims), None, None, None, None) == -1
    assert lib.smh_forward(ctypes.byref(dims), None, None, 0.5, 0, None, None) == -1
    assert lib.smh_barrier(None, None) == -1 and lib.smh_push_inputs(None, None, 0, 0, None) == -1
    assert lib.smh_l2norm_fwd(None, None, None, 4, 4, 1e-12, None) == -1


@pytest.mark.parametrize("n,world", [(3, 1), (64, 1), (96, 1), (200, 1), (256, 1), (256, 2), (1024, 8),
                                     (330, 2), (8192, 8)])
def test_plan_covers_every_ordered_pair_once(n, world):
    m = 2 * n
    tp = (m + 127) // 128
    cover = np.zeros((tp, 2 * tp), np.int32)          # (row block, 64-col tile)
    stored = np.zeros((tp, tp), np.int32)
    per_rank = []
    for rank in range(world):
        lay, (h, tiles, tasks, strips) = L.build_plan(n, 128, world, rank)
        assert h["magic"] == 0x534D4831 and h["m"] == m
        per_rank.append(len(tiles))
        for I, J in tiles:
            assert I <= J
            stored[I, J] += 1
        for row, cj, lt, flags in tasks:
            I, J = tiles[lt]
            if flags & L.TASK_TRANSPOSED:
                assert row == J and cj // 2 == I and I != J
            else:
                assert row == I and cj // 2 == J
                assert bool(flags & L.TASK_DIAGONAL) == (I == J)
            ragged = (row * 128 + 128 > m) or (cj * 64 + 64 > m)
            assert bool(flags & L.TASK_RAGGED) == ragged
            cover[row, cj] += 1
        # both sweeps cut the same task list: strips partition it and never mix row blocks; the sweep CTAs own contiguous
        # task ranges of equal COST (task = 1, every strip a CTA opens = 4 task times backward, 0.5 forward)
        assert lay.n_strips == len(strips) and lay.n_strips_fwd == len(h["strips_fwd"])
        for table, cp, strip_cost in ((strips, h["cta_ptr"], 4.0), (h["strips_fwd"], h["cta_ptr_fwd"], 0.5)):
            order = np.argsort(table[:, 0])
            ss = table[order]
            assert ss[0, 0] == 0 and ss[-1, 1] == len(tasks)
            assert (ss[1:, 0] == ss[:-1, 1]).all()
            for a, b in table:
                assert 0 < b - a <= lay.strip_len
                assert len(set(tasks[a:b, 0])) == 1
            assert cp[0] == 0 and cp[-1] == len(table) and (np.diff(cp) >= 0).all()
            per_cta = [int(table[cp[c]:cp[c + 1], 1].max() - table[cp[c]:cp[c + 1], 0].min()) if cp[c + 1] > cp[c] else 0
                       for c in range(L.NUM_CTAS)]
            assert sum(per_cta) == len(tasks)
            cost = [per_cta[c] + strip_cost * (cp[c + 1] - cp[c]) for c in range(L.NUM_CTAS) if per_cta[c]]
            if len(tasks) >= 4 * L.NUM_CTAS:
                assert max(cost) - min(cost) <= 1.5 * strip_cost + 1.0 + 0.02 * max(cost), (min(cost), max(cost))
            assert len(cost) == min(L.NUM_CTAS, len(tasks))
    assert (stored[np.triu_indices(tp)] == 1).all() and stored.sum() == tp * (tp + 1) // 2
    live = np.array([[cj * 64 < m for cj in range(2 * tp)]] * tp)
    assert (cover[live] == 1).all() and (cover[~live] == 0).all()
    if tp >= 2 * world:
        assert max(per_rank) - min(per_rank) <= tp + 1      # balanced over ranks


@pytest.mark.parametrize("n", [3, 96, 200, 1024])
def test_dense_plans_visit_every_tile(n):
    """Materialised-weights path: every (I, J) tile is stored; the forward list visits each 128 x 64 piece once
    (direct), the backward list once direct and once through the transposed read of the mirror tile."""
    m = 2 * n
    tp = (m + 127) // 128
    lay_f, (hf, tiles_f, tasks_f, _) = L.build_plan(n, 128, 1, 0, 0, _lib.DIMS_DENSE_WEIGHTS)
    lay_b, (hb, tiles_b, tasks_b, strips_b) = L.build_plan(n, 128, 1, 0, 0,
                                                           _lib.DIMS_DENSE_WEIGHTS | _lib.DIMS_DENSE_BACKWARD)
    assert lay_f.ws_bytes == lay_b.ws_bytes and lay_f.off_dist == lay_b.off_dist
    assert (tiles_f == tiles_b).all() and len(tiles_f) == tp * tp
    assert [tuple(t) for t in tiles_f] == [(i, j) for i in range(tp) for j in range(tp)]
    live = np.array([[cj * 64 < m for cj in range(2 * tp)]] * tp)
    for tasks, want_t in ((tasks_f, 0), (tasks_b, 1)):
        direct = np.zeros((tp, 2 * tp), np.int32)
        trans = np.zeros((tp, 2 * tp), np.int32)
        for row, cj, lt, flags in tasks:
            I, J = tiles_f[lt]
            if flags & L.TASK_TRANSPOSED:
                assert (I, J) == (cj // 2, row)
                trans[row, cj] += 1
            else:
                assert (I, J) == (row, cj // 2)
                direct[row, cj] += 1
            assert bool(flags & L.TASK_DIAGONAL) == (row == cj // 2)
        assert (direct[live] == 1).all() and (direct[~live] == 0).all()
        assert (trans[live] == want_t).all() and (trans[~live] == 0).all()
    for a, b in strips_b:
        assert len(set(tasks_b[a:b, 0])) == 1
    # dims validation of the path
    lib = _lib.load()
    lay = _lib.Layout()
    assert lib.smh_layout(ctypes.byref(_lib.Dims(n, 128, 1, 0, 0, _lib.DIMS_DENSE_BACKWARD)), ctypes.byref(lay)) == -1
    if n % 2 == 0:
        assert lib.smh_layout(ctypes.byref(_lib.Dims(n, 128, 2, 0, 0, _lib.DIMS_DENSE_WEIGHTS)), ctypes.byref(lay)) == -2


def test_layout_indices_are_bijections():
    r, c = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    assert sorted(L.dist_index(r, c).ravel()) == list(range(128 * 128))
    r, c = np.meshgrid(np.arange(192), np.arange(128), indexing="ij")
    assert sorted(L.zt_index(r, c).ravel()) == list(range(192 * 128))
    # a 16-byte chunk stays together and lands in the XOR-swizzled slot of its 128-byte row
    assert L.zt_index(5, 8) == 5 * 32 + ((2 ^ 5) * 4)
    assert sorted(L.jp_index(j, c) for j in range(21) for c in range(2)) == list(range(40)) + [40, 41]


def test_staged_tile_reads_match_matrix():
    """Producer piece mapping + epilogue reads (smh_sweep_tc.cu) reproduce D[gi, gj] for every task."""
    n = 200
    m = 2 * n
    rng = np.random.default_rng(0)
    dfull = rng.random((512, 512)).astype(np.float32)
    dfull = np.triu(dfull) + np.triu(dfull, 1).T             # symmetric like the MPJPE matrix
    lay, (h, tiles, tasks, strips) = L.build_plan(n)
    r, c = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    store = np.zeros((len(tiles), L.TILE_FLOATS), np.float32)
    for lt, (I, J) in enumerate(tiles):
        store[lt, L.dist_index(r, c)] = dfull[I * 128:(I + 1) * 128, J * 128:(J + 1) * 128]
    for task in tasks:
        row, cj, lt, flags = task
        stage = L.stage_task(store[lt], task)
        assert stage.size == L.STAGE_FLOATS
        for rr in (0, 1, 5, 63, 64, 77, 127):
            for jl in (0, 3, 4, 31, 32, 63):
                assert L.staged_read(stage, task, rr, jl) == dfull[row * 128 + rr, cj * 64 + jl]


def test_staged_reads_are_bank_conflict_free():
    """32 lanes of an epilogue warp hit 32 distinct banks (transposed, 4-byte reads) or 8 distinct 16-byte slots per
    quarter warp (direct, 16-byte reads) for every column."""
    for w4 in range(4):
        rows = np.arange(32) + 32 * w4
        for jl in range(64):
            c4 = rows >> 2
            word = (c4 * 64 + (jl ^ (c4 & 7))) * 4 + (rows & 3)
            assert len(set(word % 32)) == 32
            c4l = jl >> 2
            slot = ((rows >> 6) * 16 + c4l) * 64 + ((rows & 63) ^ (c4l & 7))
            for q in range(4):
                assert len(set(slot[8 * q:8 * q + 8] % 8)) == 8


def test_q16_tile_layout_mirrors():
    """16-bit tile image: index bijection, producer piece mapping + epilogue reads reproduce q[gi, gj] for every task,
    and both read patterns are free of shared-memory bank conflicts."""
    r, c = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    assert sorted(L.distq_index(r, c).ravel()) == list(range(128 * 128))
    n = 200
    rng = np.random.default_rng(1)
    qfull = rng.integers(0, 65000, (512, 512)).astype(np.uint16)
    qfull = np.triu(qfull) + np.triu(qfull, 1).T
    lay, (h, tiles, tasks, strips) = L.build_plan(n, 128, 1, 0, 0, _lib.DIMS_Q16_TILES)
    lay32, _ = L.build_plan(n)
    assert lay.ws_bytes - lay.off_dist == (lay32.ws_bytes - lay32.off_dist) // 2          # half the tile bytes
    store = np.zeros((len(tiles), 128 * 128), np.uint16)
    for lt, (I, J) in enumerate(tiles):
        store[lt, L.distq_index(r, c)] = qfull[I * 128:(I + 1) * 128, J * 128:(J + 1) * 128]
    for task in tasks:
        row, cj, lt, flags = task
        stage = L.stage_task_q16(store[lt], task)
        assert stage.size == L.STAGE_Q16
        for rr in (0, 1, 7, 8, 63, 64, 77, 127):
            for jl in (0, 3, 4, 7, 8, 31, 32, 63):
                assert L.staged_read_q16(stage, task, rr, jl) == qfull[row * 128 + rr, cj * 64 + jl]
    for w4 in range(4):
        rows = np.arange(32) + 32 * w4
        for jl in range(64):
            # transposed: 2-byte reads; lanes may share a 32-bit word (broadcast) but never a bank with another word
            c8 = rows >> 3
            byte = (c8 * 64 + (jl ^ (c8 & 7))) * 16 + (rows & 7) * 2
            words = byte // 4
            assert len(set(words % 32)) == len(set(words))
        for c8l in range(8):
            # direct: 16-byte reads, conflict-free within every quarter warp
            slot = (rows & 63) ^ (c8l & 7)
            for qw in range(4):
                assert len(set((slot[8 * qw:8 * qw + 8] * 4) % 32)) == 8
    # the flag is rejected where the fp32 tiles are required
    lib = _lib.load()
    lo = _lib.Layout()
    assert lib.smh_layout(ctypes.byref(_lib.Dims(64, 128, 1, 0, 0, _lib.DIMS_Q16_TILES, 1)), ctypes.byref(lo)) == -6
    assert lib.smh_layout(ctypes.byref(_lib.Dims(64, 128, 1, 0, 0, _lib.DIMS_Q16_TILES | _lib.DIMS_DENSE_WEIGHTS)),
                          ctypes.byref(lo)) == -6


def test_plan_properties_random_shapes():
    """Seeded sweep over odd shapes: every ordered 128 x 64 piece is covered exactly once across the ranks, strips never
    mix row blocks, and the 16-bit tile flag changes the layout size but not the task list."""
    rng = np.random.default_rng(7)
    for _ in range(12):
        world = int(rng.choice([1, 2, 4, 8]))
        n = int(rng.integers(1, 90)) * world
        m = 2 * n
        tp = (m + 127) // 128
        cover = np.zeros((tp, 2 * tp), np.int32)
        for rank in range(world):
            lay, (h, tiles, tasks, strips) = L.build_plan(n, 128, world, rank)
            layq, (hq, tiles_q, tasks_q, strips_q) = L.build_plan(n, 128, world, rank, 0, _lib.DIMS_Q16_TILES)
            assert layq.n_tasks == lay.n_tasks and layq.n_stored_tiles == lay.n_stored_tiles
            assert sorted(map(tuple, tasks_q[:, :3])) == sorted(map(tuple, tasks[:, :3]))
            assert lay.ws_bytes - lay.off_dist == 2 * (layq.ws_bytes - layq.off_dist)
            for row, cj, lt, flags in tasks:
                cover[row, cj] += 1
            for a, b in strips:
                assert len(set(tasks[a:b, 0])) == 1
        live = np.array([[cj * 64 < m for cj in range(2 * tp)]] * tp)
        assert (cover[live] == 1).all() and (cover[~live] == 0).all(), (n, world)
