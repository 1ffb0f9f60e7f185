"""Host-side (numpy) mirrors of the HBM layouts and the task plan of libsimhand_b200.so.

Used by the CPU tests to check the plan (every ordered pair of samples is covered exactly once, ranks are
balanced) and the index arithmetic shared by the kernels (`smh_common.cuh`), and handy when inspecting a
workspace dumped from the GPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

TILE = 128
NUM_CTAS = 148                   # persistent sweep CTAs the plan is cut for
TASK_N = 64
TILE_FLOATS = TILE * TILE
BLOCK_ROWS = 64
BLOCK_FLOATS = BLOCK_ROWS * 128
JP = 44
STAGE_FLOATS = 8192              # floats of a stored tile staged per task (32 KiB, no padding)

TASK_TRANSPOSED, TASK_DIAGONAL, TASK_RAGGED = 1, 2, 4


def zt_index(row, col):
    """float index of z[row, col] in the pre-swizzled 64-row block image (SWIZZLE_128B, K-major)."""
    row, col = np.asarray(row), np.asarray(col)
    blk, r = row // BLOCK_ROWS, row % BLOCK_ROWS
    kb, cc = col >> 5, col & 31
    chunk = (cc >> 2) ^ (r & 7)
    return blk * BLOCK_FLOATS + kb * (BLOCK_ROWS * 32) + r * 32 + chunk * 4 + (cc & 3)


def dist_index(row, col):
    """float index of D[row, col] inside a stored 128x128 tile: [row/64][col/4][(row%64) ^ (col/4 % 8)][col%4]."""
    row, col = np.asarray(row), np.asarray(col)
    c4 = col >> 2
    return ((((row >> 6) * 32 + c4) * 64) + ((row & 63) ^ (c4 & 7))) * 4 + (col & 3)


def jp_index(joint, coord):
    """float index of (joint, coord) in a packed 44-float joint row."""
    if joint == 20:
        return 40 + coord
    return 4 * (joint // 2) + 2 * coord + (joint % 2)


def parse_plan(plan_bytes: np.ndarray):
    """Decodes a plan blob built by smh_plan_build into (header dict, tiles[n,2], tasks[n,4], strips[n,2])."""
    hdr = np.frombuffer(plan_bytes[:64].tobytes(), dtype=np.uint32)
    names = ("magic", "m", "world", "rank", "tiles_per_side", "n_stored", "n_tasks", "n_strips", "strip_len",
             "off_tiles", "off_tasks", "off_strips", "off_cta", "off_strips_fwd", "off_cta_fwd", "n_strips_fwd")
    h = {k: int(v) for k, v in zip(names, hdr)}
    raw = plan_bytes.tobytes()
    tiles = np.frombuffer(raw, np.int32, h["n_stored"] * 2, h["off_tiles"]).reshape(-1, 2)
    tasks = np.frombuffer(raw, np.int32, h["n_tasks"] * 4, h["off_tasks"]).reshape(-1, 4)
    strips = np.frombuffer(raw, np.int32, h["n_strips"] * 2, h["off_strips"]).reshape(-1, 2)
    h["cta_ptr"] = np.frombuffer(raw, np.int32, NUM_CTAS + 1, h["off_cta"])
    # the forward sweep's own cuts of the same task list
    h["strips_fwd"] = np.frombuffer(raw, np.int32, h["n_strips_fwd"] * 2, h["off_strips_fwd"]).reshape(-1, 2)
    h["cta_ptr_fwd"] = np.frombuffer(raw, np.int32, NUM_CTAS + 1, h["off_cta_fwd"])
    return h, tiles, tasks, strips


def build_plan(n: int, d: int = 128, world: int = 1, rank: int = 0, strip_len: int = 0, flags: int = 0):
    from . import _lib
    lib = _lib.load()
    dims = _lib.Dims(n, d, world, rank, strip_len, flags)
    lay = _lib.Layout()
    _lib.check(lib.smh_layout(ctypes.byref(dims), ctypes.byref(lay)), "smh_layout")
    buf = np.zeros(int(lay.plan_bytes), np.uint8)
    _lib.check(lib.smh_plan_build(ctypes.byref(dims), buf.ctypes.data, buf.size), "smh_plan_build")
    return lay, parse_plan(buf)


def stage_task(tile: np.ndarray, task) -> np.ndarray:
    """The 32 KiB the tile producer stages for a task (smh_sweep_tc.cu): direct -> the two 16 KiB column-half
    slabs (rows 0..63, rows 64..127); transposed -> the 32 KiB row-half slab."""
    half = task[1] & 1
    if task[3] & TASK_TRANSPOSED:
        return tile[half * 8192:(half + 1) * 8192].copy()
    lo = tile[(half * 16) * 256:(half * 16 + 16) * 256]
    hi = tile[(32 + half * 16) * 256:(32 + half * 16 + 16) * 256]
    return np.concatenate([lo, hi])


def staged_read(stage: np.ndarray, task, r: int, jl: int) -> float:
    """The value the epilogue thread of row r reads for task column jl from the staged 8192 floats."""
    if task[3] & TASK_TRANSPOSED:
        c4 = r >> 2
        return stage[(c4 * 64 + (jl ^ (c4 & 7))) * 4 + (r & 3)]
    c4l = jl >> 2
    return stage[(((r >> 6) * 16 + c4l) * 64 + ((r & 63) ^ (c4l & 7))) * 4 + (jl & 3)]


# ----------------------------------------------------------------------------------------------------
# 16-bit image of the distance tiles (SMH_DIMS_Q16_TILES): mirrors of smh_common.cuh / smh_sweep_tc.cu
# ----------------------------------------------------------------------------------------------------
STAGE_Q16 = 8192                 # u16 elements staged per task (16 KiB)
Q16_LEVELS = 65000.0


def distq_index(row, col):
    """u16 index of q[row, col] inside a stored tile: [row/64][col/8][(row%64) ^ (col/8 % 8)][col%8]."""
    row, col = np.asarray(row), np.asarray(col)
    c8 = col >> 3
    return ((((row >> 6) * 16 + c8) * 64) + ((row & 63) ^ (c8 & 7))) * 8 + (col & 7)


def stage_task_q16(tile: np.ndarray, task) -> np.ndarray:
    """The 16 KiB the tile producer stages for a task from the 16-bit image: direct -> two 8 KiB column-half slabs,
    transposed -> the 16 KiB row-half slab."""
    half = task[1] & 1
    if task[3] & TASK_TRANSPOSED:
        return tile[half * 8192:(half + 1) * 8192].copy()
    lo = tile[half * 4096:(half + 1) * 4096]
    hi = tile[8192 + half * 4096:8192 + (half + 1) * 4096]
    return np.concatenate([lo, hi])


def staged_read_q16(stage: np.ndarray, task, r: int, jl: int):
    """The value the epilogue thread of row r reads for task column jl from the staged 8192 u16."""
    if task[3] & TASK_TRANSPOSED:
        c8 = r >> 3
        return stage[(c8 * 64 + (jl ^ (c8 & 7))) * 8 + (r & 7)]
    c8l = jl >> 3
    return stage[(((r >> 6) * 8 + c8l) * 64 + ((r & 63) ^ (c8l & 7))) * 8 + (jl & 7)]
