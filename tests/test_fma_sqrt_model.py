"""Host-side model of the MUFU-free square root of the 16-bit tile image (smh_common.cuh: sqrt2_fma_pipe).

The kernel form is integer arithmetic and fp32 FMAs only, so numpy can restate it exactly (an fp32 FMA = the fp64 product
of two fp32 values plus an fp32 value, rounded once more: double rounding can differ in the last bit, which is far below
the bound checked here).  The constants are read out of the header, so the test follows the kernel.  The exhaustive
on-device check is tests/test_gpu_parity.py::test_fma_pipe_sqrt_selftest.
"""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "simhand_b200", "csrc", "smh_common.cuh")
f32 = np.float32


def _constants():
    src = open(HDR).read()
    magic = int(re.search(r"kRsqMagic\s*=\s*(0x[0-9a-fA-F]+)u", src).group(1), 16)
    g1, g2 = re.search(r"kGold1\s*=\s*([0-9.eE+-]+)f\s*,\s*kGold2\s*=\s*([0-9.eE+-]+)f", src).groups()
    return magic, f32(g1), f32(g2)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def model_sqrt_fma_pipe(x):
    magic, c1, c2 = _constants()
    x = np.ascontiguousarray(x, dtype=f32)
    s = x.view(np.uint32) >> np.uint32(1)
    y = (np.uint32(magic) - s).view(f32)
    nh = (np.uint32((magic + 0x7F800000) & 0xFFFFFFFF) - s).view(f32)
    g = x * y
    r = _fma(g, nh, np.full_like(x, c1))
    g = _fma(g, r, g)
    nh = _fma(nh, r, nh)
    r = _fma(g, nh, np.full_like(x, c2))
    return _fma(g, r, g)


def model_sqrt_fma_pipe_acc(x, acc):
    """sqrt2_fma_pipe_acc: the last step takes the running sum as its addend (r' = 1 + r from the constant 1 + c2)."""
    magic, c1, c2 = _constants()
    x = np.ascontiguousarray(x, dtype=f32)
    s = x.view(np.uint32) >> np.uint32(1)
    y = (np.uint32(magic) - s).view(f32)
    nh = (np.uint32((magic + 0x7F800000) & 0xFFFFFFFF) - s).view(f32)
    g = x * y
    r = _fma(g, nh, np.full_like(x, c1))
    g = _fma(g, r, g)
    nh = _fma(nh, r, nh)
    r = _fma(g, nh, np.full_like(x, f32(1.0) + c2))
    return _fma(g, r, np.ascontiguousarray(acc, dtype=f32))


def test_accumulating_form_bound():
    bits = np.arange(0x3F800000, 0x40800000, 3, dtype=np.uint32)
    x = np.concatenate([bits.view(f32), np.exp2(np.random.default_rng(2).uniform(-100, 126, 1_000_000)).astype(f32)])
    with np.errstate(over="ignore"):
        g = model_sqrt_fma_pipe_acc(x, np.zeros_like(x))
    ref = np.sqrt(x.astype(np.float64))
    rel = (g.astype(np.float64) - ref) / ref
    assert np.abs(rel).max() < 8.0e-7, np.abs(rel).max()
    assert abs(rel.mean()) < 2e-7
    # with a running sum: acc + sqrt(x) to fp32 rounding of the sum
    acc = np.full_like(x[:1000], 123.456)
    out = model_sqrt_fma_pipe_acc(x[:1000], acc)
    want = acc.astype(np.float64) + np.sqrt(x[:1000].astype(np.float64))
    assert np.abs(out - want).max() <= 1.2e-7 * np.abs(want).max() + 8e-7 * np.sqrt(x[:1000]).max()


def test_negated_half_seed_is_exact():
    magic, _, _ = _constants()
    x = np.random.default_rng(0).uniform(1e-12, 1e12, 100000).astype(f32)
    s = x.view(np.uint32) >> np.uint32(1)
    y = (np.uint32(magic) - s).view(f32)
    nh = (np.uint32((magic + 0x7F800000) & 0xFFFFFFFF) - s).view(f32)
    assert np.array_equal(nh, -(y * f32(0.5)))


def test_fma_pipe_sqrt_bound():
    # the seed's error depends on the mantissa and the exponent's parity only: two adjacent octaves, every 3rd float,
    # plus a spread of exponents over the whole input domain of the kernel (coordinates up to 2^60)
    bits = np.arange(0x3F800000, 0x40800000, 3, dtype=np.uint32)
    x = np.concatenate([bits.view(f32), np.exp2(np.random.default_rng(1).uniform(-100, 126, 2_000_000)).astype(f32)])
    with np.errstate(over="ignore"):
        g = model_sqrt_fma_pipe(x)
    ref = np.sqrt(x.astype(np.float64))
    rel = (g.astype(np.float64) - ref) / ref
    assert np.abs(rel).max() < 7.5e-7, np.abs(rel).max()
    assert abs(rel.mean()) < 2e-7                                  # centred: no systematic shift of D


def test_fma_pipe_sqrt_special_values():
    with np.errstate(all="ignore"):
        out = model_sqrt_fma_pipe(np.array([0.0, np.inf, np.nan], dtype=f32))
    assert out[0] == 0.0 and not np.signbit(out[0])               # coincident joints: exactly +0, no NaN
    assert np.isnan(out[1]) and np.isnan(out[2])                   # non-finite inputs surface as NaN (the step is flagged)
