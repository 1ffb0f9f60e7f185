"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference hot path.

Three layers, each citing what it follows (paths relative to the reference checkout):

* `port_*`  -- the same torch ops, in the same order, as `src/models/utils.py:229-259`
  (weights) and `:407-426` (loss), evaluated in row chunks so the `[2N,2N,21,2]` temporary
  of `:252` never exceeds a few hundred MB.  Chunking over rows does not change any bit
  (SURVEY.md A.3).  This is the `"port"` CPU baseline timed by bench.py.
* `c_*`     -- ctypes bindings of `oracle/smh_oracle.c`, the plain-C restatement with the
  explicit fp32 operation order and a double-precision loss/gradient.
* `closed_form_fp64` -- torch fp64 closed form of SURVEY.md section 7.2 on dense fp32
  weights (small sizes).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsmh_oracle.so")


# --------------------------------------------------------------------------------------
# torch port (same ATen ops as the reference)
# --------------------------------------------------------------------------------------
def port_mpjpe_rows(batch_joints: torch.Tensor, r0: int, r1: int) -> torch.Tensor:
    """Rows [r0, r1) of `neg_dist_matrix` exactly as utils.py:252-253 builds it."""
    nd = torch.norm(batch_joints[r0:r1].unsqueeze(1) - batch_joints.unsqueeze(0), dim=-1)
    return nd.mean(dim=2)


def port_get_weights_linear(joints1, joints2, chunk: int = 256, dense: bool = True):
    """utils.py:218-261 with diff_type == 'mpjpe'.  Returns (pos_w, neg_w or None, stats)."""
    pos_distance = torch.norm(joints1 - joints2, dim=-1)                 # :230
    pos_dist = pos_distance.mean(dim=1)                                  # :231
    pmax, pmin = pos_dist.max(), pos_dist.min()                          # :233-234
    pos_w = (pmax - pos_dist) / (pmax - pmin)                            # :235
    bj = torch.cat((joints1, joints2), dim=0)                            # :237-239
    m = bj.shape[0]
    blocks = []
    dmax = torch.tensor(-float("inf"))
    dmin = torch.tensor(float("inf"))
    for r0 in range(0, m, chunk):
        blk = port_mpjpe_rows(bj, r0, min(m, r0 + chunk))
        dmax, dmin = torch.maximum(dmax, blk.max()), torch.minimum(dmin, blk.min())
        if dense:
            blocks.append(blk)
    stats = dict(dmax=dmax.float(), dmin=dmin.float(), pmax=pmax, pmin=pmin)
    if not dense:
        return pos_w, None, stats
    d = torch.cat(blocks, 0)
    neg_w = (dmax - d) / (dmax - dmin)                                   # :259
    return pos_w, neg_w, stats


def port_loss(z1, z2, pos_w, neg_w, temperature: float = 0.5):
    """utils.py:407-426, op for op (kept differentiable in z1, z2)."""
    z = torch.cat([z1, z2], dim=0)
    n = len(z)
    cov = torch.mm(z, z.t().contiguous())
    sim = torch.exp(cov * neg_w / temperature)
    mask = ~torch.eye(n, device=sim.device).bool()
    neg = sim.masked_select(mask).view(n, -1).sum(dim=-1)
    pos = torch.exp(torch.sum(z1 * z2, dim=-1) * pos_w / temperature)
    pos = torch.cat([pos, pos], dim=0)
    return -torch.log(pos / neg).mean()


def port_step(z1, z2, joints1, joints2, temperature: float = 0.5):
    """One fwd+bwd step of the reference path on the CPU (small sizes: dense weights)."""
    z1 = z1.detach().clone().requires_grad_(True)
    z2 = z2.detach().clone().requires_grad_(True)
    pos_w, neg_w, _ = port_get_weights_linear(joints1, joints2)
    loss = port_loss(z1, z2, pos_w, neg_w, temperature)
    loss.backward()
    return loss.detach(), z1.grad, z2.grad, pos_w, neg_w


def port_step_rows(z, bj, r0: int, r1: int, dmax: float, temperature: float = 0.5):
    """Bounded sample of one step at sizes whose dense form does not fit the host
    (2N = 16384 needs 63 GiB): the reference's ops restricted to the row
    block [r0, r1) -- weights, logits, exp, row sums and the autograd backward of that block.
    Used only for timing by bench.py's CPU legs."""
    zr = z[r0:r1].detach().clone().requires_grad_(True)
    zc = z.detach()
    d = port_mpjpe_rows(bj, r0, r1)
    w = (dmax - d) / dmax
    cov = torch.mm(zr, zc.t().contiguous())
    sim = torch.exp(cov * w / temperature)
    idx = torch.arange(r0, r1)
    mask = torch.ones_like(sim, dtype=torch.bool)
    mask[idx - r0, idx] = False
    neg = sim.masked_select(mask).view(r1 - r0, -1).sum(dim=-1)
    loss = torch.log(neg).sum()
    loss.backward()
    return float(loss.detach()), zr.grad


# --------------------------------------------------------------------------------------
# fp64 closed form on dense fp32 weights
# --------------------------------------------------------------------------------------
def closed_form_fp64(z1, z2, pos_w, neg_w, temperature: float = 0.5):
    z = torch.cat([z1, z2], 0).double()
    m, n = z.shape[0], z1.shape[0]
    wn, wp = neg_w.double(), pos_w.double()
    s = z @ z.t()
    e = torch.exp(s * wn / temperature)
    e.fill_diagonal_(0.0)
    neg = e.sum(1)
    part = torch.cat([z[n:], z[:n]], 0)
    wp2 = torch.cat([wp, wp], 0)
    pos = (z * part).sum(1) * wp2 / temperature
    loss = (torch.log(neg) - pos).mean()
    a = wn * e
    rn = 1.0 / neg
    g = (a * (rn[:, None] + rn[None, :])) @ z / (m * temperature)
    g = g - (2.0 / (m * temperature)) * wp2[:, None] * part
    return loss, g[:n], g[n:], neg


# --------------------------------------------------------------------------------------
# C oracle bindings
# --------------------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "smh_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def c_lib():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c_oracle())
        fp, dp, ci, cd = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double),
                          ctypes.c_int, ctypes.c_double)
        lib.smh_oracle_mpjpe_rows.argtypes = [fp, ci, ci, ci, fp]
        lib.smh_oracle_mpjpe_minmax.argtypes = [fp, ci, fp, fp]
        lib.smh_oracle_pos_weights.argtypes = [fp, ci, fp, fp, fp]
        lib.smh_oracle_neg_weights_rows.argtypes = [fp, ci, ci, ci, ctypes.c_float, ctypes.c_float, fp]
        lib.smh_oracle_step.argtypes = [fp, fp, ci, ci, cd, dp, dp, dp, fp, fp]
        lib.smh_oracle_step_rows.argtypes = [fp, fp, ci, ci, cd, ci, ci]
        lib.smh_oracle_step_rows.restype = cd
        _lib = lib
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def pack_joints(joints1, joints2) -> np.ndarray:
    """[2N, 42] contiguous fp32 from the two `[N,21,2]` (possibly strided) views (utils.py:237)."""
    bj = torch.cat((joints1, joints2), dim=0).contiguous().float()
    return bj.reshape(bj.shape[0], 42).numpy()


def c_mpjpe_rows(bj42: np.ndarray, r0: int, r1: int) -> np.ndarray:
    j, jp = _f32(bj42)
    out = np.empty((r1 - r0, j.shape[0]), np.float32)
    rc = c_lib().smh_oracle_mpjpe_rows(jp, j.shape[0], r0, r1,
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    assert rc == 0
    return out


def c_minmax(bj42: np.ndarray):
    j, jp = _f32(bj42)
    a, b = ctypes.c_float(), ctypes.c_float()
    assert c_lib().smh_oracle_mpjpe_minmax(jp, j.shape[0], ctypes.byref(a), ctypes.byref(b)) == 0
    return np.float32(a.value), np.float32(b.value)


def c_neg_weights_rows(bj42, r0, r1, dmax, dmin) -> np.ndarray:
    j, jp = _f32(bj42)
    out = np.empty((r1 - r0, j.shape[0]), np.float32)
    assert c_lib().smh_oracle_neg_weights_rows(
        jp, j.shape[0], r0, r1, float(dmax), float(dmin),
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
    return out


def c_step(z1, z2, joints1, joints2, temperature: float = 0.5, want_grad: bool = True):
    """Full step through the C oracle.  Returns dict(loss, dz1, dz2, neg, pos_w, stats)."""
    z = torch.cat([z1, z2], 0).contiguous().float().numpy()
    zz, zp = _f32(z)
    j, jp = _f32(pack_joints(joints1, joints2))
    m, d = zz.shape
    n = m // 2
    loss = ctypes.c_double()
    dz = np.empty((m, d), np.float64) if want_grad else None
    neg = np.empty(m, np.float64)
    pw = np.empty(n, np.float32)
    stats = np.empty(4, np.float32)
    dp = ctypes.POINTER(ctypes.c_double)
    fpp = ctypes.POINTER(ctypes.c_float)
    rc = c_lib().smh_oracle_step(zp, jp, n, d, float(temperature), ctypes.byref(loss),
                                 dz.ctypes.data_as(dp) if want_grad else None,
                                 neg.ctypes.data_as(dp), pw.ctypes.data_as(fpp),
                                 stats.ctypes.data_as(fpp))
    assert rc == 0
    return dict(loss=loss.value, dz1=None if dz is None else dz[:n], dz2=None if dz is None else dz[n:],
                neg=neg, pos_w=pw,
                stats=dict(dmax=stats[0], dmin=stats[1], pmax=stats[2], pmin=stats[3]))


def c_step_rows(z, bj42, n, temperature, r0, r1) -> float:
    zz, zp = _f32(z)
    j, jp = _f32(bj42)
    return c_lib().smh_oracle_step_rows(zp, jp, n, zz.shape[1], float(temperature), r0, r1)


# --------------------------------------------------------------------------------------
# comparison helpers
# --------------------------------------------------------------------------------------
def ulp_distance(a, b) -> np.ndarray:
    """Distance in units of fp32 representable values."""
    ai = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def grad_metrics(g, ref):
    """(cosine similarity, max|g - ref| / max|ref|) -- the two gradient figures north_star names."""
    g = np.asarray(g, np.float64).ravel()
    ref = np.asarray(ref, np.float64).ravel()
    cos = float(g @ ref / (np.linalg.norm(g) * np.linalg.norm(ref) + 1e-300))
    maxabs = float(np.abs(g - ref).max() / (np.abs(ref).max() + 1e-300))
    return cos, maxabs


# --------------------------------------------------------------------------------------
# projection-space transform: restatement of get_transformed_projections
# (src/models/unsupervised/simhand_w_model.py:55-94) with translate_encodings (utils.py:661-684)
# and rotate_encoding / get_rotation_2D_matrix (utils.py:606-658), out-of-place and differentiable
# --------------------------------------------------------------------------------------
def port_transform(projections, translate_x=None, translate_y=None, angle=None, eps: float = 1e-12):
    """projections [rows, d] -> normalize(rotate(translate(normalize(projections)))); d/2 2-D points per row.
    translate_* / angle are the values the reference passes to its helpers (the models pass -jitter, -angles)."""
    rows, d = projections.shape
    y = torch.nn.functional.normalize(projections, dim=1, eps=eps).view(rows, d // 2, 2)
    px, py = y[..., 0], y[..., 1]
    if translate_x is not None:
        yd = y.detach()
        ext = yd.max(dim=1).values - yd.min(dim=1).values                 # utils.py:673-674
        px = px + (translate_x * ext[:, 0]).view(-1, 1)                   # :676-678
        py = py + (translate_y * ext[:, 1]).view(-1, 1)                   # :679-681
    if angle is not None:
        cx, cy = px.detach().mean(dim=1), py.detach().mean(dim=1)         # :649
        rad = angle * np.pi / 180                                         # :622
        al, be = torch.cos(rad), torch.sin(rad)                           # :623-624 (scale = 1)
        ox = (1 - al) * cx - be * cy                                      # :627
        oy = (1 - al) * cy + be * cx                                      # :630
        qx = px * al.view(-1, 1) + py * be.view(-1, 1) + ox.view(-1, 1)   # [x, y, 1] @ rot_mat[:, :, 0]
        qy = -px * be.view(-1, 1) + py * al.view(-1, 1) + oy.view(-1, 1)  # [x, y, 1] @ rot_mat[:, :, 1]
        px, py = qx, qy
    out = torch.stack([px, py], dim=-1).reshape(rows, d)
    return torch.nn.functional.normalize(out, dim=1, eps=eps)


# --------------------------------------------------------------------------------------
# the other weightings of the reference: diff_type w_abs / w_o_abs and weight_type non_linear,
# restated from get_weights_linear (utils.py:218-261) and get_weights_nonlinear (utils.py:304-346)
# --------------------------------------------------------------------------------------
def port_distances(joints1, joints2, diff_type: str):
    """(positive distances [N], all-pairs distances [2N, 2N]) for one diff_type; dense, small sizes only."""
    bj = torch.cat((joints1, joints2), dim=0)
    if diff_type == "mpjpe":
        pos = torch.norm(joints1 - joints2, dim=-1).mean(dim=1)                               # :229-231
        neg = torch.norm(bj.unsqueeze(1) - bj.unsqueeze(0), dim=-1).mean(dim=2)               # :251-253
    elif diff_type in ("w_abs", "w_o_abs"):
        fn = torch.abs if diff_type == "w_abs" else (lambda t: t)
        pos = torch.norm(fn(joints1 - joints2).mean(dim=1), dim=1)                            # :219-227: mean over joints
        neg = torch.norm(torch.mean(fn(bj.unsqueeze(1) - bj.unsqueeze(0)), dim=-1), dim=2)    # :241-249: mean over x, y
    elif diff_type == "pca":
        # *_with_pca (utils.py:264-301, :349-388): joints are [N, K] coordinate vectors; every reference diff_type is
        # the Euclidean distance between them
        j1, j2 = joints1.reshape(joints1.shape[0], -1), joints2.reshape(joints2.shape[0], -1)
        bj = torch.cat((j1, j2), dim=0)
        pos = torch.norm(j1 - j2, dim=-1)                                                     # :265-274
        neg = torch.norm(bj.unsqueeze(1) - bj.unsqueeze(0), dim=-1)                            # :282-293
    else:
        raise ValueError(diff_type)
    return pos, neg


def port_get_weights(joints1, joints2, weight_type: str = "linear", diff_type: str = "mpjpe",
                     lambda_pos: float = 0.0, lambda_neg: float = 0.0):
    pos, neg = port_distances(joints1, joints2, diff_type)
    if weight_type == "linear":
        pos_w = (pos.max() - pos) / (pos.max() - pos.min())                                   # :233-235
        neg_w = (neg.max() - neg) / (neg.max() - neg.min())                                   # :255-259
    else:
        pos_w = 1 / (1 + torch.exp(lambda_pos * (pos - pos.mean())))                          # :323-325
        neg_w = 1 / (1 + torch.exp(lambda_neg * (neg - neg.mean())))                          # :343-346
    return pos_w, neg_w
