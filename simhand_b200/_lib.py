"""ctypes binding of libsimhand_b200.so (the C ABI declared in include/simhand_b200.h).

The library is the product: if it is missing the import fails loudly -- there is no fallback path.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMH_LIB") or os.path.join(_HERE, "lib", "libsimhand_b200.so")      # SMH_LIB: development builds

ENGINE_TC_TF32 = 0
ENGINE_FP32 = 1
ENGINE_TC_BF16 = 2
ENGINE_TC_FP16 = 3
ENGINES = {"tf32": ENGINE_TC_TF32, "tc": ENGINE_TC_TF32, "fp32": ENGINE_FP32, "bf16": ENGINE_TC_BF16,
           "fp16": ENGINE_TC_FP16}

PREP_NO_ZERO = 0x100
BACKWARD_RN_ONLY = 0x200
UNIT_NEG_WEIGHTS = 0x400
UNIT_POS_WEIGHTS = 0x800
DENSE_WEIGHTS = 0x2000
FINALIZE_LOSS_PART = 0x4000
FINALIZE_GRAD = 0x8000
SHARD_PREP_NO_IMAGES = 0x10000
DIMS_DENSE_WEIGHTS = 1
DIMS_DENSE_BACKWARD = 2
DIMS_Q16_TILES = 4
DIFF_TYPES = {"mpjpe": 0, "w_abs": 1, "w_o_abs": 2, "pca": 3}
WEIGHT_TYPES = {"linear": 0, "non_linear": 1}
FLAG_SLOW_DOMAIN = 1
FLAG_NONFINITE = 2


class Dims(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("d", ctypes.c_int32), ("world", ctypes.c_int32),
                ("rank", ctypes.c_int32), ("strip_len", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("diff_type", ctypes.c_int32), ("weight_type", ctypes.c_int32),
                ("lambda_pos", ctypes.c_float), ("lambda_neg", ctypes.c_float)]


class Layout(ctypes.Structure):
    _fields_ = [("ws_bytes", ctypes.c_int64), ("plan_bytes", ctypes.c_int64), ("off_stats", ctypes.c_int64),
                ("off_zt", ctypes.c_int64), ("off_zb", ctypes.c_int64), ("off_zh", ctypes.c_int64),
                ("off_jp", ctypes.c_int64), ("off_posd", ctypes.c_int64),
                ("off_neg", ctypes.c_int64), ("off_rn", ctypes.c_int64), ("off_rowloss", ctypes.c_int64),
                ("off_dzacc", ctypes.c_int64), ("off_negparts", ctypes.c_int64), ("off_dzparts", ctypes.c_int64),
                ("off_dist", ctypes.c_int64), ("off_posinfo", ctypes.c_int64),
                ("m", ctypes.c_int32), ("tiles_per_side", ctypes.c_int32), ("n_stored_tiles", ctypes.c_int32),
                ("n_tasks", ctypes.c_int32), ("n_strips", ctypes.c_int32), ("strip_len", ctypes.c_int32),
                ("n_strips_fwd", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class Inputs(ctypes.Structure):
    _fields_ = [("z1", ctypes.c_void_p), ("z2", ctypes.c_void_p), ("z_row_stride", ctypes.c_int64),
                ("j1", ctypes.c_void_p), ("j2", ctypes.c_void_p),
                ("j_sample_stride", ctypes.c_int64), ("j_joint_stride", ctypes.c_int64),
                ("j_coord_stride", ctypes.c_int64), ("n_local", ctypes.c_int32),
                ("z_rank_stride", ctypes.c_int64), ("j_rank_stride", ctypes.c_int64)]


MAX_PEERS = 8


class Exchange(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("ws_peer", ctypes.c_void_p * MAX_PEERS), ("xin_peer", ctypes.c_void_p * MAX_PEERS),
                ("signal_peer", ctypes.c_void_p * MAX_PEERS), ("fused", ctypes.c_int32),
                ("timeout_ms", ctypes.c_uint32)]


class Head(ctypes.Structure):
    _fields_ = [("rows", ctypes.c_int64), ("in_dim", ctypes.c_int32), ("hidden", ctypes.c_int32), ("out_dim", ctypes.c_int32),
                ("fp16", ctypes.c_int32), ("training", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("x", ctypes.c_void_p), ("x_row_stride", ctypes.c_int64), ("w1", ctypes.c_void_p), ("b1", ctypes.c_void_p),
                ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("running_mean", ctypes.c_void_p),
                ("running_var", ctypes.c_void_p), ("bn_eps", ctypes.c_float), ("bn_momentum", ctypes.c_float),
                ("w2", ctypes.c_void_p), ("h", ctypes.c_void_p), ("colsum", ctypes.c_void_p), ("save_mean", ctypes.c_void_p),
                ("save_rstd", ctypes.c_void_p), ("y", ctypes.c_void_p), ("norm", ctypes.c_void_p), ("norm_eps", ctypes.c_float)]


class HeadBwd(ctypes.Structure):
    _fields_ = [("dy", ctypes.c_void_p), ("w2t", ctypes.c_void_p), ("dp", ctypes.c_void_p), ("a", ctypes.c_void_p),
                ("dhn", ctypes.c_void_p), ("dh", ctypes.c_void_p), ("colsum", ctypes.c_void_p), ("dgamma", ctypes.c_void_p),
                ("dbeta", ctypes.c_void_p)]


SIGNAL_WORDS = 256
SIG_EPOCH = 16
SIG_POISON = 17


class Stats(ctypes.Structure):
    _fields_ = [("dmax_bits", ctypes.c_uint32), ("pmax_bits", ctypes.c_uint32), ("pmin_inv", ctypes.c_uint32),
                ("flags", ctypes.c_uint32), ("loss", ctypes.c_float), ("counter", ctypes.c_uint32),
                ("fail_site", ctypes.c_uint32), ("ticket2", ctypes.c_uint32), ("dsum", ctypes.c_double),
                ("dbound_bits", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


# every symbol include/simhand_b200.h declares (tests check that the library exports all of them)
EXPORTS = ("smh_version", "smh_last_error", "smh_layout", "smh_plan_build", "smh_prep", "smh_mpjpe",
           "smh_forward", "smh_backward", "smh_finalize", "smh_weights_dense", "smh_l2norm_fwd",
           "smh_l2norm_bwd", "smh_selftest", "smh_tc_probe", "smh_tc_default_params", "smh_push_inputs",
           "smh_barrier", "smh_prep_zero", "smh_exchange_neg", "smh_exchange_dz", "smh_import_weights", "smh_transform_fwd", "smh_transform_bwd",
           "smh_scale_grads", "smh_shard_prep", "smh_shard_push_z", "smh_head_forward", "smh_head_backward")

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m simhand_b200.build` "
            "(nvcc, sm_100a).  simhand_b200 has no fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    pd, pl, pi = ctypes.POINTER(Dims), ctypes.POINTER(Layout), ctypes.POINTER(Inputs)
    lib.smh_version.restype = ctypes.c_int
    lib.smh_last_error.restype = ctypes.c_char_p
    lib.smh_layout.argtypes = [pd, pl]
    lib.smh_plan_build.argtypes = [pd, vp, i64]
    lib.smh_prep.argtypes = [pd, pi, vp, ctypes.c_int, vp]
    px = ctypes.POINTER(Exchange)
    lib.smh_mpjpe.argtypes = [pd, vp, vp, px, vp]
    lib.smh_forward.argtypes = [pd, vp, vp, f32, ctypes.c_int, px, vp]
    lib.smh_backward.argtypes = [pd, vp, vp, f32, ctypes.c_int, px, vp]
    lib.smh_push_inputs.argtypes = [px, pi, i32, i32, vp]
    lib.smh_barrier.argtypes = [px, vp]
    lib.smh_shard_prep.argtypes = [pd, pi, vp, px, ctypes.c_int, vp]
    lib.smh_shard_push_z.argtypes = [pd, pi, vp, px, ctypes.c_int, vp]
    lib.smh_exchange_neg.argtypes = [pd, vp, px, vp]
    lib.smh_exchange_dz.argtypes = [pd, vp, px, vp]
    lib.smh_prep_zero.argtypes = [pd, vp, vp]
    lib.smh_finalize.argtypes = [pd, pi, vp, vp, f32, f32, vp, vp, vp, i64, ctypes.c_int, ctypes.POINTER(Exchange), vp]
    lib.smh_weights_dense.argtypes = [pd, vp, vp, vp, vp, vp]
    lib.smh_import_weights.argtypes = [pd, vp, vp, vp, i64, vp, vp]
    lib.smh_head_forward.argtypes = [ctypes.POINTER(Head), vp]
    lib.smh_head_backward.argtypes = [ctypes.POINTER(Head), ctypes.POINTER(HeadBwd), vp]
    lib.smh_scale_grads.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.smh_l2norm_fwd.argtypes = [vp, vp, vp, i64, i32, f32, vp]
    lib.smh_l2norm_bwd.argtypes = [vp, vp, vp, vp, i64, i32, f32, vp]
    lib.smh_transform_fwd.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp, i64, i32, f32, vp]
    lib.smh_transform_bwd.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, i64, i64, i32, f32, vp]
    lib.smh_selftest.argtypes = [ctypes.c_int, vp, i64, vp]
    lib.smh_tc_probe.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), vp, vp, vp, vp, vp]
    lib.smh_tc_default_params.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
    lib.smh_tc_default_params.restype = None
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("smh_version", "smh_last_error", "smh_tc_default_params"):
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


class SmhError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().smh_last_error().decode(errors="replace")
        raise SmhError(f"{what or 'simhand_b200'} failed with code {rc}: {msg}")
