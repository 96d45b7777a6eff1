"""ctypes binding of libgode.so (the C ABI in include/gode.h).

The product path has NO fallback: if the shared library is missing or a kernel is not compiled for the
requested shape, callers get an exception, never a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GODE_LIB") or os.path.join(HERE, "csrc", "libgode.so")   # GODE_LIB: developer trace build only

METHODS = {"rk4": 0, "euler": 1, "midpoint": 2}
TABLEAUS = {"dopri5": 0, "bosh3": 1, "adaptive_heun": 2}
PREC = {"fp32": 0, "tf32": 1, "bf16": 2}
LAYOUT_TBD, LAYOUT_BTD = 0, 1
NORM_BATCH, NORM_TRAJ = 0, 1
MAX_HOST_STEPS = 255
ADAPTIVE_MAX_T, SDE_MAX_STEPS, SDE_MAX_FRAMES, SDE_MAX_CELLS, SDE_MAX_REV_STEPS = 256, 320, 64, 768, 384   # include/gode.h
SYNC_REGION_BYTES = 6400 * 1024   # include/gode.h GODE_SYNC_REGION_BYTES
LAUNCH_PDL_BWD = 1

ST_DT_UNDERFLOW, ST_NONFINITE, ST_MAX_STEPS, ST_CKPT_OVERFLOW, ST_PEER_TIMEOUT = 1, 2, 4, 8, 16


class GodeStepLog(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_attempts", C.c_int32), ("n_accepted", C.c_int32), ("nfe", C.c_int32),
                ("dt0", C.c_double), ("t_final", C.c_double)]


class GodeAdaptiveOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("first_step", C.c_double), ("safety", C.c_double),
                ("ifactor", C.c_double), ("dfactor", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double),
                ("max_num_steps", C.c_int32), ("norm_scope", C.c_int32), ("log_capacity", C.c_int32),
                ("ckpt_capacity", C.c_int32), ("fsign", C.c_float), ("tableau", C.c_int32)]


class GodeWorld(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("total_B", C.c_int64), ("slots_dev", C.c_void_p),
                ("launch_ctr", C.c_void_p)]


class GodeError(RuntimeError):
    pass


_P = C.c_void_p
_I = C.c_int
_SIGS = {
    "gode_strerror": (C.c_char_p, [_I]),
    "gode_version": (C.c_char_p, []),
    "gode_workspace_init": (_I, [_P, C.c_size_t, _P]),
    "gode_stream_capture_id": (_I, [_P, C.POINTER(C.c_ulonglong)]),
    "gode_set_thread_launch_flags": (_I, [_I]),
    "gode_set_status_mailbox": (_I, [_P]),
    "gode_supported": (_I, [_I, _I, _I]),
    "gode_param_count": (_I, [_I, _I]),
    "gode_rk4_fwd": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "gode_bwd_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "gode_fixed_adjoint_bwd_substep": (_I, [_I] + [_P] * 6 + [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "gode_rk4_sampler_fwd": (_I, [_P, _P, _P, _P, C.c_float, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_uint64, C.c_int64, _P, _I,
                                  _P, _I, _P, _P]),
    "gode_rk4_adjoint_bwd_strided": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "gode_rk4_bwd_workspace_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "gode_rk4_adjoint_bwd": (_I, [_P] * 6 + [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "gode_rk4_backprop_bwd": (_I, [_P] * 6 + [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "gode_dopri5_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "gode_dopri5_backprop_workspace_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "gode_dopri5_fwd": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, C.POINTER(GodeAdaptiveOpts), _I] + [_P] * 10 + [C.c_size_t, _P]),
    "gode_dopri5_fwd_world": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, C.POINTER(GodeAdaptiveOpts), _I] + [_P] * 10 +
                              [C.c_size_t, C.POINTER(GodeWorld), _P]),
    "gode_rk4_bwd_world": (_I, [_I] + [_P] * 6 + [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, C.POINTER(GodeWorld), _P]),
    "gode_dopri5_backprop_bwd_world": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, C.c_float, _P, _P, _P, C.c_size_t,
                                            C.POINTER(GodeWorld), _P]),
    "gode_dopri5_adjoint_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "gode_dopri5_adjoint_bwd": (_I, [_P] * 6 + [_P, _I, _I, _I, _I, _I, C.POINTER(GodeAdaptiveOpts), _I] + [_P] * 7 + [C.c_size_t, _P]),
    "gode_adaptive_backprop_bwd": (_I, [_I] + [_P] * 5 + [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, C.c_float, _P, _P, _P, C.c_size_t, _P]),
    "gode_dopri5_backprop_bwd": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, C.c_float, _P, _P, _P, C.c_size_t, _P]),
    "gode_dopri5_traj_fwd": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, C.POINTER(GodeAdaptiveOpts), _I] + [_P] * 11),
    "gode_dopri5_traj_backprop_bwd": (_I, [_P] * 5 + [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, C.c_float, _P, _P, _P,
                                           C.c_size_t, _P]),
    "gode_gru_param_count": (_I, [_I]),
    "gode_gru_jump_fwd": (_I, [_P] * 6 + [_I, _I, _P, _P]),
    "gode_gru_jump_bwd": (_I, [_P] * 7 + [_I, _I, _P, _P, _P, _P, C.c_size_t, _P]),
    "gode_odernn_log_stride": (C.c_size_t, [_I]),
    "gode_odernn_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "gode_odernn_fwd": (_I, [_P] * 10 + [_I, _I, _I, _I, C.POINTER(GodeAdaptiveOpts)] + [_P] * 7 + [C.c_size_t, _P]),
    "gode_odernn_bwd": (_I, [_P] * 10 + [_I, _I, _I, _I, _I, _I] + [_P] * 6 + [_I] + [_P] * 6 + [C.c_size_t, _P]),
    "gode_fixed_fwd": (_I, [_I] + [_P] * 5 + [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "gode_fixed_adjoint_bwd": (_I, [_I] + [_P] * 6 + [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "gode_fixed_backprop_bwd": (_I, [_I] + [_P] * 6 + [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "gode_allreduce_p2p": (_I, [_P, _I, _P, _P, _I, _I, _I, _P, _P]),
    "gode_sde_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "gode_sde_em_fwd": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P, C.c_uint64, C.c_int64, _I, _P, _P, _P]),
    "gode_sde_em_fwd_cells": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, C.c_uint64, C.c_int64, _I, _P, _P]),
    "gode_sde_adjoint_bwd": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, C.c_uint64, C.c_int64, _I,
                                  _P, _P, _P, C.c_size_t, _P]),
    "gode_sde_em_bwd": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P, C.c_uint64, C.c_int64, _I, _P, _P, _P,
                             C.c_size_t, _P]),
}

_lib = None


def lib():
    """Load libgode.so once.  Raises loudly if it has not been built (python -m gan_ode_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GodeError("libgode.so not found at {} — build it with `python -m gan_ode_b200.build`; "
                            "there is no CPU/PyTorch fallback for this path".format(LIB_PATH))
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


ERR_COOP = -5  # include/gode.h GODE_ERR_COOP


def check(rc: int, what: str):
    if rc != 0:
        raise GodeError("{} failed: {} (code {})".format(what, lib().gode_strerror(rc).decode(), rc))


def declared_symbols():
    """Every function name declared in include/gode.h (used by the CPU test that the library exports them all)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), "include", "gode.h")
    src = open(hdr).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gode_[a-z0-9_]+)\s*\(", src)))
