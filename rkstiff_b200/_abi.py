"""ctypes binding of include/rkstiff_b200.h.

The engine has no CPU fallback: if the shared library is missing or does not load, importing
this module raises, and every solver constructor raises with it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p)

from . import _build

RKS_ABI_VERSION = 1

METHOD_IDS = {"IF4": 0, "ETD4": 1, "ETD5": 2, "IF34": 3, "ETD34": 4, "ETD35": 5, "IF45DP": 6}
MODEL_NONE, MODEL_UUX_RFFT, MODEL_NLS_FFT, MODEL_CUBIC_RFFT, MODEL_SINE_GORDON = 0, 1, 2, 3, 4
MODEL_DERIV_FFT, MODEL_DERIV_RFFT_PAIR = 5, 6      # row transforms only (derivatives.py)
CTRL_RUNNING, CTRL_DONE, CTRL_MAX_LOOPS, CTRL_MIN_STEP = 0, 1, 2, 3
LOG_CAP = 4096
ROW_LOG_CAP = 64


class RksConfig(Structure):
    _fields_ = [("epsilon", c_double), ("incr_f", c_double), ("decr_f", c_double), ("safety_f", c_double),
                ("adapt_cutoff", c_double), ("minh", c_double), ("modecutoff", c_double),
                ("contour_radius", c_double), ("contour_points", c_int32), ("if45dp_r4_fix", c_int32)]


class RksCtrl(Structure):
    _fields_ = [("h", c_double), ("h_last", c_double), ("h_coeff", c_double), ("t", c_double), ("tf", c_double),
                ("s_last", c_double), ("step_count", c_int64), ("trial_count", c_int64), ("nl_evals", c_int64),
                ("coeff_updates", c_int64), ("status", c_int32), ("accept", c_int32), ("numloops", c_int32),
                ("u_sel", c_int32), ("n_sel", c_int32), ("need_n1", c_int32), ("log_count", c_int32),
                ("snap_count", c_int32)]


class RksTrialRec(Structure):
    _fields_ = [("h", c_double), ("s", c_double), ("t_after", c_double), ("accepted", c_int32), ("pad", c_int32)]


class EngineError(RuntimeError):
    """A call into librkstiff_b200.so failed."""


def _load():
    path = _build.LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA engine has not been built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). rkstiff_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    P = c_void_p
    sig = {
        "rks_abi_version": (c_int, []),
        "rks_last_error": (c_char_p, []),
        "rks_num_stages": (c_int, [c_int]),
        "rks_num_nl_buffers": (c_int, [c_int]),
        "rks_is_adaptive": (c_int, [c_int]),
        "rks_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int64, c_int]),
        "rks_plan_create": (c_int, [POINTER(P), c_int, c_int64, c_int64, P, c_int, c_int64, POINTER(RksConfig), P,
                                    c_size_t, P]),
        "rks_plan_destroy": (None, [P]),
        "rks_workspace_bytes_indexed": (c_size_t, [c_int, c_int64, c_int64, c_int64, c_int]),
        "rks_plan_create_indexed": (c_int, [POINTER(P), c_int, c_int64, c_int64, P, c_int, c_int64, P,
                                            POINTER(RksConfig), P, c_size_t, P]),
        "rks_workspace_bytes_separable": (c_size_t, [c_int, c_int64, c_int, POINTER(c_int64), c_int]),
        "rks_plan_create_separable": (c_int, [POINTER(P), c_int, c_int64, c_int, POINTER(c_int64), P, c_int,
                                              POINTER(RksConfig), P, c_size_t, P]),
        "rks_workspace_bytes_independent": (c_size_t, [c_int, c_int64, c_int64, c_int]),
        "rks_plan_create_independent": (c_int, [POINTER(P), c_int, c_int64, c_int64, P, c_int, POINTER(RksConfig), P,
                                                c_size_t, P]),
        "rks_set_config": (c_int, [P, POINTER(RksConfig), P]),
        "rks_set_model": (c_int, [P, c_int, c_int64, P, POINTER(c_double), c_int, P]),
        "rks_set_model_nd": (c_int, [P, c_int, c_int, POINTER(c_int64), c_double, P]),
        "rks_begin": (c_int, [P, c_double, c_double, c_double, c_int64, c_int, c_int, P]),
        "rks_set_h": (c_int, [P, c_double, P]),
        "rks_set_u": (c_int, [P, P, P]),
        "rks_get_u": (c_int, [P, P, P]),
        "rks_update_coeffs": (c_int, [P, P]),
        "rks_stage": (c_int, [P, c_int, P]),
        "rks_nl": (c_int, [P, c_int, P]),
        "rks_stage_nl": (c_int, [P, c_int, P]),
        "rks_stage_nl_part": (c_int, [P, c_int, c_int, P]),
        "rks_nl_input": (P, [P, c_int]),
        "rks_nl_output": (P, [P, c_int]),
        "rks_error_control": (c_int, [P, P]),
        "rks_norm_override": (c_int, [P, P, P]),
        "rks_error_sums": (c_int, [P, P]),
        "rks_controller": (c_int, [P, P]),
        "rks_reduction_scalars": (P, [P]),
        "rks_run_trials": (c_int, [P, c_int, P, P, c_int, P]),
        "rks_run_fixed": (c_int, [P, c_int, P]),
        "rks_snapshot": (c_int, [P, P, P, c_int, P]),
        "rks_pointwise": (c_int, [c_int, P, P, c_int64, c_double, P]),
        "rks_gemv": (c_int, [P, P, P, c_int64, c_int64, P]),
        "rks_rows_create": (c_int, [POINTER(P), c_int, c_int64, P, c_double, P]),
        "rks_rows_apply": (c_int, [P, P, P, c_int64, P]),
        "rks_rows_destroy": (None, [P]),
        "rks_axis_create": (c_int, [POINTER(P), c_int64, P]),
        "rks_axis_apply": (c_int, [P, P, P, c_int64, c_int64, c_int, P]),
        "rks_axis_apply_chunked": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int, P]),
        "rks_axis_apply_scatter": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int, P]),
        "rks_axis_destroy": (None, [P]),
        "rks_peer_barrier": (c_int, [P, c_int, c_int, c_uint64, P]),
        "rks_read_ctrl": (c_int, [P, POINTER(RksCtrl), P]),
        "rks_read_log": (c_int, [P, POINTER(RksTrialRec), c_int, c_int, P]),
        "rks_read_rows": (c_int, [P, POINTER(RksCtrl), c_int64, P]),
        "rks_read_row_log": (c_int, [P, c_int64, POINTER(RksTrialRec), c_int, c_int, P]),
        "rks_array": (P, [P, c_char_p]),
        "rks_kernel_launches": (c_int64, [P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.rks_abi_version() != RKS_ABI_VERSION:
        raise ImportError("librkstiff_b200.so ABI version mismatch; rebuild it")
    return lib, sorted(sig)


lib, EXPORTS = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(f"rkstiff_b200 engine error {rc}: {lib.rks_last_error().decode()}")
