"""ctypes binding of libfos_b200.so (C ABI: include/fos.h).

Loading fails loudly: there is no CPU fallback and no pure-Python path behind any
solver entry point.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FOS_LIB_PATH: load another build of the library (A/B builds: `python -m fastoptsolver_b200.build --lean`)
LIB_PATH = os.environ.get("FOS_LIB_PATH") or os.path.join(HERE, "libfos_b200.so")

FOS_OK = 0
FOS_ERR_INVALID = -1
FOS_ERR_CUDA = -2
FOS_ERR_NOMEM = -3
FOS_ERR_UNSUPPORTED = -4
FOS_ERR_COMM = -5

FOS_F64, FOS_F32 = 0, 1
SCHEME_NESTEROV, SCHEME_DELTA, SCHEME_ISTA = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class PGParams(C.Structure):
    _fields_ = [
        ("scheme", C.c_int),
        ("alpha1", C.c_double), ("alpha2", C.c_double),
        ("obj_terms", C.c_int),
        ("delta", C.c_double),
        ("backtracking", C.c_int),
        ("eta", C.c_double),
        ("armijo_c", C.c_double),
        ("step0", C.c_double),
        ("max_iter", C.c_int),
        ("tol", C.c_double), ("tol_ratio", C.c_double),
        ("adaptive_restart", C.c_int),
        ("restart_threshold", C.c_double),
        ("want_history", C.c_int),
        ("x0", c_double_p),
    ]


class PGResult(C.Structure):
    _fields_ = [
        ("x", c_double_p), ("x_hist", c_double_p), ("obj_hist", c_double_p),
        ("t_hist", c_double_p), ("step_hist", c_double_p),
        ("ls_iters", c_int_p), ("grad_ms", c_float_p), ("ls_ms", c_float_p),
        ("n_iters", C.c_int), ("n_grad_calls", C.c_int), ("n_passes", C.c_int),
        ("stop_reason", C.c_int), ("loop_ms", C.c_float), ("kernel_launches", C.c_int64),
        ("grad_kernel_ms", C.c_float), ("grad_kernel_launches", C.c_int),
        ("epilogue_ms", C.c_float), ("exchange_ms", C.c_float),
        ("host_setup_ms", C.c_float), ("host_loop_ms", C.c_float), ("host_finish_ms", C.c_float),
    ]


class PathParams(C.Structure):
    _fields_ = [("alphas1", c_double_p), ("n_lambda", C.c_int), ("alpha2", C.c_double), ("step", C.c_double),
                ("max_iter", C.c_int), ("tol", C.c_double), ("check_every", C.c_int), ("X0", c_double_p)]


class PathResult(C.Structure):
    _fields_ = [("X", c_double_p), ("obj", c_double_p), ("n_iters", C.c_int), ("last_max_step", C.c_double),
                ("tile_rows", C.c_int), ("loop_ms", C.c_float), ("kernel_launches", C.c_int64)]


class LbfgsParams(C.Structure):
    _fields_ = [("m", C.c_int), ("max_iter", C.c_int), ("maxfun", C.c_int), ("maxls", C.c_int), ("obj_terms", C.c_int),
                ("alpha1", C.c_double), ("alpha2", C.c_double), ("pgtol", C.c_double), ("factr", C.c_double),
                ("x0", c_double_p)]


class LbfgsResult(C.Structure):
    _fields_ = [("x", c_double_p), ("obj_hist", c_double_p), ("f_final", C.c_double),
                ("n_iters", C.c_int), ("n_fg", C.c_int), ("n_skipped", C.c_int), ("stop_reason", C.c_int),
                ("loop_ms", C.c_float), ("kernel_launches", C.c_int64)]


# name -> (restype, argtypes); the CPU test-suite checks every one is exported
SIGNATURES = {
    "fos_abi_version": (C.c_int, []),
    "fos_last_error": (C.c_char_p, []),
    "fos_device_count": (C.c_int, []),
    "fos_trim": (C.c_int, []),
    "fos_device_info": (C.c_int, [C.c_int, c_int_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "fos_design_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64,
                                    C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "fos_design_create_begin": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fos_design_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "fos_design_create_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64,
                                           C.c_int, C.POINTER(C.c_void_p)]),
    "fos_design_create_synthetic": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_uint64, C.c_double, C.c_double,
                                              C.c_double, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "fos_design_destroy": (C.c_int, [C.c_void_p]),
    "fos_design_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), c_int_p,
                                   C.POINTER(C.c_int64)]),
    "fos_design_download": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "fos_design_pointers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "fos_design_upload_gram": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), c_int_p, c_float_p, c_float_p]),
    "fos_design_upload_gram_set": (C.c_int, [C.c_void_p, C.c_int]),
    "fos_design_column_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, c_double_p]),
    "fos_design_affine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "fos_design_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "fos_time_grad_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "fos_debug_cta_times": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, c_int_p]),
    "fos_debug_solve_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "fos_debug_gram_staging": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fos_debug_path_staging": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fos_debug_path_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fos_design_solve_kernel_ok": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "fos_design_solve_kernel_disable": (C.c_int, [C.c_void_p]),
    "fos_design_lambda_max": (C.c_int, [C.c_void_p, c_double_p]),
    "fos_comm_window_alloc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fos_comm_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "fos_comm_window_alloc_fd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p]),
    "fos_comm_attach_fd": (C.c_int, [C.c_void_p, c_int_p, C.c_int]),
    "fos_comm_window_free": (C.c_int, [C.c_void_p]),
    "fos_comm_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "fos_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, c_double_p]),
    "fos_objective": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, c_double_p]),
    "fos_power_iter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, c_double_p, c_int_p, c_float_p]),
    "fos_prox_l1": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int]),
    "fos_prox_elastic_net": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                       C.c_int]),
    "fos_prox_grad": (C.c_int, [C.c_void_p, C.POINTER(PGParams), C.POINTER(PGResult)]),
    "fos_lbfgs": (C.c_int, [C.c_void_p, C.POINTER(LbfgsParams), C.POINTER(LbfgsResult)]),
    "fos_gram_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "fos_gram_destroy": (C.c_int, [C.c_void_p]),
    "fos_gram_info": (C.c_int, [C.c_void_p, c_int_p, c_double_p, c_float_p, c_int_p]),
    "fos_gram_pointers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "fos_gram_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fos_gram_set_btb": (C.c_int, [C.c_void_p, C.c_double]),
    "fos_gram_subset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "fos_gram_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fos_gram_path_fista": (C.c_int, [C.c_void_p, C.POINTER(PathParams), C.POINTER(PathResult)]),
    "fos_mrhs_fista": (C.c_int, [C.c_void_p, C.POINTER(PathParams), C.POINTER(PathResult)]),
}

_lib = None


def load():
    """Return the loaded library; raise ImportError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fastoptsolver_b200.build` "
            "(nvcc, sm_100a). fastoptsolver_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.fos_abi_version() != 1:
        raise ImportError("libfos_b200.so ABI version mismatch")
    _lib = lib
    return lib


class FosError(RuntimeError):
    pass


def check(status):
    """Map a fos_status to the exception types of the reference's boundary: bad arguments
    -> ValueError (objective_functions.py:28, lbfgs.py:35), everything else RuntimeError."""
    if status == FOS_OK:
        return
    msg = load().fos_last_error().decode("utf-8", "replace")
    if status == FOS_ERR_INVALID:
        raise ValueError(msg)
    if status == FOS_ERR_NOMEM:
        raise MemoryError(msg)
    if status == FOS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise FosError(f"libfos_b200 error {status}: {msg}")
