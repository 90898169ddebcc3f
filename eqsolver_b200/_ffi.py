"""ctypes binding of include/eqsolver_b200.h -- one Python function per C symbol, nothing else.

The product path is the CUDA library; if it is missing or no GPU is present, calls fail loudly
(`KernelLibraryMissing` / `SolverError`) -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EQS_B200_LIBRARY: load a differently built variant of the same library (tuning sweeps, user-objective builds)
LIB_PATH = os.environ.get("EQS_B200_LIBRARY") or os.path.join(_HERE, "lib", "libeqsolver_b200.so")

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)

(OK, ERR_MAX_ITER, ERR_NAN, ERR_INCORRECT_INPUT, ERR_BAD_JACOBIAN, ERR_TYPE_CONVERSION, ERR_EXTERNAL) = range(7)
STALL_WINDOW = 50


class KernelLibraryMissing(RuntimeError):
    pass


class PsoParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("objective_id", C.c_int32), ("mode", C.c_int32), ("exchange", C.c_int32),
                ("particle_count", C.c_int64), ("iter_max", C.c_int64),
                ("tol", C.c_double), ("w", C.c_double), ("c1", C.c_double), ("c2", C.c_double),
                ("lower", dp), ("upper", dp), ("seed", C.c_uint64), ("replay_init", dp),
                ("virtual_rank", C.c_int32), ("virtual_world", C.c_int32)]


class CeParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("objective_id", C.c_int32), ("mean_divisor", C.c_int32), ("exchange", C.c_int32),
                ("sample_size", C.c_int64), ("elite_count", C.c_int64), ("iter_max", C.c_int64),
                ("tol", C.c_double), ("std_dev", dp), ("seed", C.c_uint64),
                ("virtual_rank", C.c_int32), ("virtual_world", C.c_int32)]


class McParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("integrand_id", C.c_int32), ("sample_count", C.c_int64),
                ("from_", dp), ("to", dp), ("seed", C.c_uint64), ("replay_u", dp)]


class MiserParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("integrand_id", C.c_int32), ("sample_count", C.c_int64),
                ("minimum_dimensional_sample_count", C.c_int64), ("maximum_depth", C.c_int32), ("reserved", C.c_int32),
                ("alpha", C.c_double), ("dither", C.c_double), ("from_", dp), ("to", dp), ("seed", C.c_uint64)]


class Report(C.Structure):
    _fields_ = [("iters_run", C.c_int64), ("evals", C.c_int64), ("best_cost", C.c_double),
                ("stalled", C.c_int32), ("world", C.c_int32), ("kernel_launches", C.c_int64),
                ("init_ms", C.c_double), ("loop_ms", C.c_double)]


# every symbol include/eqsolver_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
SYMBOLS = {
    "eqs_abi_version": (C.c_int, []),
    "eqs_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "eqs_ctx_destroy": (C.c_int, [_vp]),
    "eqs_last_error": (C.c_char_p, [_vp]),
    "eqs_ctx_sm_count": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "eqs_comm_unique_id": (C.c_int, [_vp]),
    "eqs_ctx_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "eqs_ctx_rank": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "eqs_shard_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, i64p, i64p]),
    "eqs_record_doubles": (C.c_int64, [C.c_int32]),
    "eqs_objective_eval": (C.c_int, [_vp, C.c_int32, dp, C.c_int32, C.c_int64, dp]),
    "eqs_philox_blocks": (C.c_int, [_vp, _u32p, _u32p, C.c_int64, _u32p]),
    "eqs_draws": (C.c_int, [_vp, C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, dp, dp]),
    "eqs_device_cos": (C.c_int, [_vp, dp, C.c_int64, dp]),
    "eqs_measure_fp64": (C.c_int, [_vp, dp]),
    "eqs_measure_copy": (C.c_int, [_vp, C.c_int64, dp]),
    "eqs_measure_pattern": (C.c_int, [_vp, C.c_int32, C.c_int64, dp]),
    "eqs_pso_validate": (C.c_int, [C.POINTER(PsoParams)]),
    "eqs_pso_solve": (C.c_int, [_vp, C.POINTER(PsoParams), dp, dp, C.POINTER(Report)]),
    "eqs_pso_create": (C.c_int, [_vp, C.POINTER(PsoParams), C.POINTER(_vp)]),
    "eqs_pso_destroy": (C.c_int, [_vp]),
    "eqs_pso_init": (C.c_int, [_vp, dp]),
    "eqs_pso_step": (C.c_int, [_vp, C.c_int64, dp]),
    "eqs_pso_sync": (C.c_int, [_vp]),
    "eqs_pso_last_step_ms": (C.c_int, [_vp, dp, i64p]),
    "eqs_pso_read_gbest": (C.c_int, [_vp, dp, dp, i64p, C.POINTER(C.c_int32), i64p]),
    "eqs_pso_read_state": (C.c_int, [_vp, dp, dp, dp, dp]),
    "eqs_pso_read_ring": (C.c_int, [_vp, dp, i64p]),
    "eqs_pso_local_range": (C.c_int, [_vp, i64p, i64p]),
    "eqs_pso_local_indices": (C.c_int, [_vp, i64p]),
    "eqs_pso_init_local": (C.c_int, [_vp, dp, C.c_int32]),
    "eqs_pso_init_apply": (C.c_int, [_vp]),
    "eqs_pso_step_local": (C.c_int, [_vp, dp]),
    "eqs_pso_step_apply": (C.c_int, [_vp]),
    "eqs_pso_get_record": (C.c_int, [_vp, dp]),
    "eqs_pso_set_records": (C.c_int, [_vp, dp]),
    "eqs_ce_validate": (C.c_int, [C.POINTER(CeParams)]),
    "eqs_ce_solve": (C.c_int, [_vp, C.POINTER(CeParams), dp, dp, C.POINTER(Report)]),
    "eqs_ce_create": (C.c_int, [_vp, C.POINTER(CeParams), dp, C.POINTER(_vp)]),
    "eqs_ce_destroy": (C.c_int, [_vp]),
    "eqs_ce_step": (C.c_int, [_vp, C.c_int64, dp]),
    "eqs_ce_sync": (C.c_int, [_vp]),
    "eqs_ce_last_step_ms": (C.c_int, [_vp, dp, i64p]),
    "eqs_ce_read_state": (C.c_int, [_vp, dp, dp, i64p, C.POINTER(C.c_int32)]),
    "eqs_ce_read_elites": (C.c_int, [_vp, i64p, i64p]),
    "eqs_ce_read_costs": (C.c_int, [_vp, dp]),
    "eqs_ce_local_range": (C.c_int, [_vp, i64p, i64p]),
    "eqs_ce_phase": (C.c_int, [_vp, C.c_int32, dp]),
    "eqs_ce_exchange_words": (C.c_int64, [_vp, C.c_int32]),
    "eqs_ce_get_exchange": (C.c_int, [_vp, C.c_int32, _vp]),
    "eqs_ce_set_exchange": (C.c_int, [_vp, C.c_int32, _vp]),
    "eqs_mc_validate": (C.c_int, [C.POINTER(McParams)]),
    "eqs_mc_integrate": (C.c_int, [_vp, C.POINTER(McParams), dp, dp, C.POINTER(Report)]),
    "eqs_miser_validate": (C.c_int, [C.POINTER(MiserParams)]),
    "eqs_miser_integrate": (C.c_int, [_vp, C.POINTER(MiserParams), dp, dp, C.POINTER(Report)]),
}

_lib = None


def lib():
    """Load the kernel library (built in-tree by `python -m eqsolver_b200.build`); never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KernelLibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -m eqsolver_b200.build` "
                "(nvcc, sm_100a).  eqsolver_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
