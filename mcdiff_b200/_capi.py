"""ctypes view of include/mcdiff_b200.h and the loader of libmcdiff_b200.so.

The product path has no CPU fallback: if the shared library is missing or does not load,
`load()` raises and nothing else in the package can compute.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCDIFF_B200_LIB", os.path.join(_HERE, "libmcdiff_b200.so"))

MCD_NSCAL = 8
MCD_NCOUNT = 8
S_LL, S_T0, S_DV, S_DW, S_DT0, S_TEMP = range(6)
C_NACCV, C_NACCW, C_NACCT0, C_NACCV_UPD, C_NACCW_UPD, C_STEP, C_STATUS, C_SWEEPS = range(8)
ST_NONFINITE, ST_NAN_TRIAL, ST_BAD_LAGS = 1, 2, 4
ST_EIG_NOCONV = ST_NONFINITE
METHOD_AUTO, METHOD_EIGEN, METHOD_SQUARING, METHOD_SQUARING_DFMA = 0, 1, 2, 3

c_dp = C.c_void_p  # device (or, for the test emulator, host) pointers are passed as integers


class ProblemDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("pbc", C.c_int32), ("nlag", C.c_int32), ("ncosF", C.c_int32),
        ("ncosD", C.c_int32), ("use_pull", C.c_int32),
        ("lagtimes", c_dp), ("counts", c_dp), ("v_basis", c_dp), ("w_basis", c_dp),
        ("string_vecs", c_dp), ("spring_k", C.c_double), ("clip", C.c_double), ("pull", C.c_double),
    ]


class McParams(C.Structure):
    _fields_ = [
        ("num_mc_update", C.c_double), ("dtemp", C.c_double), ("min_lt", C.c_double),
        ("move_timezero", C.c_int32), ("sample_every", C.c_int32), ("refresh_every", C.c_int32),
        ("use_tape", C.c_int32), ("method", C.c_int32), ("taylor_m", C.c_int32),
        ("seed", C.c_uint64), ("chain_offset", C.c_int64),
        ("tape_stride", C.c_int64), ("nsamples_cap", C.c_int64),
    ]


class ChainIO(C.Structure):
    _fields_ = [
        ("v", c_dp), ("w", c_dp), ("vcoef", c_dp), ("wcoef", c_dp), ("scal", c_dp),
        ("counters", c_dp), ("naccv_coeff", c_dp), ("naccw_coeff", c_dp), ("basis", c_dp),
        ("basis_valid", c_dp), ("tape_u", c_dp), ("tape_idx", c_dp), ("decisions", c_dp),
        ("ll_try", c_dp), ("samples", c_dp),
    ]


class RadProblemDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("pbc", C.c_int32), ("nlag", C.c_int32), ("dim_rad", C.c_int32), ("lmax", C.c_int32),
        ("ncosDrad", C.c_int32),
        ("v", c_dp), ("w", c_dp), ("lagtimes", c_dp), ("counts", c_dp), ("bessels", c_dp), ("b0zeros", c_dp),
        ("wrad_basis", c_dp), ("clip", C.c_double),
    ]


class RadMcParams(C.Structure):
    _fields_ = [
        ("num_mc_update", C.c_double), ("dtemp", C.c_double), ("sample_every", C.c_int32),
        ("use_tape", C.c_int32), ("ncosDrad", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64),
        ("chain_offset", C.c_int64), ("tape_stride", C.c_int64), ("nsamples_cap", C.c_int64),
    ]


class RadChainIO(C.Structure):
    _fields_ = [
        ("wrad", c_dp), ("wcoef", c_dp), ("scal", c_dp), ("counters", c_dp), ("nacc_coeff", c_dp),
        ("wrad_basis", c_dp), ("tape_u", c_dp), ("tape_idx", c_dp), ("decisions", c_dp), ("ll_try", c_dp),
        ("samples", c_dp),
    ]


EXPORTS = (
    "mcd_version", "mcd_error_string", "mcd_create", "mcd_destroy", "mcd_problem_create",
    "mcd_problem_destroy", "mcd_problem_info", "mcd_loglike_batch", "mcd_mc_run",
    "mcd_rad_problem_create", "mcd_rad_problem_destroy", "mcd_rad_problem_info",
    "mcd_rad_loglike_batch", "mcd_rad_mc_run", "mcd_propagator_batch", "mcd_traj_counts", "mcd_traj_counts_rad",
    "mcd_fp64_peak", "mcd_launch_count",
)

_lib = None


class McdError(RuntimeError):
    pass


def load():
    """Load libmcdiff_b200.so (built by __graft_entry__.build()).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McdError(
            "libmcdiff_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "mcdiff_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.mcd_version.restype = C.c_int
    lib.mcd_error_string.restype = C.c_char_p
    lib.mcd_error_string.argtypes = [C.c_int]
    lib.mcd_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.mcd_destroy.argtypes = [C.c_void_p]
    lib.mcd_problem_create.argtypes = [C.c_void_p, C.POINTER(ProblemDesc), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mcd_problem_destroy.argtypes = [C.c_void_p]
    lib.mcd_problem_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.mcd_loglike_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                      C.c_void_p]
    lib.mcd_mc_run.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(McParams), C.POINTER(ChainIO), C.c_int32,
                               C.c_int32, C.c_void_p]
    lib.mcd_rad_problem_create.argtypes = [C.c_void_p, C.POINTER(RadProblemDesc), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mcd_rad_problem_destroy.argtypes = [C.c_void_p]
    lib.mcd_rad_problem_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
    lib.mcd_rad_loglike_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_dp, C.c_void_p]
    lib.mcd_rad_mc_run.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RadMcParams), C.POINTER(RadChainIO), C.c_int32,
                                   C.c_int32, C.c_void_p]
    lib.mcd_propagator_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_void_p]
    lib.mcd_traj_counts.argtypes = [C.c_void_p, c_dp, C.c_int64, C.c_int64, C.c_double, c_dp, C.c_int32,
                                    C.POINTER(C.c_int32), C.c_int32, c_dp, C.c_void_p]
    lib.mcd_traj_counts_rad.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, C.c_int64, C.c_int64, C.c_double, c_dp, C.c_int32,
                                        c_dp, C.c_int32, C.POINTER(C.c_int32), C.c_int32, c_dp, C.c_void_p]
    lib.mcd_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.mcd_launch_count.restype = C.c_int64
    lib.mcd_launch_count.argtypes = [C.c_void_p]
    for f in ("mcd_create", "mcd_destroy", "mcd_problem_create", "mcd_problem_destroy", "mcd_problem_info",
              "mcd_loglike_batch", "mcd_mc_run", "mcd_rad_problem_create", "mcd_rad_problem_destroy",
              "mcd_rad_problem_info", "mcd_rad_loglike_batch", "mcd_rad_mc_run", "mcd_propagator_batch",
              "mcd_traj_counts", "mcd_traj_counts_rad", "mcd_fp64_peak"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().mcd_error_string(code).decode()
        raise McdError("%s failed: %s (code %d)" % (what or "mcdiff_b200 call", msg, code))
