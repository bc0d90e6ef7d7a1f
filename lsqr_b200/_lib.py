"""ctypes binding of include/lsqr_b200.h (the C ABI).  No CPU fallback: if the CUDA library is
missing this module raises, and every compute call fails with an LsqrError on a box without a GPU."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LSQR_B200_LIB") or os.path.join(HERE, "lib", "liblsqr_b200.so")

LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


class IterRecord(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "itn", "istop", "x1", "rnorm", "test1", "test2", "anorm", "acond",
        "phi", "dknorm", "dxk", "alfopt", "alpha", "beta", "xnorm", "arnorm")]


ITER_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(IterRecord))
APROD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)
APROD_HOST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double))


class Options(C.Structure):
    """struct lsqr_b200_options"""
    _fields_ = [
        ("atol", C.c_double), ("btol", C.c_double), ("conlim", C.c_double),
        ("itnlim", C.c_int32), ("device", C.c_int32),
        ("stream", C.c_void_p),
        ("log", LOG_FN), ("log_user", C.c_void_p),
        ("iter", ITER_FN), ("iter_user", C.c_void_p),
        ("engine", C.c_int32), ("use_graph", C.c_int32), ("profile", C.c_int32), ("spmv_variant", C.c_int32),
        ("world_size", C.c_int32), ("rank", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
        ("m_global", C.c_int64),
    ]


class KernelTimes(C.Structure):
    """struct lsqr_b200_kernel_times"""
    _fields_ = [
        ("aprod_ms", C.c_double), ("atprod_ms", C.c_double), ("update_ms", C.c_double), ("other_ms", C.c_double),
        ("aprod_launches", C.c_int64), ("atprod_launches", C.c_int64), ("update_launches", C.c_int64),
        ("other_launches", C.c_int64), ("total_launches", C.c_int64),
        ("loop_ms", C.c_double), ("init_ms", C.c_double), ("iteration_launches", C.c_int64),
    ]


class PlanInfo(C.Structure):
    """struct lsqr_b200_plan_info"""
    _fields_ = [
        ("nblocks", C.c_int64), ("ntiles", C.c_int64),
        ("grid_ctas", C.c_int32), ("ctas_per_sm", C.c_int32), ("window_doubles", C.c_int32), ("balanced", C.c_int32),
        ("windowed_fraction", C.c_double), ("imbalance", C.c_double),
        ("span_median", C.c_int64), ("span_max", C.c_int64),
        ("single_launch", C.c_int32), ("peer_exchange", C.c_int32),
        ("entries_per_lane", C.c_int32), ("flavour", C.c_int32), ("lines_per_gather", C.c_double),
    ]


ERR_NAMES = {
    0: "OK", 1: "ERR_SIZES", 2: "ERR_IROW", 3: "ERR_ICOL", 4: "ERR_NOINIT", 5: "ERR_MODE", 6: "ERR_INDEX_LOW",
    10: "ERR_NO_DEVICE", 11: "ERR_CUDA", 12: "ERR_NCCL", 13: "ERR_ARG", 14: "ERR_TOO_LARGE", 15: "ERR_CALLBACK",
}


class LsqrError(RuntimeError):
    """Raised where the reference would `error stop '<message>'` (codes 1..5) or on an engine failure."""

    def __init__(self, code: int, message: str, detail: str = ""):
        self.code = code
        self.name = ERR_NAMES.get(code, str(code))
        self.message = message
        super().__init__(f"{message}" + (f" [{detail}]" if detail else ""))


_lib = None


def load() -> C.CDLL:
    """Loads liblsqr_b200.so; raises ImportError (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine has not been built (python -m lsqr_b200.build). "
            "lsqr_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.lsqr_b200_error_message.restype = C.c_char_p
    L.lsqr_b200_error_message.argtypes = [C.c_int]
    L.lsqr_b200_last_error.restype = C.c_char_p
    L.lsqr_b200_last_error.argtypes = []
    L.lsqr_b200_version.restype = C.c_int
    L.lsqr_b200_device_count.restype = C.c_int
    L.lsqr_b200_default_options.restype = None
    L.lsqr_b200_default_options.argtypes = [C.POINTER(Options)]
    L.lsqr_b200_nccl_unique_id.restype = C.c_int
    L.lsqr_b200_nccl_unique_id.argtypes = [vp]
    L.lsqr_b200_ez_initialize.restype = C.c_int
    L.lsqr_b200_ez_initialize.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32, C.c_int64, vp, C.c_int64, vp,
                                          C.c_int64, vp, C.POINTER(Options)]
    L.lsqr_b200_ez_solve.restype = C.c_int
    L.lsqr_b200_ez_solve.argtypes = [vp, vp, C.c_double, vp, ip, vp, ip, dp, dp, dp, dp, dp]
    L.lsqr_b200_ez_aprod.restype = C.c_int
    L.lsqr_b200_ez_aprod.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]
    L.lsqr_b200_ez_destroy.restype = None
    L.lsqr_b200_ez_destroy.argtypes = [vp]
    L.lsqr_b200_ez_set_options.restype = C.c_int
    L.lsqr_b200_ez_set_options.argtypes = [vp, C.POINTER(Options)]
    L.lsqr_b200_ez_get_csr.restype = C.c_int
    L.lsqr_b200_ez_get_csr.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.lsqr_b200_ez_blocks.restype = C.c_int
    L.lsqr_b200_ez_blocks.argtypes = [vp, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.lsqr_b200_ez_schedule.restype = C.c_int
    L.lsqr_b200_ez_schedule.argtypes = [vp, C.c_int32, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.lsqr_b200_ez_plan.restype = C.c_int
    L.lsqr_b200_ez_plan.argtypes = [vp, C.c_int32, C.POINTER(PlanInfo)]
    L.lsqr_b200_ez_get_csr_device.restype = C.c_int
    L.lsqr_b200_ez_get_csr_device.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.lsqr_b200_ez_nnz.restype = C.c_int64
    L.lsqr_b200_ez_nnz.argtypes = [vp]
    L.lsqr_b200_ez_get_kernel_times.restype = C.c_int
    L.lsqr_b200_ez_get_kernel_times.argtypes = [vp, C.POINTER(KernelTimes)]
    L.lsqr_b200_ez_aprod_device.restype = C.c_int
    L.lsqr_b200_ez_aprod_device.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    L.lsqr_b200_lsqr.restype = C.c_int
    L.lsqr_b200_lsqr.argtypes = [APROD_FN, vp, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                 vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_int32,
                                 C.POINTER(Options), ip, ip, dp, dp, dp, dp, dp]
    L.lsqr_b200_acheck.restype = C.c_int
    L.lsqr_b200_acheck.argtypes = [APROD_FN, vp, C.c_int32, C.c_int32, C.c_double, vp, vp, vp, vp,
                                   C.POINTER(Options), ip, dp]
    L.lsqr_b200_xcheck.restype = C.c_int
    L.lsqr_b200_xcheck.argtypes = [APROD_FN, vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                   vp, vp, vp, vp, vp, C.POINTER(Options), ip, dp, dp, dp, dp]
    L.lsqr_b200_lsqr_host.restype = C.c_int
    L.lsqr_b200_lsqr_host.argtypes = [APROD_HOST_FN, vp, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                      vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_int32,
                                      C.POINTER(Options), ip, ip, dp, dp, dp, dp, dp]
    L.lsqr_b200_acheck_host.restype = C.c_int
    L.lsqr_b200_acheck_host.argtypes = [APROD_HOST_FN, vp, C.c_int32, C.c_int32, C.c_double, vp, vp, vp, vp,
                                        C.POINTER(Options), ip, dp]
    L.lsqr_b200_xcheck_host.restype = C.c_int
    L.lsqr_b200_xcheck_host.argtypes = [APROD_HOST_FN, vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                        vp, vp, vp, vp, vp, C.POINTER(Options), ip, dp, dp, dp, dp]
    L.lsqr_b200_dnrm2.restype = C.c_int
    L.lsqr_b200_dnrm2.argtypes = [C.c_int64, vp, dp, vp]
    L.lsqr_b200_ddot.restype = C.c_int
    L.lsqr_b200_ddot.argtypes = [C.c_int64, vp, vp, dp, vp]
    L.lsqr_b200_dscal.restype = C.c_int
    L.lsqr_b200_dscal.argtypes = [C.c_int64, C.c_double, vp, vp]
    L.lsqr_b200_dcopy.restype = C.c_int
    L.lsqr_b200_dcopy.argtypes = [C.c_int64, vp, vp, vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        L = load()
        raise LsqrError(rc, L.lsqr_b200_error_message(rc).decode(), L.lsqr_b200_last_error().decode())


def default_options() -> Options:
    o = Options()
    load().lsqr_b200_default_options(C.byref(o))
    return o
