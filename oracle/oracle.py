"""ctypes front end of the CPU ORACLE (test infrastructure, NOT product code).

Loads ``oracle/build/liblsqr_oracle.so`` (built by ``oracle/Makefile``), the plain-C
restatement of the reference's ``src/lsqr.f90`` / ``src/lsqrblas.f90`` /
``test/lsqrtest_module.f90``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and
only as the checker / the reported CPU baseline.  Nothing under ``lsqr_b200/`` imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liblsqr_oracle.so")

FOURPI_F64 = 0  # current source: 4*acos(-1)           (test/lsqrtest_module.f90:433)
FOURPI_F32 = 1  # what the committed LSQR.LIS was made with (single-precision 4.0*3.141592)

APROD_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))
LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


class IterRec(C.Structure):
    _fields_ = [("itn", C.c_int)] + [
        (k, C.c_double)
        for k in ("x1", "rnorm", "test1", "test2", "anorm", "acond", "phi", "dknorm", "dxk", "alfopt",
                  "alpha", "beta", "xnorm", "arnorm")
    ]


ITER_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(IterRec))


class EzOpts(C.Structure):
    _fields_ = [("atol", C.c_double), ("btol", C.c_double), ("conlim", C.c_double),
                ("itnlim", C.c_int), ("has_log", C.c_int)]


class LstpResult(C.Structure):
    _fields_ = [
        ("gen_acond", C.c_double), ("gen_rnorm", C.c_double),
        ("acheck_inform", C.c_int), ("acheck_relerr", C.c_double),
        ("istop", C.c_int), ("itn", C.c_int),
        ("anorm", C.c_double), ("acond", C.c_double), ("rnorm", C.c_double),
        ("arnorm", C.c_double), ("xnorm", C.c_double),
        ("xcheck_inform", C.c_int), ("xtest1", C.c_double), ("xtest2", C.c_double), ("xtest3", C.c_double),
        ("xnorms", C.c_double * 6),
        ("enorm", C.c_double),
        ("x_head", C.c_double * 8),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


_BUILD_FLAGS = "-O2 -ffp-contract=off"


def use_native_build() -> str:
    """For the TIMED CPU baseline only: compile a copy of the oracle for the CPU of the box it runs on
    (gcc -O3 -march=native -ffp-contract=off) and make lib() load it.  Call before the first
    use of the oracle in a process.  Returns the flags in use; falls back to the portable -O2 build if the compile
    fails or the library is already loaded."""
    global _LIB_PATH, _BUILD_FLAGS
    if _lib is not None:
        return _BUILD_FLAGS          # (already loaded: report what was loaded)
    flags = _BUILD_FLAGS
    native = os.path.join(_HERE, "build", "liblsqr_oracle_native.so")
    try:
        subprocess.run(["make", "-B", "-C", _HERE, "native"], check=True, capture_output=True)   # always for THIS CPU
        if os.path.exists(native):
            _LIB_PATH = native
            flags = "-O3 -march=native -ffp-contract=off"
    except Exception:
        pass
    _BUILD_FLAGS = flags
    return flags


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.oracle_dnrm2.restype = C.c_double
        L.oracle_dnrm2.argtypes = [C.c_int, dp, C.c_int]
        L.oracle_ddot.restype = C.c_double
        L.oracle_ddot.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int]
        L.oracle_dscal.restype = None
        L.oracle_dscal.argtypes = [C.c_int, C.c_double, dp, C.c_int]
        L.oracle_dcopy.restype = None
        L.oracle_dcopy.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int]
        L.oracle_d2norm.restype = C.c_double
        L.oracle_d2norm.argtypes = [C.c_double, C.c_double]
        L.oracle_error_message.restype = C.c_char_p
        L.oracle_error_message.argtypes = [C.c_int]
        L.oracle_lsqr.restype = None
        L.oracle_lsqr.argtypes = [APROD_FN, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int,
                                  dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                  LOG_FN, C.c_void_p, ITER_FN, C.c_void_p,
                                  ip, ip, dp, dp, dp, dp, dp]
        L.oracle_acheck.restype = None
        L.oracle_acheck.argtypes = [APROD_FN, C.c_void_p, C.c_int, C.c_int, LOG_FN, C.c_void_p, C.c_double,
                                    dp, dp, dp, dp, ip, dp]
        L.oracle_xcheck.restype = None
        L.oracle_xcheck.argtypes = [APROD_FN, C.c_void_p, C.c_int, C.c_int, LOG_FN, C.c_void_p,
                                    C.c_double, C.c_double, C.c_double,
                                    dp, dp, dp, dp, dp, ip, dp, dp, dp, dp]
        L.oracle_ez_initialize.restype = C.c_int
        L.oracle_ez_initialize.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                           C.c_int64, dp, C.c_int64, C.POINTER(C.c_int32),
                                           C.c_int64, C.POINTER(C.c_int32), C.POINTER(EzOpts)]
        L.oracle_ez_destroy.restype = None
        L.oracle_ez_destroy.argtypes = [C.c_void_p]
        L.oracle_ez_aprod.restype = C.c_int
        L.oracle_ez_aprod.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp]
        L.oracle_ez_solve.restype = None
        L.oracle_ez_solve.argtypes = [C.c_void_p, dp, C.c_double, dp, ip, dp, ip, dp, dp, dp, dp, dp,
                                      LOG_FN, C.c_void_p, ITER_FN, C.c_void_p]
        L.oracle_lstp_test.restype = None
        L.oracle_lstp_test.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                       LOG_FN, C.c_void_p, ITER_FN, C.c_void_p,
                                       C.POINTER(LstpResult), dp]
        L.oracle_lstp_new.restype = C.c_void_p
        L.oracle_lstp_new.argtypes = [C.c_int, C.c_int]
        L.oracle_lstp_free.restype = None
        L.oracle_lstp_free.argtypes = [C.c_void_p]
        L.oracle_lstp_generate.restype = None
        L.oracle_lstp_generate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp]
        L.oracle_lstp_aprod.restype = None
        L.oracle_lstp_aprod.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp]
        L.oracle_coo_to_csr.restype = None
        L.oracle_coo_to_csr.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp,
                                        C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32), dp,
                                        C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a, copy=False) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    return out.copy() if copy and out is a else out


# ---------------------------------------------------------------------------
# BLAS-1 and d2norm
# ---------------------------------------------------------------------------
def dnrm2(x) -> float:
    x = _f64(x)
    return lib().oracle_dnrm2(int(x.size), _dp(x), 1)


def ddot(x, y) -> float:
    x, y = _f64(x), _f64(y)
    return lib().oracle_ddot(int(x.size), _dp(x), 1, _dp(y), 1)


def dscal(a: float, x: np.ndarray) -> None:
    assert x.dtype == np.float64 and x.flags.c_contiguous
    lib().oracle_dscal(int(x.size), float(a), _dp(x), 1)


def d2norm(a: float, b: float) -> float:
    return lib().oracle_d2norm(float(a), float(b))


# ---------------------------------------------------------------------------
# results
# ---------------------------------------------------------------------------
@dataclass
class LsqrResult:
    x: np.ndarray
    istop: int
    itn: int
    anorm: float
    acond: float
    rnorm: float
    arnorm: float
    xnorm: float
    se: np.ndarray | None = None
    log: list = field(default_factory=list)
    trace: list = field(default_factory=list)   # per-iteration dicts


class _Capture:
    """Keeps the ctypes callbacks alive and collects log lines / iteration records."""

    def __init__(self, want_log: bool, want_trace: bool):
        self.lines: list[str] = []
        self.trace: list[dict] = []
        self.log_cb = LOG_FN(self._on_log) if want_log else C.cast(None, LOG_FN)
        self.iter_cb = ITER_FN(self._on_iter) if want_trace else C.cast(None, ITER_FN)

    def _on_log(self, _user, line):
        self.lines.append(line.decode())

    def _on_iter(self, _user, rec):
        r = rec.contents
        self.trace.append({k: getattr(r, k) for k, _ in IterRec._fields_})


class OracleError(RuntimeError):
    """Stands for the reference's ``error stop '<message>'``."""

    def __init__(self, code: int):
        self.code = code
        super().__init__(lib().oracle_error_message(code).decode())


# ---------------------------------------------------------------------------
# lsqr_solver_ez  (src/lsqr.f90:32-65)
# ---------------------------------------------------------------------------
class SolverEz:
    """Oracle twin of ``lsqr_solver_ez``: ``initialize`` happens in the constructor
    (src/lsqr.f90:91-127), ``solve`` follows src/lsqr.f90:207-259, ``aprod`` :134-200."""

    def __init__(self, m, n, a, irow, icol, atol=0.0, btol=0.0, conlim=0.0, itnlim=100):
        self._a = _f64(a)
        self._irow = np.ascontiguousarray(irow, dtype=np.int32)
        self._icol = np.ascontiguousarray(icol, dtype=np.int32)
        self.m, self.n = int(m), int(n)
        opts = EzOpts(atol, btol, conlim, itnlim, 0)
        h = C.c_void_p()
        rc = lib().oracle_ez_initialize(C.byref(h), self.m, self.n,
                                        self._a.size, _dp(self._a),
                                        self._irow.size, _i32p(self._irow),
                                        self._icol.size, _i32p(self._icol), C.byref(opts))
        if rc != 0:
            raise OracleError(rc)
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().oracle_ez_destroy(h)
            self._h = None

    def aprod(self, mode: int, x: np.ndarray, y: np.ndarray, m=None, n=None) -> None:
        rc = lib().oracle_ez_aprod(self._h, mode, self.m if m is None else m, self.n if n is None else n,
                                   _dp(x), _dp(y))
        if rc != 0:
            raise OracleError(rc)

    def solve(self, b, damp=0.0, wantse=False, log=False, trace=False) -> LsqrResult:
        b = _f64(b)
        assert b.size == self.m
        x = np.zeros(max(self.n, 1))
        se = np.zeros(max(self.n, 1)) if wantse else None
        istop, itn = C.c_int(), C.c_int()
        sc = [C.c_double() for _ in range(5)]
        cap = _Capture(log, trace)
        lib().oracle_ez_solve(self._h, _dp(b), float(damp), _dp(x), C.byref(istop),
                              _dp(se) if wantse else None, C.byref(itn),
                              *[C.byref(s) for s in sc], cap.log_cb, None, cap.iter_cb, None)
        return LsqrResult(x[: self.n], istop.value, itn.value, *[s.value for s in sc],
                          se=se[: self.n] if wantse else None, log=cap.lines, trace=cap.trace)


# ---------------------------------------------------------------------------
# low-level lsqr with a Python operator (src/lsqr.f90:432-882)
# ---------------------------------------------------------------------------
def _wrap_aprod(fn, m, n):
    def thunk(_user, mode, m_, n_, xp, yp):
        x = np.ctypeslib.as_array(xp, shape=(n_,))
        y = np.ctypeslib.as_array(yp, shape=(m_,))
        fn(mode, m_, n_, x, y)
    return APROD_FN(thunk)


def lsqr(aprod, m, n, b, damp=0.0, wantse=False, atol=0.0, btol=0.0, conlim=0.0, itnlim=100,
         log=False, trace=False) -> LsqrResult:
    """``aprod(mode, m, n, x, y)`` updates y in place for mode 1 (y += A x) and x for mode 2."""
    u = _f64(b, copy=True).copy()
    v, w, x = np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
    se = np.zeros(max(n, 1))
    istop, itn = C.c_int(), C.c_int()
    sc = [C.c_double() for _ in range(5)]
    cap = _Capture(log, trace)
    cb = _wrap_aprod(aprod, m, n)
    lib().oracle_lsqr(cb, None, m, n, float(damp), int(wantse), _dp(u), _dp(v), _dp(w), _dp(x), _dp(se),
                      float(atol), float(btol), float(conlim), int(itnlim),
                      cap.log_cb, None, cap.iter_cb, None,
                      C.byref(istop), C.byref(itn), *[C.byref(s) for s in sc])
    return LsqrResult(x[:n], istop.value, itn.value, *[s.value for s in sc],
                      se=se[:n] if wantse else None, log=cap.lines, trace=cap.trace)


def acheck(aprod, m, n, eps=np.finfo(np.float64).eps):
    v, x = np.zeros(max(n, 1)), np.zeros(max(n, 1))
    w, y = np.zeros(max(m, 1)), np.zeros(max(m, 1))
    inform, rel = C.c_int(), C.c_double()
    cb = _wrap_aprod(aprod, m, n)
    lib().oracle_acheck(cb, None, m, n, C.cast(None, LOG_FN), None, float(eps),
                        _dp(v), _dp(w), _dp(x), _dp(y), C.byref(inform), C.byref(rel))
    return inform.value, rel.value


def xcheck(aprod, m, n, anorm, damp, b, x, eps=np.finfo(np.float64).eps):
    """Returns dict(inform, test1..3, bnorm, xnorm, rho1, sigma1, rho2, sigma2, r, Atr, w)."""
    b, x = _f64(b), _f64(x)
    u, v, w = np.zeros(max(m, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
    inform = C.c_int()
    t = [C.c_double() for _ in range(3)]
    norms = (C.c_double * 6)()
    cb = _wrap_aprod(aprod, m, n)
    lib().oracle_xcheck(cb, None, m, n, C.cast(None, LOG_FN), None, float(anorm), float(damp), float(eps),
                        _dp(b), _dp(u), _dp(v), _dp(w), _dp(x), C.byref(inform),
                        *[C.byref(s) for s in t], norms)
    keys = ("bnorm", "xnorm", "rho1", "sigma1", "rho2", "sigma2")
    out = dict(inform=inform.value, test1=t[0].value, test2=t[1].value, test3=t[2].value,
               r=u[:m], Atr=v[:n], w=w[:n])
    out.update({k: norms[i] for i, k in enumerate(keys)})
    return out


# ---------------------------------------------------------------------------
# LSTP problems (test/lsqrtest_module.f90)
# ---------------------------------------------------------------------------
def lstp_test(m, n, nduplc, npower, damp, fourpi_mode=FOURPI_F64, log=False, trace=False):
    res = LstpResult()
    cap = _Capture(log, trace)
    x = np.zeros(max(n, 1))
    lib().oracle_lstp_test(m, n, nduplc, npower, float(damp), fourpi_mode,
                           cap.log_cb, None, cap.iter_cb, None, C.byref(res), _dp(x))
    out = {k: getattr(res, k) for k, _ in LstpResult._fields_ if k not in ("xnorms", "x_head")}
    out["xnorms"] = list(res.xnorms)
    out["x_head"] = list(res.x_head)
    out["x"] = x[:n]
    out["log"] = cap.lines
    out["trace"] = cap.trace
    return out


class Lstp:
    """The LSTP operator A = HY*D*HZ plus its generated right-hand side."""

    def __init__(self, m, n, nduplc, npower, damp, fourpi_mode=FOURPI_F64):
        self.m, self.n, self.damp = m, n, damp
        self._h = lib().oracle_lstp_new(m, n)
        self.xtrue = np.arange(1, n + 1, dtype=np.float64) * 0.1
        self.b = np.zeros(m)
        ac, rn = C.c_double(), C.c_double()
        lib().oracle_lstp_generate(self._h, nduplc, npower, float(damp), fourpi_mode,
                                   _dp(self.xtrue), _dp(self.b), C.byref(ac), C.byref(rn))
        self.acond, self.rnorm = ac.value, rn.value
        minmn = min(m, n)
        # the struct starts with 4 ints then d, hy, hz, w pointers
        class _S(C.Structure):
            _fields_ = [("m", C.c_int), ("n", C.c_int), ("maxmn", C.c_int), ("minmn", C.c_int),
                        ("d", C.POINTER(C.c_double)), ("hy", C.POINTER(C.c_double)),
                        ("hz", C.POINTER(C.c_double)), ("w", C.POINTER(C.c_double))]
        s = C.cast(self._h, C.POINTER(_S)).contents
        self.d = np.ctypeslib.as_array(s.d, shape=(minmn,)).copy()
        self.hy = np.ctypeslib.as_array(s.hy, shape=(m,)).copy()
        self.hz = np.ctypeslib.as_array(s.hz, shape=(n,)).copy()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().oracle_lstp_free(h)
            self._h = None

    def aprod(self, mode, m, n, x, y):
        lib().oracle_lstp_aprod(self._h, mode, m, n, _dp(x), _dp(y))


# ---------------------------------------------------------------------------
# COO -> CSR host reference
# ---------------------------------------------------------------------------
def coo_to_csr(nkeys, irow, icol, a, by_col=False):
    irow = np.ascontiguousarray(irow, dtype=np.int32)
    icol = np.ascontiguousarray(icol, dtype=np.int32)
    a = _f64(a)
    nnz = a.size
    ptr = np.zeros(nkeys + 1, dtype=np.int64)
    idx = np.zeros(max(nnz, 1), dtype=np.int32)
    val = np.zeros(max(nnz, 1), dtype=np.float64)
    perm = np.zeros(max(nnz, 1), dtype=np.int64)
    lib().oracle_coo_to_csr(nkeys, nnz, _i32p(irow), _i32p(icol), _dp(a), int(by_col),
                            ptr.ctypes.data_as(C.POINTER(C.c_int64)), _i32p(idx), _dp(val),
                            perm.ctypes.data_as(C.POINTER(C.c_int64)))
    return ptr, idx[:nnz], val[:nnz], perm[:nnz]
