"""Host-side mirror of the reference's public interface (src/lsqr.f90:16-65) over the C ABI.

    reference (Fortran 2008)                          here
    ------------------------------------------------  -------------------------------------------
    type(lsqr_solver_ez) :: s                         s = LsqrSolverEz()
    call s%initialize(m,n,a,irow,icol,atol,...,nout)  s.initialize(m, n, a, irow, icol, atol=..., nout=...)
    call s%solve(b,damp,x,istop,se,itn,anorm,...)     r = s.solve(b, damp, want_se=...)  # r.x, r.istop, ...
    call s%aprod(mode,m,n,x,y)                        s.aprod(mode, m, n, x, y)
    type,extends(lsqr_solver) :: my ; aprod => ...    class My(LsqrSolver): def aprod(self, mode, m, n, x, y, stream)
    call my%lsqr(m,n,damp,wantse,u,v,w,x,se,...)      my.lsqr(m, n, damp, wantse, u, v, w, x, se, atol, ...)
    call my%acheck(...) / my%xcheck(...)              my.acheck(...) / my.xcheck(...)

Index arrays are the reference's: 1-based int32.  Arrays may be numpy arrays (host) or anything with
``data_ptr()`` (torch CUDA / CPU tensors).  Where the reference would ``error stop '<msg>'`` an
``LsqrError`` carrying the same message is raised.  All arithmetic happens on the GPU through
``liblsqr_b200.so``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import LsqrError, Options, KernelTimes, PlanInfo


# --------------------------------------------------------------------------------------------
# array plumbing
# --------------------------------------------------------------------------------------------
def _is_tensor(a) -> bool:
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda")


def _as_f64(a, name: str):
    """Returns (keepalive, pointer, length) of a float64 contiguous view of a."""
    if _is_tensor(a):
        import torch
        if a.dtype != torch.float64 or not a.is_contiguous():
            a = a.to(torch.float64).contiguous()
        return a, a.data_ptr(), a.numel()
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return arr, arr.ctypes.data, arr.size


def _as_i32(a, name: str):
    if _is_tensor(a):
        import torch
        if a.dtype != torch.int32 or not a.is_contiguous():
            a = a.to(torch.int32).contiguous()
        return a, a.data_ptr(), a.numel()
    arr = np.ascontiguousarray(a, dtype=np.int32)
    return arr, arr.ctypes.data, arr.size


def _ptr(a) -> int:
    """Raw pointer of an array that is passed by reference and written in place."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if _is_tensor(a):
        import torch
        if a.dtype != torch.float64 or not a.is_contiguous():
            raise TypeError("in/out tensors must be contiguous float64")
        return a.data_ptr()
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
        raise TypeError("in/out arrays must be contiguous float64 numpy arrays or tensors")
    return a.ctypes.data


@dataclass
class SolveResult:
    """Outputs of ``solve`` / ``lsqr`` (src/lsqr.f90:207-208, 432-435)."""
    x: object
    istop: int
    itn: int
    anorm: float
    acond: float
    rnorm: float
    arnorm: float
    xnorm: float
    se: object = None
    log: list = field(default_factory=list)
    trace: list = field(default_factory=list)


class _Callbacks:
    """Keeps the ctypes trampolines of one options struct alive."""

    def __init__(self, nout=None, trace: Optional[list] = None):
        self.lines: list[str] = []
        self.trace = trace
        self._nout = nout

        def on_log(_u, line):
            s = line.decode()
            self.lines.append(s)
            if callable(self._nout):
                self._nout(s)
            elif hasattr(self._nout, "write"):
                self._nout.write(s + "\n")

        def on_iter(_u, rec):
            r = rec.contents
            self.trace.append({k: getattr(r, k) for k, _ in _lib.IterRecord._fields_})

        self.log_cb = _lib.LOG_FN(on_log) if nout is not None else C.cast(None, _lib.LOG_FN)
        self.iter_cb = _lib.ITER_FN(on_iter) if trace is not None else C.cast(None, _lib.ITER_FN)

    def install(self, o: Options) -> None:
        o.log = self.log_cb
        o.iter = self.iter_cb


# --------------------------------------------------------------------------------------------
# lsqr_solver_ez
# --------------------------------------------------------------------------------------------
class LsqrSolverEz:
    """``type(lsqr_solver_ez)`` (src/lsqr.f90:32-65): the matrix is given as COO triplets."""

    def __init__(self):
        self._h = C.c_void_p()
        self.m = 0
        self.n = 0
        self._opts: Optional[Options] = None
        self._cb: Optional[_Callbacks] = None

    # initialize_ez, src/lsqr.f90:91-127
    def initialize(self, m: int, n: int, a, irow, icol, atol: float = 0.0, btol: float = 0.0,
                   conlim: float = 0.0, itnlim: int = 100, nout=None, *,
                   device: int = -1, stream: int = 0, engine: int = 0, use_graph: bool = True,
                   profile: bool = False, spmv_variant: int = 0, world_size: int = 1, rank: int = 0,
                   nccl_unique_id: Optional[bytes] = None, m_global: int = 0) -> "LsqrSolverEz":
        L = _lib.load()
        self.destroy()                       # `me` is intent(out): re-initialising resets the object (:95)
        ka, pa, na = _as_f64(a, "a")
        kr, pr, nr = _as_i32(irow, "irow")
        kc, pc, nc = _as_i32(icol, "icol")
        o = _lib.default_options()
        o.atol, o.btol, o.conlim, o.itnlim = float(atol), float(btol), float(conlim), int(itnlim)
        o.device, o.stream = int(device), int(stream) or None
        o.engine, o.use_graph, o.profile = int(engine), int(bool(use_graph)), int(bool(profile))
        o.spmv_variant = int(spmv_variant)
        o.world_size, o.rank, o.m_global = int(world_size), int(rank), int(m_global)
        idbuf = None
        if nccl_unique_id is not None:
            idbuf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            o.nccl_unique_id = C.cast(idbuf, C.c_void_p)
        self._cb = _Callbacks(nout=nout)
        self._cb.install(o)
        h = C.c_void_p()
        rc = L.lsqr_b200_ez_initialize(C.byref(h), int(m), int(n), na, pa, nr, pr, nc, pc, C.byref(o))
        del ka, kr, kc, idbuf
        _lib.check(rc)
        self._h, self.m, self.n, self._opts = h, int(m), int(n), o
        return self

    def set_tolerances(self, atol=None, btol=None, conlim=None, itnlim=None, nout="keep", trace=None,
                       engine=None, use_graph=None, profile=None) -> None:
        """Changes the optional arguments of ``initialize`` without rebuilding the matrix."""
        o = self._opts
        if atol is not None: o.atol = float(atol)
        if btol is not None: o.btol = float(btol)
        if conlim is not None: o.conlim = float(conlim)
        if itnlim is not None: o.itnlim = int(itnlim)
        if engine is not None: o.engine = int(engine)
        if use_graph is not None: o.use_graph = int(bool(use_graph))
        if profile is not None: o.profile = int(bool(profile))
        if nout != "keep" or trace is not None:
            self._cb = _Callbacks(nout=None if nout == "keep" else nout, trace=trace)
            self._cb.install(o)
        _lib.check(_lib.load().lsqr_b200_ez_set_options(self._h, C.byref(o)))

    # solve_ez, src/lsqr.f90:207-259
    def solve(self, b, damp: float = 0.0, want_se: bool = False, x=None, se=None, trace: bool = False) -> SolveResult:
        if not self._h:
            raise LsqrError(4, "lsqr_solver_ez class not properly initialized")
        L = _lib.load()
        kb, pb, nb = _as_f64(b, "b")
        if nb != self.m:
            raise LsqrError(13, "invalid argument", f"b has {nb} entries, expected m = {self.m}")
        on_device = _is_tensor(b) and b.is_cuda
        if x is None:
            if on_device:
                import torch
                x = torch.empty(self.n, dtype=torch.float64, device=b.device)
            else:
                x = np.empty(self.n, dtype=np.float64)
        if want_se and se is None:
            if on_device:
                import torch
                se = torch.empty(self.n, dtype=torch.float64, device=b.device)
            else:
                se = np.empty(self.n, dtype=np.float64)
        tr: Optional[list] = [] if trace else None
        if trace:
            keep_nout = self._cb._nout if self._cb else None
            self._cb = _Callbacks(nout=keep_nout, trace=tr)
            self._cb.install(self._opts)
            _lib.check(L.lsqr_b200_ez_set_options(self._h, C.byref(self._opts)))
        elif self._cb is not None:
            self._cb.lines.clear()
        istop, itn = C.c_int32(), C.c_int32()
        sc = [C.c_double() for _ in range(5)]
        rc = L.lsqr_b200_ez_solve(self._h, pb, float(damp), _ptr(x), C.byref(istop),
                                  _ptr(se) if want_se else None, C.byref(itn), *[C.byref(s) for s in sc])
        del kb
        _lib.check(rc)
        return SolveResult(x, istop.value, itn.value, *[s.value for s in sc], se=se if want_se else None,
                           log=list(self._cb.lines) if self._cb else [], trace=tr or [])

    # aprod_ez, src/lsqr.f90:134-200:  mode 1: y += A x ; mode 2: x += A'y  (in place)
    def aprod(self, mode: int, m: int, n: int, x, y) -> None:
        if not self._h:
            raise LsqrError(4, "lsqr_solver_ez class not properly initialized")
        _lib.check(_lib.load().lsqr_b200_ez_aprod(self._h, int(mode), int(m), int(n), _ptr(x), _ptr(y)))

    def aprod_device(self, mode: int, m: int, n: int, x, y, stream: int = 0) -> None:
        """aprod on DEVICE vectors, enqueued on `stream` without synchronising (lsqr_b200_ez_aprod_device)."""
        if not self._h:
            raise LsqrError(4, "lsqr_solver_ez class not properly initialized")
        _lib.check(_lib.load().lsqr_b200_ez_aprod_device(self._h, int(mode), int(m), int(n), _ptr(x), _ptr(y),
                                                        int(stream) or None))

    # parity inspection
    @property
    def nnz(self) -> int:
        return int(_lib.load().lsqr_b200_ez_nnz(self._h))

    def blocks(self, transpose: bool = False):
        """(nblocks, block_size) of the stored A (column-blocked when v does not fit in L2) or A' (row-blocked
        when u does not)."""
        nb, bs = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().lsqr_b200_ez_blocks(self._h, int(transpose), C.byref(nb), C.byref(bs)))
        return nb.value, bs.value

    def schedule(self, transpose: bool = False, block: int = 0) -> dict:
        """Work schedule of the SpMV kernel over one block of A (or A'): tiles, entries per tile, whether the
        balanced (largest-first) schedule for uneven rows is in use, most-loaded-warp / mean load."""
        nt, te, bal, imb = C.c_int64(), C.c_int64(), C.c_int32(), C.c_double()
        _lib.check(_lib.load().lsqr_b200_ez_schedule(self._h, int(transpose), int(block), C.byref(nt), C.byref(te),
                                                     C.byref(bal), C.byref(imb)))
        return {"ntiles": nt.value, "tile_entries": te.value, "balanced": bool(bal.value), "imbalance": imb.value}

    def plan(self, transpose: bool = False) -> dict:
        """Work plan of the SpMV kernel over the stored A (or A'): blocks, tiles, persistent grid, kernel flavour,
        shared-memory gather window and the fraction of the entries it serves, schedule balance."""
        p = PlanInfo()
        _lib.check(_lib.load().lsqr_b200_ez_plan(self._h, int(transpose), C.byref(p)))
        return {k: getattr(p, k) for k, _ in PlanInfo._fields_}

    def transpose_blocks(self):
        return self.blocks(True)

    def get_csr(self, transpose: bool = False):
        """Host copies (ptr, idx, val, perm) of the device-built CSR of A (or of A').  For a blocked layout ptr
        has nblocks*nkeys + 1 entries (see ``blocks``)."""
        nkeys = (self.n if transpose else self.m) * self.blocks(transpose)[0]
        nnz = self.nnz
        ptr = np.zeros(nkeys + 1, np.int64)
        idx = np.zeros(max(nnz, 1), np.int32)
        val = np.zeros(max(nnz, 1), np.float64)
        perm = np.zeros(max(nnz, 1), np.int64)
        _lib.check(_lib.load().lsqr_b200_ez_get_csr(self._h, int(transpose), ptr.ctypes.data, idx.ctypes.data,
                                                   val.ctypes.data, perm.ctypes.data))
        return ptr, idx[:nnz], val[:nnz], perm[:nnz]

    def csr_device(self, transpose: bool = False) -> dict:
        """The device-resident CSR arrays of A (or A') as zero-copy torch views (valid while the solver lives):
        ptr (int32 view of the 32-bit unsigned pointers), idx (int32), val (float64), perm (int32 view of uint32)."""
        import torch
        p = [C.c_void_p() for _ in range(4)]
        _lib.check(_lib.load().lsqr_b200_ez_get_csr_device(self._h, int(transpose), *[C.byref(q) for q in p]))
        nkeys = (self.n if transpose else self.m) * self.blocks(transpose)[0]
        nnz = self.nnz

        class _View:   # minimal __cuda_array_interface__ holder
            def __init__(self, ptr, count, typestr):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}

        def view(q, count, typestr):
            return torch.as_tensor(_View(q.value, count, typestr), device="cuda")
        return {"ptr": view(p[0], nkeys + 1, "<i4"), "idx": view(p[1], max(nnz, 1), "<i4")[:nnz],
                "val": view(p[2], max(nnz, 1), "<f8")[:nnz], "perm": view(p[3], max(nnz, 1), "<i4")[:nnz]}

    def kernel_times(self) -> dict:
        t = KernelTimes()
        _lib.check(_lib.load().lsqr_b200_ez_get_kernel_times(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in KernelTimes._fields_}

    @property
    def handle(self) -> int:
        return self._h.value or 0

    def destroy(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            _lib.load().lsqr_b200_ez_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------
# abstract lsqr_solver (operator hook)
# --------------------------------------------------------------------------------------------
class LsqrSolver:
    """``type,abstract :: lsqr_solver`` (src/lsqr.f90:16-30).  Extend it and override ``aprod``:

        def aprod(self, mode, m, n, x_ptr, y_ptr, stream) -> None
            # mode 1: y(m) += A x(n) ; mode 2: x(n) += A' y(m)
            # x_ptr / y_ptr are DEVICE pointers (ints); enqueue on `stream`, do not synchronize.
    """

    def aprod(self, mode: int, m: int, n: int, x_ptr: int, y_ptr: int, stream: int) -> None:   # deferred (:26)
        raise NotImplementedError("aprod is deferred: extend LsqrSolver and provide it")

    def _trampoline(self):
        def thunk(_user, mode, m, n, xp, yp, stream):
            try:
                self.aprod(mode, m, n, xp or 0, yp or 0, stream or 0)
                return 0
            except Exception as e:   # surfaced as LSQR_B200_ERR_CALLBACK
                self._cb_error = e
                return 1
        return _lib.APROD_FN(thunk)

    def _options(self, nout, trace, device, stream):
        o = _lib.default_options()
        o.device, o.stream = int(device), int(stream) or None
        cb = _Callbacks(nout=nout, trace=trace)
        cb.install(o)
        return o, cb

    # LSQR, src/lsqr.f90:432-882.  u,v,w,x,se are DEVICE arrays (tensors or raw pointers).
    def lsqr(self, m, n, damp, wantse, u, v, w, x, se, atol, btol, conlim, itnlim, nout=None,
             trace: bool = False, device: int = -1, stream: int = 0) -> SolveResult:
        L = _lib.load()
        tr = [] if trace else None
        o, cb = self._options(nout, tr, device, stream)
        fn = self._trampoline()
        istop, itn = C.c_int32(), C.c_int32()
        sc = [C.c_double() for _ in range(5)]
        self._cb_error = None
        rc = L.lsqr_b200_lsqr(fn, None, int(m), int(n), float(damp), int(bool(wantse)),
                              _ptr(u), _ptr(v), _ptr(w), _ptr(x), _ptr(se) if wantse else None,
                              float(atol), float(btol), float(conlim), int(itnlim), C.byref(o),
                              C.byref(istop), C.byref(itn), *[C.byref(s) for s in sc])
        if rc == 15 and self._cb_error is not None:
            raise self._cb_error
        _lib.check(rc)
        return SolveResult(x, istop.value, itn.value, *[s.value for s in sc], se=se if wantse else None,
                           log=cb.lines, trace=tr or [])

    # acheck, src/lsqr.f90:908-994
    def acheck(self, m, n, v, w, x, y, eps=float(np.finfo(np.float64).eps), nout=None, device=-1, stream=0):
        o, cb = self._options(nout, None, device, stream)
        fn = self._trampoline()
        inform, rel = C.c_int32(), C.c_double()
        self._cb_error = None
        rc = _lib.load().lsqr_b200_acheck(fn, None, int(m), int(n), float(eps), _ptr(v), _ptr(w), _ptr(x), _ptr(y),
                                          C.byref(o), C.byref(inform), C.byref(rel))
        if rc == 15 and self._cb_error is not None:
            raise self._cb_error
        _lib.check(rc)
        return inform.value, rel.value

    # xcheck, src/lsqr.f90:1015-1154
    def xcheck(self, m, n, anorm, damp, b, u, v, w, x, eps=float(np.finfo(np.float64).eps), nout=None,
               device=-1, stream=0) -> dict:
        o, cb = self._options(nout, None, device, stream)
        fn = self._trampoline()
        inform = C.c_int32()
        t = [C.c_double() for _ in range(3)]
        norms = (C.c_double * 6)()
        self._cb_error = None
        rc = _lib.load().lsqr_b200_xcheck(fn, None, int(m), int(n), float(anorm), float(damp), float(eps),
                                          _ptr(b), _ptr(u), _ptr(v), _ptr(w), _ptr(x), C.byref(o),
                                          C.byref(inform), *[C.byref(s) for s in t], norms)
        if rc == 15 and self._cb_error is not None:
            raise self._cb_error
        _lib.check(rc)
        keys = ("bnorm", "xnorm", "rho1", "sigma1", "rho2", "sigma2")
        out = dict(inform=inform.value, test1=t[0].value, test2=t[1].value, test3=t[2].value, log=cb.lines)
        out.update({k: norms[i] for i, k in enumerate(keys)})
        return out


class LsqrSolverHost(LsqrSolver):
    """The abstract ``lsqr_solver`` with the reference's OWN signatures (src/lsqr.f90:16-30,67-82): every vector is a
    host (numpy) array and the operator is host code.  Extend it and override

        def aprod(self, mode, m, n, x, y) -> None        # mode 1: y += A x ; mode 2: x += A' y   (numpy views, in place)

    The products make a host round trip per call; all vector arithmetic and the scalar recurrence run on the GPU.  This
    is what an unmodified ``type,extends(lsqr_solver)`` of the reference binds to; a device operator (``LsqrSolver``)
    is the fast path."""

    def aprod(self, mode: int, m: int, n: int, x: np.ndarray, y: np.ndarray) -> None:   # deferred (:26)
        raise NotImplementedError("aprod is deferred: extend LsqrSolverHost and provide it")

    def _trampoline(self):
        def thunk(_user, mode, m, n, xp, yp):
            try:
                x = np.ctypeslib.as_array(xp, shape=(max(n, 1),))[:n]
                y = np.ctypeslib.as_array(yp, shape=(max(m, 1),))[:m]
                self.aprod(mode, m, n, x, y)
                return 0
            except Exception as e:   # surfaced as LSQR_B200_ERR_CALLBACK
                self._cb_error = e
                return 1
        return _lib.APROD_HOST_FN(thunk)

    # LSQR, src/lsqr.f90:432-882, reference argument list: u (holds b, overwritten), v, w, x, se are host arrays
    def lsqr(self, m, n, damp, wantse, u, v, w, x, se, atol, btol, conlim, itnlim, nout=None,
             trace: bool = False, device: int = -1, stream: int = 0) -> SolveResult:
        L = _lib.load()
        tr = [] if trace else None
        o, cb = self._options(nout, tr, device, stream)
        fn = self._trampoline()
        istop, itn = C.c_int32(), C.c_int32()
        sc = [C.c_double() for _ in range(5)]
        self._cb_error = None
        rc = L.lsqr_b200_lsqr_host(fn, None, int(m), int(n), float(damp), int(bool(wantse)),
                                   _ptr(u), _ptr(v), _ptr(w), _ptr(x), _ptr(se) if wantse else None,
                                   float(atol), float(btol), float(conlim), int(itnlim), C.byref(o),
                                   C.byref(istop), C.byref(itn), *[C.byref(s) for s in sc])
        if rc == 15 and self._cb_error is not None:
            raise self._cb_error
        _lib.check(rc)
        return SolveResult(x, istop.value, itn.value, *[s.value for s in sc], se=se if wantse else None,
                           log=cb.lines, trace=tr or [])

    def acheck(self, m, n, v, w, x, y, eps=float(np.finfo(np.float64).eps), nout=None, device=-1, stream=0):
        o, cb = self._options(nout, None, device, stream)
        fn = self._trampoline()
        inform, rel = C.c_int32(), C.c_double()
        self._cb_error = None
        rc = _lib.load().lsqr_b200_acheck_host(fn, None, int(m), int(n), float(eps), _ptr(v), _ptr(w), _ptr(x), _ptr(y),
                                               C.byref(o), C.byref(inform), C.byref(rel))
        if rc == 15 and self._cb_error is not None:
            raise self._cb_error
        _lib.check(rc)
        return inform.value, rel.value

    def xcheck(self, m, n, anorm, damp, b, u, v, w, x, eps=float(np.finfo(np.float64).eps), nout=None,
               device=-1, stream=0) -> dict:
        o, cb = self._options(nout, None, device, stream)
        fn = self._trampoline()
        inform = C.c_int32()
        t = [C.c_double() for _ in range(3)]
        norms = (C.c_double * 6)()
        self._cb_error = None
        rc = _lib.load().lsqr_b200_xcheck_host(fn, None, int(m), int(n), float(anorm), float(damp), float(eps),
                                               _ptr(b), _ptr(u), _ptr(v), _ptr(w), _ptr(x), C.byref(o),
                                               C.byref(inform), *[C.byref(s) for s in t], norms)
        if rc == 15 and self._cb_error is not None:
            raise self._cb_error
        _lib.check(rc)
        keys = ("bnorm", "xnorm", "rho1", "sigma1", "rho2", "sigma2")
        out = dict(inform=inform.value, test1=t[0].value, test2=t[1].value, test3=t[2].value, log=cb.lines)
        out.update({k: norms[i] for i, k in enumerate(keys)})
        return out


class EzAsOperator(LsqrSolver):
    """Drives an initialized ``LsqrSolverEz`` through the low-level path, the way
    ``class(lsqr_solver_ez)`` *is a* ``lsqr_solver`` in the reference."""

    def __init__(self, ez: LsqrSolverEz):
        self.ez = ez

    def aprod(self, mode, m, n, x_ptr, y_ptr, stream):
        _lib.check(_lib.load().lsqr_b200_ez_aprod_device(self.ez.handle, mode, m, n, x_ptr, y_ptr, stream))


# --------------------------------------------------------------------------------------------
# device BLAS-1 (src/lsqrblas.f90) on device arrays
# --------------------------------------------------------------------------------------------
def dnrm2(n, x, stream: int = 0) -> float:
    r = C.c_double()
    _lib.check(_lib.load().lsqr_b200_dnrm2(int(n), _ptr(x), C.byref(r), int(stream) or None))
    return r.value


def ddot(n, x, y, stream: int = 0) -> float:
    r = C.c_double()
    _lib.check(_lib.load().lsqr_b200_ddot(int(n), _ptr(x), _ptr(y), C.byref(r), int(stream) or None))
    return r.value


def dscal(n, da, x, stream: int = 0) -> None:
    _lib.check(_lib.load().lsqr_b200_dscal(int(n), float(da), _ptr(x), int(stream) or None))


def dcopy(n, x, y, stream: int = 0) -> None:
    _lib.check(_lib.load().lsqr_b200_dcopy(int(n), _ptr(x), _ptr(y), int(stream) or None))
