"""Device-side synthetic workloads (include/lsqr_b200_synth.h, csrc/synth.cu): the same triplets and
vectors as synth.py, generated straight into HBM as torch CUDA tensors (torch is only the allocator)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, synth


def _bind():
    L = _lib.load()
    if getattr(L, "_synth_bound", False):
        return L
    vp = C.c_void_p
    L.lsqr_b200_synth_row_ptr.restype = C.c_int
    L.lsqr_b200_synth_row_ptr.argtypes = [C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_int32, vp, C.c_int32,
                                          vp, C.POINTER(C.c_int64), vp]
    L.lsqr_b200_synth_fill.restype = C.c_int
    L.lsqr_b200_synth_fill.argtypes = [C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       vp, C.c_int64, vp, vp, vp, vp]
    L.lsqr_b200_synth_vector.restype = C.c_int
    L.lsqr_b200_synth_vector.argtypes = [C.c_uint64, C.c_int32, C.c_double, C.c_int64, C.c_int64, vp, vp]
    L._synth_bound = True
    return L


def coo_block(kind: str, seed: int, m: int, n: int, k: int, row0: int, nrows: int, device):
    """(irow, icol, a) of global rows [row0, row0+nrows) as CUDA tensors; bit-identical to synth.coo_block."""
    import torch
    L = _bind()
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        ptr = torch.empty(nrows + 1, dtype=torch.int64, device=device)
        table = synth.powerlaw_table() if kind == "powerlaw" else None
        nnz = C.c_int64(0)
        _lib.check(L.lsqr_b200_synth_row_ptr(synth.KINDS[kind], seed, row0, nrows, k,
                                             table.ctypes.data if table is not None else None,
                                             table.size if table is not None else 0,
                                             ptr.data_ptr(), C.byref(nnz), stream))
        nz = int(nnz.value)
        irow = torch.empty(nz, dtype=torch.int32, device=device)
        icol = torch.empty(nz, dtype=torch.int32, device=device)
        a = torch.empty(nz, dtype=torch.float64, device=device)
        _lib.check(L.lsqr_b200_synth_fill(synth.KINDS[kind], seed, m, n, row0, nrows, ptr.data_ptr(), nz,
                                          irow.data_ptr(), icol.data_ptr(), a.data_ptr(), stream))
    return irow, icol, a


def vector(seed: int, tag: int, scale: float, offset: int, count: int, device):
    import torch
    L = _bind()
    with torch.cuda.device(device):
        out = torch.empty(count, dtype=torch.float64, device=device)
        _lib.check(L.lsqr_b200_synth_vector(seed, tag, float(np.float64(scale) * np.float64(synth.SQRT3)), offset, count,
                                            out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


def x_true(seed: int, n: int, device):
    return vector(seed, synth.TAG_XTRUE, 1.0, 0, n, device)


def noise(seed: int, row0: int, nrows: int, device, scale: float = 1e-3):
    return vector(seed, synth.TAG_NOISE, scale, row0, nrows, device)
