"""Multi-GPU host logic: A is row-partitioned, one process per GPU.

Every rank owns a contiguous block of rows of A (and of b / u); v, w, x are replicated.  Aprod is
local; the per-rank partial A_p'u_p (n entries) plus the partial sum(u_p^2) are combined by ONE
NCCL all-reduce of n+1 doubles per iteration inside the engine (csrc/engine.cu).  This module holds
what the launcher side needs: the row partition, the bootstrap of the engine's NCCL communicator
over torch.distributed (any backend: nccl on GPUs, gloo in the CPU tests), and block generation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, synth


def row_block(m: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced row range [row0, row1) of `rank`; the first m % world ranks get one more row."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(m, world)
    row0 = rank * base + min(rank, extra)
    return row0, row0 + base + (1 if rank < extra else 0)


def row_blocks_by_nnz(row_nnz_prefix: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Row ranges with (nearly) equal numbers of stored entries, for skewed row lengths (C4).
    `row_nnz_prefix` is the CSR row pointer (length m+1)."""
    m = row_nnz_prefix.size - 1
    total = int(row_nnz_prefix[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(row_nnz_prefix, total * r / world, side="left")))
    cuts.append(m)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, m))
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(world)]


def new_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().lsqr_b200_nccl_unique_id(buf))
    return buf.raw


def exchange_unique_id(world: int, rank: int, make_id=new_unique_id) -> bytes:
    """Rank 0 creates the engine's ncclUniqueId; torch.distributed broadcasts the 128 bytes."""
    import torch
    import torch.distributed as td
    on_gpu = td.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone().to(dev)
    td.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def generate_block(cfg: dict, row0: int, nrows: int, device=None):
    """COO triplets of one row block: on the device when the CUDA generator is built, else numpy."""
    if device is not None:
        try:
            from . import synth_device
            return synth_device.coo_block(cfg["kind"], cfg["seed"], cfg["m"], cfg["n"], cfg["k"], row0, nrows, device)
        except ImportError:
            pass
    return synth.coo_block(cfg["kind"], cfg["seed"], cfg["m"], cfg["n"], cfg["k"], row0, nrows)
