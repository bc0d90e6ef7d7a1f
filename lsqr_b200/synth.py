"""Deterministic synthetic workloads of BASELINE.json, host (numpy) side.

Every entry is a pure function of (seed, global row, slot) through a splitmix64-style counter hash,
so any row block can be generated independently (multi-GPU row partitioning) and the CUDA generator
(csrc/synth.cu, lsqr_b200_synth_*) produces bit-identical triplets.

    uniform   (C2, C5): k entries per row, col ~ U{1..n},                           val ~ U(-1,1)
    banded    (C3)    : k entries per row, col = (floor(row*n/m) + U{-100..100}) mod n + 1, val ~ U(0,1)
    powerlaw  (C4)    : L = #{j in 1..10000 : U <= j^-0.75} entries, col ~ U{1..n},  val ~ U(-1,1)/sqrt(L)

Rows are emitted in order (row-sorted COO, 1-based indices); `shuffle_coo` gives the globally
shuffled variant that exercises the device sort.
"""
from __future__ import annotations

import numpy as np

KINDS = {"uniform": 0, "banded": 1, "powerlaw": 2}
TAG_COL, TAG_VAL, TAG_LEN, TAG_XTRUE, TAG_NOISE = 1, 2, 3, 4, 5
POWERLAW_MAX = 10000
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_ROWMUL = np.uint64(0xD1342543DE82EF95)
SQRT3 = 1.7320508075688772


def _mix(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def hash3(seed: int, tag: int, r, k):
    """h(seed, tag, r, k) = mix(mix(mix(seed + tag*GOLD) + r*ROWMUL) + k)  (uint64 wrap-around)."""
    with np.errstate(over="ignore"):
        base = _mix(np.uint64(seed) + np.uint64(tag) * _GOLD)
        return _mix(_mix(base + np.asarray(r, dtype=np.uint64) * _ROWMUL) + np.asarray(k, dtype=np.uint64))


def u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def powerlaw_table() -> np.ndarray:
    """T[j-1] = j^-0.75, j = 1..10000.  The SAME array is handed to the device generator."""
    return np.arange(1, POWERLAW_MAX + 1, dtype=np.float64) ** -0.75


def row_lengths(kind: str, seed: int, row0: int, nrows: int, k: int, table=None) -> np.ndarray:
    if kind != "powerlaw":
        return np.full(nrows, k, dtype=np.int64)
    table = powerlaw_table() if table is None else table
    u = u01(hash3(seed, TAG_LEN, np.arange(row0, row0 + nrows, dtype=np.uint64), 0))
    # number of j with u <= T[j]; T is decreasing, so search in the negated (increasing) table
    return np.searchsorted(-table, -u, side="right").astype(np.int64)


def coo_block(kind: str, seed: int, m: int, n: int, k: int, row0: int = 0, nrows: int | None = None, table=None):
    """COO triplets (irow, icol, a) of global rows [row0, row0+nrows); irow is 1-based INSIDE the block."""
    nrows = m - row0 if nrows is None else nrows
    lens = row_lengths(kind, seed, row0, nrows, k, table)
    ptr = np.concatenate([[0], np.cumsum(lens)])
    nnz = int(ptr[-1])
    local_row = np.repeat(np.arange(nrows, dtype=np.int64), lens)
    slot = np.arange(nnz, dtype=np.int64) - ptr[local_row]
    grow = (local_row + row0).astype(np.uint64)
    hc = hash3(seed, TAG_COL, grow, slot)
    hv = hash3(seed, TAG_VAL, grow, slot)
    if kind == "banded":
        center = (grow * np.uint64(n)) // np.uint64(m)
        off = ((hc >> np.uint64(32)) * np.uint64(201)) >> np.uint64(32)          # 0..200
        col = (center + np.uint64(n) + off - np.uint64(100)) % np.uint64(n)
        val = u01(hv)
    else:
        col = ((hc >> np.uint64(32)) * np.uint64(n)) >> np.uint64(32)             # 0..n-1
        val = 2.0 * u01(hv) - 1.0
        if kind == "powerlaw":
            val = val / np.sqrt(lens.astype(np.float64))[local_row]
    return (local_row + 1).astype(np.int32), (col + np.uint64(1)).astype(np.int32), val


def vector(seed: int, tag: int, scale: float, offset: int, count: int) -> np.ndarray:
    """scale * sqrt(3) * (2u-1): zero mean, variance scale^2."""
    u = u01(hash3(seed, tag, np.arange(offset, offset + count, dtype=np.uint64), 0))
    return (scale * SQRT3) * (2.0 * u - 1.0)


def x_true(seed: int, n: int) -> np.ndarray:
    return vector(seed, TAG_XTRUE, 1.0, 0, n)


def noise(seed: int, row0: int, nrows: int, scale: float = 1e-3) -> np.ndarray:
    return vector(seed, TAG_NOISE, scale, row0, nrows)


def rhs_block(irow, icol, a, nrows: int, xt: np.ndarray, seed: int, row0: int = 0) -> np.ndarray:
    """b = A x_true + 1e-3 * noise for one row block (host SpMV with numpy)."""
    b = np.zeros(nrows)
    np.add.at(b, irow.astype(np.int64) - 1, a * xt[icol.astype(np.int64) - 1])
    return b + noise(seed, row0, nrows)


def shuffle_coo(irow, icol, a, seed: int):
    """Globally shuffled triplets (a deterministic permutation): exercises the device key sort."""
    p = np.argsort(hash3(seed, 99, np.arange(irow.size, dtype=np.uint64), 0), kind="stable")
    return irow[p], icol[p], a[p]


# The BASELINE.json configurations.  `k` is entries per row (ignored for powerlaw).
CONFIGS = {
    "C2": dict(kind="uniform", m=1_000_000, n=100_000, k=10, damp=0.0, seed=1),
    "C3": dict(kind="banded", m=10_000_000, n=2_000_000, k=50, damp=1e-3, seed=2),
    "C4": dict(kind="powerlaw", m=20_000_000, n=5_000_000, k=0, damp=0.0, seed=3),
    "C5": dict(kind="uniform", m=100_000_000, n=10_000_000, k=20, damp=0.0, seed=4),
}


def scaled(name: str, scale: float) -> dict:
    """A configuration with m and n divided by `scale` (same kind, k, damp, seed)."""
    c = dict(CONFIGS[name])
    c["m"] = max(1, int(round(c["m"] / scale)))
    c["n"] = max(1, int(round(c["n"] / scale)))
    return c


def b_iter_bytes(nnz: int, m: int, n: int, wantse: bool = False) -> int:
    """Algorithmic bytes per LSQR iteration: 24 nnz + 28 m + 68 n + 8 (+16 n with se)."""
    return 24 * nnz + 28 * m + 68 * n + 8 + (16 * n if wantse else 0)
