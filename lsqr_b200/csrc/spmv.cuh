// spmv.cuh -- K3/K4: the warp-autonomous segmented CSR SpMV of the engine, hand-written for sm_100a.
//
// ONE persistent launch computes one whole product, A v or A'u, even when the matrix is stored blocked along the
// gathered coordinate (column blocks of A / row blocks of A', so that the gathered slice of the dense vector stays
// in L2).  The rows are cut into row-aligned tiles, the SAME row ranges in every block; a warp owns the same tiles
// in every block and walks the blocks in order, so "block b+1 accumulates onto what block b stored" is program
// order inside one warp and needs neither atomics nor a grid barrier.  A soft guard keeps the warps of a launch
// within two blocks of each other, so the gathered working set stays at two L2-sized slices at most.
//
// A tile of one block (a "piece") is consumed in chunks of 128 stored entries, 4 per lane:
//   * val[] / idx[] of chunk c+1 are fetched with one 256-bit and one 128-bit streaming load per lane (evict-first)
//     while chunk c is reduced (register double buffer);
//   * x[idx]: if the piece's gathered indices span a window that fits the warp's shared-memory buffer (banded /
//     local matrices; the span is recorded per piece at initialize), the window is staged into shared memory with
//     coalesced loads and the gathers are shared-memory reads (conflict-bound, ~5x the L1TEX rate of a divergent
//     global gather); otherwise 4 independent global gathers per lane through L1/L2 (evict-last);
//   * row heads inside the chunk come from a 32-row window of ptr[] held in registers (lane j <-> row wb+j), turned
//     into a 128-bit head mask with warp-wide OR reductions;
//   * a segmented scan (4 entries in the lane, then 5 shuffle steps over lanes) forms the running row sums; a row
//     that spans chunks is carried in a register, so rows of ANY length need no special path;
//   * the running sums go through a 1 KB per-warp shared buffer so that lane j picks up the sum of row wb+j: the
//     epilogue is a coalesced read-modify-write of consecutive rows.
// Epilogue of a block: STORE (part = s), ACC (part += s) or FINAL, the fused LSQR step
//   u' = ca_mat (part + s) + ca_vec u, sum u'^2   /   v' = ct_mat (part + s) + ct_vec v, sum v'^2
// followed by the scalar recurrence in the block that finishes last (steps.cuh), or -- multi-GPU peer path -- the
// push of (part + s) into the owner rank's receive buffer over NVLink.
// The order of every floating-point addition is fixed by the tile map, never by scheduling: results are
// run-to-run reproducible.  Sums of squares use Blue's scaled accumulators (common.cuh).
#pragma once

#include "peer.cuh"

namespace lsqrb {

constexpr int kWThreads = 256;               // 8 warps per CTA
constexpr int kWWarps = kWThreads / 32;
constexpr uint32_t kChunk = 128;             // stored entries per warp step (4 per lane)
constexpr uint32_t kPtrSentinel = 0xFFFFFFFFu;
constexpr uint32_t kNoTile = 0xFFFFFFFFu;
constexpr int kMaxSpmvBlocks = 64;           // DevState::blk_done

struct CsrView {
    const uint32_t *ptr;   // [nrows+1]
    const int32_t  *idx;   // [nnz] 0-based other coordinate
    const double   *val;   // [nnz]
    int64_t         nrows;
};

// One piece of work: rows [row, next.row) of one block, stored entries [entry, next.entry).  win_len != 0: every
// gathered index of the piece lies in [win_lo, win_lo + win_len) and win_len fits the launch's shared window.
struct __align__(16) TileDesc {
    uint32_t row, entry, win_lo, win_len;
};

enum FinalKind {
    FIN_NONE = 0,         // plain product: every block stores / accumulates (y += A x, g = A'u)
    FIN_APROD = 1,        // u' = ca_mat (part + s) + ca_vec u ; sum u'^2 ; step_after_aprod (or the partial to aux)
    FIN_ATPROD = 2,       // v' = ct_mat (part + s) + ct_vec v ; sum v'^2 ; step_after_atprod
    FIN_INIT_ATPROD = 3,  // v  = ct_mat (part + s)            ; sum v^2  ; step_init_alpha
    FIN_PUSH = 4          // multi-GPU peer path: (part + s) of column j goes to the receive buffer of j's owner rank
};
enum BlockMode { BM_STORE = 0, BM_ACC = 1, BM_FINAL = 2 };

struct SpmvArgs {
    const int32_t *idx;        // stored entries of all blocks, back to back
    const double *val;
    const uint32_t *ptr;       // pointer array of the first block of this launch; block b at ptr + b * ptr_stride
    int64_t ptr_stride;
    const TileDesc *tiles;     // pieces of the first block of this launch: [nblocks][ntiles + 1]
    int ntiles, nblocks;
    const uint32_t *order;     // balanced schedule (slot k * warps + w = k-th tile of warp w) or nullptr = round robin
    int nslots;
    int64_t nrows;
    const double *x;           // gathered dense vector
    double *out;               // rows of the result (FINAL epilogue)
    double *part;              // partial sums across blocks (STORE / ACC epilogues)
    int first_mode;            // BM_STORE or BM_ACC: what the first block of this launch does with `part`
    int last_is_final;         // the last block of this launch applies the FINAL epilogue (FIN != FIN_NONE)
    int final_adds_part;       // FINAL adds part[row] (the matrix has more than one block)
    int win_cap;               // doubles of shared-memory gather window per warp (0 = none)
    int guard;                 // 0: none; g > 0: block b starts when every warp has finished block b - g (1 = one live slice)
    int check_done;            // plain product launched from the solve loop: nothing to do once the solver has stopped
    int probe;                 // residency probe: arrive, wait for the whole grid (or 20 ms), leave
    DevState *st;
    Ssq *aux;                  // FIN_APROD, multi-GPU: where the local sum u'^2 goes instead of step_after_aprod
    // FIN_PUSH: push[q] = where this rank's contributions to the columns owned by rank q go (q's receive buffer,
    // already offset to this rank's slot); columns [q * push_cols, (q+1) * push_cols) belong to rank q
    double *const *push;
    int64_t push_cols;
    const PeerView *peer;      // FIN_PUSH: the exchange block (device copy), for the flags that follow the pushes
};

__device__ __forceinline__ void ldg_stream_s32x4(const int32_t *p, int32_t (&v)[4], uint64_t pol)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p), "l"(pol));
}

// Everything a warp needs to know about the block it is working on (warp-uniform).
struct BlockCtx {
    const uint32_t *ptr;
    int mode;                  // BlockMode
    bool adds_part;
};

// Row window: lane j holds row wb+j of the piece [.., r1): P = ptr[row], PE = ptr[row+1], O = the value the epilogue
// combines with (part[row] for ACC, out[row] for FINAL), G = part[row] for a FINAL that adds the earlier blocks.
// Rows at or beyond r1 carry the sentinel so they neither start nor end anywhere.
template <int FIN>
struct RowWindow {
    uint32_t P, PE;
    double O, G;
    __device__ __forceinline__ void load(const SpmvArgs &a, const BlockCtx &bc, uint32_t wb, uint32_t r1, int lane)
    {
        const uint32_t r = wb + (uint32_t)lane;
        P = PE = kPtrSentinel;
        O = 0.0;
        G = 0.0;
        if (r < r1) {
            P = bc.ptr[r];
            PE = bc.ptr[r + 1];
            if (bc.mode == BM_ACC) O = a.part[r];
            if (FIN != FIN_NONE && bc.mode == BM_FINAL) {
                if (FIN == FIN_APROD || FIN == FIN_ATPROD) O = a.out[r];
                if (bc.adds_part) G = a.part[r];
            }
        }
    }
};

template <int FIN>
struct Epilogue {
    double sq = 0.0;           // mid-range sum of squares of this thread (FINAL)
    double *exc;               // this thread's exceptional accumulators (shared memory)
    __device__ __forceinline__ double coef_mat(const DevState *st) const
    {
        return FIN == FIN_APROD ? __ldg(&st->ca_mat) : __ldg(&st->ct_mat);
    }
    __device__ __forceinline__ double coef_vec(const DevState *st) const
    {
        return FIN == FIN_APROD ? __ldg(&st->ca_vec) : __ldg(&st->ct_vec);
    }
    // s = the row's sum over this block's entries
    __device__ __forceinline__ void apply(const SpmvArgs &a, const BlockCtx &bc, uint32_t row, double s, double old, double g)
    {
        if (bc.mode == BM_STORE) { a.part[row] = s; return; }
        if (bc.mode == BM_ACC) { a.part[row] = old + s; return; }
        if (FIN == FIN_NONE) return;
        s += g;
        if (FIN == FIN_PUSH) {
            const int64_t q = (int64_t)row / a.push_cols;
            a.push[q][(int64_t)row - q * a.push_cols] = s;
            return;
        }
        // the coefficients are re-read from the (kernel-invariant, L1-resident) device state: the kernel is
        // register-bound and two doubles live across the whole tile loop would spill
        double r = coef_mat(a.st) * s;
        if (FIN != FIN_INIT_ATPROD) r += coef_vec(a.st) * old;
        a.out[row] = r;
        ssq_add(sq, exc, kWThreads, r);
    }
};

// One lane's share of a chunk: 4 consecutive stored entries.
struct ChunkRegs {
    double v[4];
    int32_t c[4];
};

__device__ __forceinline__ int32_t ldg_stream_s32_1(const int32_t *p, uint64_t pol)
{
    int32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}

// Loads one lane's share of the chunk that starts at entry `cb` of the piece [e0, e1).  Values: entries
// [cb+4 lane, +4) (blocked: what the segmented scan wants).  Indices: the same entries, or -- STRIPED -- entries
// cb + 32 k + lane, k = 0..3, so that one gather instruction covers 32 CONSECUTIVE stored entries: in a matrix whose
// rows hold clustered / sorted indices (banded A', FEM, tomography) those fall into a handful of 128-byte lines
// instead of 32, and the L1TEX tag stage -- one line per clock -- stops being the limiter.
// Entries outside the piece (they belong to the neighbouring pieces or to the allocation slack) get the value 0 and
// the index `safe` (inside the piece's gather window), so they add nothing to any row and gather harmlessly.
template <bool STRIPED>
__device__ __forceinline__ void load_chunk(const SpmvArgs &a, uint32_t cb, int lane, uint32_t e0, uint32_t e1,
                                           int32_t safe, uint64_t pol_stream, ChunkRegs &r)
{
    const uint32_t q = cb + 4u * (uint32_t)lane;
    if (cb >= e0 && cb + kChunk <= e1) {   // steady state: the whole chunk lies inside the piece
        ldg_stream_f64x4(a.val + q, r.v);
        if (STRIPED) {
#pragma unroll
            for (int k = 0; k < 4; ++k) r.c[k] = ldg_stream_s32_1(a.idx + cb + 32u * k + (uint32_t)lane, pol_stream);
        } else {
            ldg_stream_s32x4(a.idx + q, r.c, pol_stream);
        }
        return;
    }
    if (q < e1) {
        ldg_stream_f64x4(a.val + q, r.v);
        if (!STRIPED) ldg_stream_s32x4(a.idx + q, r.c, pol_stream);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q + k < e0 || q + k >= e1) { r.v[k] = 0.0; if (!STRIPED) r.c[k] = safe; }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { r.v[k] = 0.0; if (!STRIPED) r.c[k] = safe; }
    }
    if (STRIPED) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t e = cb + 32u * k + (uint32_t)lane;
            r.c[k] = (e >= e0 && e < e1) ? ldg_stream_s32_1(a.idx + e, pol_stream) : safe;
        }
    }
}

template <int FIN>
struct WarpTileState {
    uint32_t wb, woff;     // window base row, lanes below woff are finished rows
    RowWindow<FIN> win;
    double carry;          // running sum of the row that is open at the chunk boundary
};

// Reduces the chunk [base, base+128): `cv` = the lane's 4 stored values, x0..x3 = the gathered vector entries.
template <int FIN>
__device__ __forceinline__ void warp_chunk_core(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, double *su,
                                                WarpTileState<FIN> &ts, const double (&cv)[4],
                                                double x0, double x1, double x2, double x3,
                                                uint32_t base, uint32_t r1, uint32_t e1, int lane)
{
    const uint32_t endp = base + kChunk;
    // ---- head mask of the chunk: bit i set <=> a row starts at entry base+i
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    {
        uint32_t p = ts.win.P;
        uint32_t wb2 = ts.wb;
        for (;;) {
            const uint32_t rel = p - base;
            const uint32_t bit = rel < kChunk ? (1u << (rel & 31u)) : 0u;
            const uint32_t w = rel >> 5;
            m0 |= __reduce_or_sync(0xffffffffu, w == 0u ? bit : 0u);
            m1 |= __reduce_or_sync(0xffffffffu, w == 1u ? bit : 0u);
            m2 |= __reduce_or_sync(0xffffffffu, w == 2u ? bit : 0u);
            m3 |= __reduce_or_sync(0xffffffffu, w == 3u ? bit : 0u);
            const uint32_t p31 = __shfl_sync(0xffffffffu, p, 31);
            if (p31 >= endp || wb2 + 32u >= r1) break;     // (the sentinel ends the walk too)
            wb2 += 32u;                                    // more than a window of rows starts in this chunk
            const uint32_t r = wb2 + (uint32_t)lane;
            p = r < r1 ? bc.ptr[r] : kPtrSentinel;
        }
    }
    const uint32_t mw = (lane < 8) ? m0 : (lane < 16) ? m1 : (lane < 24) ? m2 : m3;
    const uint32_t f = (mw >> ((lane & 7) * 4)) & 0xFu;

    // ---- products; segmented running sums inside the lane (t_k = sum of the lane's entries of the segment
    // that entry k belongs to, up to and including k)
    double t0 = cv[0] * x0;
    if (lane == 0 && !(f & 1u)) t0 = ts.carry + t0;        // row that began in an earlier chunk
    double t1 = cv[1] * x1;
    if (!(f & 2u)) t1 += t0;
    double t2 = cv[2] * x2;
    if (!(f & 4u)) t2 += t1;
    double t3 = cv[3] * x3;
    if (!(f & 8u)) t3 += t2;
    // ---- ... and across lanes (Kogge-Stone; lane l takes lane l-d iff no head lies in lanes (l-d, l])
    const uint32_t hb = __ballot_sync(0xffffffffu, f != 0u);
    const uint32_t below = hb & (0xffffffffu >> (31 - lane));
    const int reach = lane - (below ? 31 - __clz(below) : 0);   // how far down this lane's open segment extends
    double vs = t3;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, vs, d);
        if (d <= reach) vs += y;
    }
    double cin = __shfl_up_sync(0xffffffffu, vs, 1);
    if (lane == 0) cin = 0.0;
    // entries that precede the lane's first head continue the segment of the lanes below
    if (!(f & 1u)) t0 += cin;
    if (!(f & 3u)) t1 += cin;
    if (!(f & 7u)) t2 += cin;
    if (!(f & 15u)) t3 += cin;
    ts.carry = __shfl_sync(0xffffffffu, t3, 31);
    *reinterpret_cast<double2 *>(su + 4 * lane) = make_double2(t0, t1);
    *reinterpret_cast<double2 *>(su + 4 * lane + 2) = make_double2(t2, t3);
    __syncwarp();

    // ---- rows that end in this chunk: lane j finishes row wb+j (the sentinel of absent rows never ends)
    for (;;) {
        const bool ends = (uint32_t)lane >= ts.woff && ts.win.PE <= endp;
        if (ends) {
            const double s = (ts.win.PE != ts.win.P) ? su[ts.win.PE - 1u - base] : 0.0;
            epi.apply(a, bc, ts.wb + (uint32_t)lane, s, ts.win.O, ts.win.G);
        }
        ts.woff += (uint32_t)__popc(__ballot_sync(0xffffffffu, ends));
        if (ts.woff >= 32u && ts.wb + 32u < r1) {          // window exhausted: more rows may end here
            ts.wb += 32u;
            ts.woff = 0;
            ts.win.load(a, bc, ts.wb, r1, lane);
            continue;
        }
        break;
    }
    __syncwarp();
    // keep the window ahead of the stream: the reload is in flight during the next chunk
    if (endp < e1 && ts.woff >= 16u && ts.wb + ts.woff < r1) {
        ts.wb += ts.woff;
        ts.woff = 0;
        ts.win.load(a, bc, ts.wb, r1, lane);
    }
}

// Processes the chunk [base, base+128) held in `cur`; `nxt` receives the following chunk, whose loads stay in
// flight while this one is reduced.  WIN: the gathers are reads of the warp's staged window.  STRIPED: the gathered
// values arrive in striped order (entry 32 k + lane) and are transposed to the blocked order of the values (entries
// 4 lane .. 4 lane + 3) through the warp's shared buffer.
template <int FIN, bool WIN, bool STRIPED>
__device__ __forceinline__ void warp_chunk(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, double *su,
                                           const double *xw /* WIN: window - win_lo */, WarpTileState<FIN> &ts,
                                           const ChunkRegs &cur, ChunkRegs &nxt, uint32_t base, uint32_t r1,
                                           uint32_t e0, uint32_t e1, int32_t safe, int lane,
                                           uint64_t pol_stream, uint64_t pol_keep)
{
    double x0, x1, x2, x3;
    if (WIN) {
        x0 = xw[cur.c[0]]; x1 = xw[cur.c[1]]; x2 = xw[cur.c[2]]; x3 = xw[cur.c[3]];
    } else {
        x0 = ldg_keep_f64(a.x + cur.c[0], pol_keep);
        x1 = ldg_keep_f64(a.x + cur.c[1], pol_keep);
        x2 = ldg_keep_f64(a.x + cur.c[2], pol_keep);
        x3 = ldg_keep_f64(a.x + cur.c[3], pol_keep);
    }
    load_chunk<STRIPED>(a, base + kChunk, lane, e0, e1, safe, pol_stream, nxt);
    if (STRIPED) {
        // (su is free here: the previous chunk's row sums were consumed before its closing __syncwarp)
        su[lane] = x0; su[32 + lane] = x1; su[64 + lane] = x2; su[96 + lane] = x3;
        __syncwarp();
        const double2 p = *reinterpret_cast<const double2 *>(su + 4 * lane);
        const double2 q = *reinterpret_cast<const double2 *>(su + 4 * lane + 2);
        x0 = p.x; x1 = p.y; x2 = q.x; x3 = q.y;
        __syncwarp();                   // every lane has its values before the scan overwrites su
    }
    warp_chunk_core<FIN>(a, bc, epi, su, ts, cur.v, x0, x1, x2, x3, base, r1, e1, lane);
}

template <int FIN, bool WIN, bool STRIPED>
__device__ __forceinline__ void warp_tile_loop(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, double *su,
                                               const double *xw, uint32_t r0, uint32_t r1, uint32_t e0, uint32_t e1,
                                               int32_t safe, int lane, uint64_t pol_stream, uint64_t pol_keep)
{
    const uint32_t a0 = e0 & ~3u;
    WarpTileState<FIN> ts;
    ts.wb = r0;
    ts.woff = 0;
    ts.carry = 0.0;
    ts.win.load(a, bc, r0, r1, lane);
    ChunkRegs ra, rb;
    load_chunk<STRIPED>(a, a0, lane, e0, e1, safe, pol_stream, ra);
    // two chunks per trip so that the register double buffer needs no copies; at least one chunk is processed
    // even for a piece without entries, so that its (empty) rows still get their epilogue
    for (uint32_t base = a0;;) {
        warp_chunk<FIN, WIN, STRIPED>(a, bc, epi, su, xw, ts, ra, rb, base, r1, e0, e1, safe, lane, pol_stream, pol_keep);
        base += kChunk;
        if (base >= e1) break;
        warp_chunk<FIN, WIN, STRIPED>(a, bc, epi, su, xw, ts, rb, ra, base, r1, e0, e1, safe, lane, pol_stream, pol_keep);
        base += kChunk;
        if (base >= e1) break;
    }
}

// One piece: stage the gather window if the piece has one, then stream its chunks.
template <int FIN, bool STRIPED>
__device__ __forceinline__ void warp_tile(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, double *su, double *wbuf,
                                          const TileDesc &d0, const TileDesc &d1, int lane,
                                          uint64_t pol_stream, uint64_t pol_keep)
{
    if (d0.win_len != 0u) {
        // coalesced 8-byte loads: no alignment requirement on x, which may be the caller's own array (aprod)
        const double *src = a.x + d0.win_lo;
        for (uint32_t i = (uint32_t)lane; i < d0.win_len; i += 32u) wbuf[i] = ldg_keep_f64(src + i, pol_keep);
        __syncwarp();
        // (a staged window is conflict-bound, not line-bound: the blocked index layout with its single 128-bit load)
        warp_tile_loop<FIN, true, false>(a, bc, epi, su, wbuf - d0.win_lo, d0.row, d1.row, d0.entry, d1.entry,
                                         (int32_t)d0.win_lo, lane, pol_stream, pol_keep);
    } else {
        warp_tile_loop<FIN, false, STRIPED>(a, bc, epi, su, nullptr, d0.row, d1.row, d0.entry, d1.entry, 0, lane, pol_stream, pol_keep);
    }
}

// MINB: resident CTAs per SM the kernel is compiled for (4: <= 64 registers, 32 warps per SM; 2: <= 128 registers,
// 16 warps per SM with room for wide gather windows).  STRIPED: lane-consecutive gathers (see load_chunk).
template <int FIN, int MINB, bool STRIPED>
__global__ void __launch_bounds__(kWThreads, MINB)
spmv_kernel(SpmvArgs a)
{
    constexpr bool kFused = (FIN == FIN_APROD || FIN == FIN_ATPROD || FIN == FIN_INIT_ATPROD);
    extern __shared__ __align__(16) double s_dyn[];          // per warp: su[128] | gather window[win_cap]
    __shared__ double s_red[kWWarps];
    __shared__ double s_exc[2 * kWThreads];

    DevState *st = a.st;
    const int tid = threadIdx.x;
    bool tracing = false;
    if (a.probe) {
        // Residency probe (initialize): do ALL CTAs of this grid run at the same time, with this kernel's registers and
        // this launch's shared memory?  The drift guard of a multi-block launch is a grid barrier and relies on it.
        if (tid == 0) {
            atomicAdd(&st->probe_count, 1u);
            const unsigned long long t0 = globaltimer_ns();
            while (*(volatile unsigned int *)&st->probe_count < gridDim.x) {
                if (globaltimer_ns() - t0 > 20000000ull) { st->probe_fail = 1; break; }
                __nanosleep(200);
            }
        }
        return;
    }
    if (kFused) {
        if (st->done) return;
        if (FIN == FIN_APROD && st->istop != 0) return;   // the stop is decided: only the x/w update of that iteration is left
        if (FIN == FIN_ATPROD && st->beta == 0.0) {
            // beta = 0: the reference skips the A' half and keeps alpha (src/lsqr.f90:691-699)
            if (blockIdx.x == 0 && tid == 0) step_after_atprod(*st, 0.0, false);
            return;
        }
        tracing = st->tr_on != 0;
        if (tracing && blockIdx.x == 0 && tid == 0) st->trace[0][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
    } else if (a.check_done && st->done) {
        return;
    }
    s_exc[tid] = 0.0;
    s_exc[kWThreads + tid] = 0.0;
    Epilogue<FIN> epi;
    epi.exc = s_exc + tid;

    {
        const uint64_t pol_stream = l2_policy_evict_first();
        const uint64_t pol_keep = l2_policy_evict_last();
        const int lane = tid & 31, wib = tid >> 5;
        const int nw = (int)gridDim.x * kWWarps;
        double *su = s_dyn + (size_t)wib * (kChunk + (size_t)a.win_cap);
        double *wbuf = su + kChunk;
        const uint32_t *__restrict__ order = a.order;
        const int nslots = order ? a.nslots : a.ntiles;
        for (int b = 0; b < a.nblocks; ++b) {
            BlockCtx bc;
            bc.ptr = a.ptr + (int64_t)b * a.ptr_stride;
            bc.mode = (b == a.nblocks - 1 && a.last_is_final) ? BM_FINAL : (b == 0 ? a.first_mode : BM_ACC);
            bc.adds_part = a.final_adds_part != 0;
            if (a.guard && b >= a.guard) {
                // guard = 1: a block starts when EVERY CTA has finished the previous one, so exactly one gathered slice
                // is live in L2 (measured: a warp that runs ahead touches the whole next slice within microseconds --
                // random gathers -- and two 48 MB slices do not fit; C5/4 Atprod 7.3 ms vs 5.4 ms).  guard = 2: one
                // block of slack.  One thread per CTA polls (4736 polling warps stole L1TEX cycles from the stragglers
                // they were waiting for: 6 % on C5/4); the launch is cooperative, so the grid is co-resident.
                if (tid == 0) {
                    const volatile unsigned int *done = &st->blk_done[b - a.guard];
                    const unsigned long long t0 = globaltimer_ns();
                    while (*done < gridDim.x) {
                        __nanosleep(400);
                        // (belt and braces: should a CTA still wait for seconds, latch an error the host reports
                        // instead of hanging the GPU)
                        if (*(volatile int *)&st->guard_error) break;
                        if (globaltimer_ns() - t0 > 4000000000ull) { st->guard_error = 1; break; }
                    }
                }
                __syncthreads();
            }
            const TileDesc *__restrict__ tiles = a.tiles + (size_t)b * ((size_t)a.ntiles + 1);
            for (int s = (int)blockIdx.x * kWWarps + wib; s < nslots; s += nw) {
                const uint32_t t = order ? order[s] : (uint32_t)s;
                if (t == kNoTile) break;                        // this warp's list is exhausted
                const TileDesc d0 = tiles[t], d1 = tiles[t + 1];
                if (d0.row == d1.row) continue;                 // no row starts in this tile (inside a long row)
                warp_tile<FIN, STRIPED>(a, bc, epi, su, wbuf, d0, d1, lane, pol_stream, pol_keep);
            }
            if (a.guard && b + 1 < a.nblocks) {
                __syncthreads();                                // every warp of this CTA has finished block b
                if (tid == 0) atomicAdd(&st->blk_done[b], 1u);
            }
        }
    }

    if (kFused) {
        Ssq total;
        if (finish_ssq<kWThreads>(st, 0, st->partial, epi.sq, s_exc, s_red, &total)) {
            if (a.guard) for (int b = 0; b + 1 < a.nblocks; ++b) st->blk_done[b] = 0;
            if (tracing) st->trace[1][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
            if (FIN == FIN_APROD) {
                if (a.aux) *a.aux = total; else step_after_aprod(*st, ssq_norm(total));
            }
            else if (FIN == FIN_ATPROD) step_after_atprod(*st, ssq_norm(total), true);
            else step_init_alpha(*st, ssq_norm(total));
            if (tracing) { st->trace[2][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns(); st->tr_n += 1; }
        }
    } else if (FIN == FIN_PUSH) {
        // this block's peer stores are fenced system-wide before its ticket; the last block publishes the flags
        __syncthreads();
        if (tid == 0) __threadfence_system();
        if (last_block_ticket(st, 0)) {
            if (a.guard) for (int b = 0; b + 1 < a.nblocks; ++b) st->blk_done[b] = 0;
            peer_publish_partials(st, *a.peer);
        }
    } else if (a.guard && a.nblocks > 1) {
        if (last_block_ticket(st, 0)) for (int b = 0; b + 1 < a.nblocks; ++b) st->blk_done[b] = 0;
    }
}

// =============================================================================================
// Tile map construction (initialize): the same row cuts for every block of one matrix.
// =============================================================================================
// Work coordinate of row r over ALL blocks: W(r) = sum_b (ptr_b[r] - ptr_b[0]) + row_w * nblocks * r  (stored
// entries before the row, plus row_w units per row and block: a run of EMPTY rows is work too -- every row gets its
// epilogue -- and must not land in one tile).  Tile t starts at the first row with W(r) >= t * tile.
__global__ void build_tiles_kernel(const uint32_t *__restrict__ ptr, int64_t nrows, int nblocks, int ntiles,
                                   uint64_t tile, uint32_t row_w, TileDesc *tiles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    int64_t lo = 0, hi = nrows;   // first r in [0, nrows] with W(r) >= target
    if (t == ntiles) {
        lo = nrows;
    } else {
        const uint64_t target = (uint64_t)t * tile;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            uint64_t w = (uint64_t)row_w * (uint64_t)nblocks * (uint64_t)mid;
            for (int b = 0; b < nblocks; ++b) w += (uint64_t)ptr[(int64_t)b * nrows + mid] - (uint64_t)ptr[(int64_t)b * nrows];
            if (w >= target) hi = mid; else lo = mid + 1;
        }
    }
    for (int b = 0; b < nblocks; ++b) {
        TileDesc d;
        d.row = (uint32_t)lo;
        d.entry = ptr[(int64_t)b * nrows + lo];
        d.win_lo = 0;
        d.win_len = 0;
        tiles[(size_t)b * ((size_t)ntiles + 1) + (size_t)t] = d;
    }
}

// Gather span of every piece: one warp per piece scans its indices (min / max).  On the way it measures how local 32
// CONSECUTIVE stored entries are: stat[0] += 128-byte lines a gather of such a group can touch at most
// (span of the group / 16 + 1, capped at 32), stat[1] += groups.  Integer atomics: order-independent.
__global__ void __launch_bounds__(256)
tile_span_kernel(const int32_t *__restrict__ idx, TileDesc *tiles, int64_t npieces_with_sentinels, int ntiles,
                 unsigned long long *stat)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long lines = 0, groups = 0;
    for (int64_t p = w; p < npieces_with_sentinels; p += nwarps) {
        if ((p % ((int64_t)ntiles + 1)) == ntiles) continue;      // the sentinel entry that closes a block's list
        const uint32_t e0 = tiles[p].entry, e1 = tiles[p + 1].entry;
        int32_t lo = INT32_MAX, hi = INT32_MIN;
        for (uint32_t eb = e0; eb < e1; eb += 32u) {
            const uint32_t e = eb + (uint32_t)lane;
            const bool ok = e < e1;
            const int32_t c = ok ? idx[e] : 0;
            const int32_t glo = __reduce_min_sync(0xffffffffu, ok ? c : INT32_MAX);
            const int32_t ghi = __reduce_max_sync(0xffffffffu, ok ? c : INT32_MIN);
            lo = min(lo, glo);
            hi = max(hi, ghi);
            lines += (unsigned long long)min(32, (ghi - glo) / 16 + 1);
            groups += 1;
        }
        if (lane == 0) {
            tiles[p].win_lo = e1 > e0 ? (uint32_t)lo : 0u;
            tiles[p].win_len = e1 > e0 ? (uint32_t)(hi - lo + 1) : 0u;   // span; the host clears it where it exceeds the window
        }
    }
    if (lane == 0 && groups) { atomicAdd(stat, lines); atomicAdd(stat + 1, groups); }
}

// win_len > cap (or an empty piece) -> 0: that piece gathers from global memory
__global__ void tile_window_cap_kernel(TileDesc *tiles, int64_t n, uint32_t cap)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && tiles[i].win_len > cap) { tiles[i].win_len = 0; tiles[i].win_lo = 0; }
}

}  // namespace lsqrb
