// spmv_warp.cuh -- K3/K4, variant 3: warp-autonomous segmented CSR SpMV for sm_100a.
//
// The matrix is cut into row-aligned tiles; every warp of a persistent grid walks its own tiles
// without any CTA-wide barrier.  A tile is consumed in chunks of 128 stored entries (4 per lane):
//   * val[] / idx[] of chunk c+1 are fetched with one 256-bit and one 128-bit streaming load per
//     lane (1.5 KB per warp, evict-first) while chunk c is being processed (register double
//     buffering keeps ~48 KB per SM in flight towards HBM);
//   * x[idx] is gathered through L1/L2 (evict-last), 4 independent loads per lane;
//   * row heads inside the chunk come from a 32-row window of ptr[] held in registers (lane j <->
//     row wb+j) and are turned into a 128-bit head mask with warp-wide OR reductions;
//   * a segmented scan (4 entries in the lane, then 5 shuffle steps over lanes) forms the running
//     row sums; a row that spans chunks is carried in a register, so rows of ANY length need no
//     special path;
//   * the running sums go through a 1 KB per-warp shared buffer so that lane j picks up the sum of
//     row wb+j: the fused epilogue (out[] old value prefetched with the window) is then a coalesced
//     read-modify-write of consecutive rows.
// The order of every floating-point addition is fixed by the tile map (chunk boundaries are relative
// to the tile start), never by scheduling: results are run-to-run reproducible.
// HBM traffic = val + idx once, ptr once, out once (+old once).
#pragma once

#include "spmv_stream.cuh"

namespace lsqrb {

constexpr int kWThreads = 256;               // 8 warps per CTA
constexpr int kWWarps = kWThreads / 32;
#ifndef LSQRB_WARP_MINBLOCKS
#define LSQRB_WARP_MINBLOCKS 4
#endif
constexpr int kWMinBlocks = LSQRB_WARP_MINBLOCKS;   // 4: 32 warps per SM, <= 64 registers per thread
constexpr uint32_t kChunk = 128;             // stored entries per warp step (4 per lane)
#ifndef LSQRB_L2_PREFETCH_CHUNKS
#define LSQRB_L2_PREFETCH_CHUNKS 0           // D > 0: one lane bulk-prefetches val/idx of the chunk D steps past the register
#endif                                       // double buffer into L2 (more bytes in flight towards HBM at no register cost)
#ifndef LSQRB_WIN_PREFETCH
#define LSQRB_WIN_PREFETCH 0                 // 1: prefetch the ptr / out lines of the next row window into L2 (measured: no gain)
#endif
#ifndef LSQRB_GATHER_AHEAD
#define LSQRB_GATHER_AHEAD 0                 // 1: gathers issued one chunk ahead of their use (measured 15-25 % SLOWER: the kernel is
                                             // bound by L1TEX wavefronts, not by gather latency, and the extra registers spill)
#endif
constexpr uint32_t kPtrSentinel = 0xFFFFFFFFu;

__device__ __forceinline__ void ldg_stream_s32x4(const int32_t *p, int32_t (&v)[4], uint64_t pol)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p), "l"(pol));
}

// Row window: lane j holds row wb+j of the tile [.., r1): P = ptr[row], PE = ptr[row+1], O = out[row].
// Rows at or beyond r1 carry the sentinel so they neither start nor end anywhere.
template <int EPI>
struct RowWindow {
    uint32_t P, PE;
    double O;
    double W, X;   // fused deferred update only: w[row], x[row]
    __device__ __forceinline__ void load(const StreamArgs &a, const RowEpilogue<EPI, true> &epi, uint32_t wb, uint32_t r1, int lane)
    {
        const uint32_t r = wb + (uint32_t)lane;
        P = PE = kPtrSentinel;
        O = 0.0;
        if (EPI == SEPI_ATPROD_UPD) W = X = 0.0;
        if (EPI == SEPI_APROD_ACC) W = 0.0;
        if (r < r1) {
            P = a.A.ptr[r];
            PE = a.A.ptr[r + 1];
            if (epi.needs_old()) O = a.out[r];
            if (EPI == SEPI_ATPROD_UPD && epi.upd) { W = a.uw[r]; X = a.ux[r]; }
            if (EPI == SEPI_APROD_ACC) W = a.uw[r];
        }
#if LSQRB_WIN_PREFETCH
        // the next window is needed a chunk or two from now and its reload sits on the critical path of the head
        // mask: pull its lines (ptr: one line per 32 rows, out: two) towards the SM now; no registers are held
        if (r + 32u < r1) {
            if ((lane & 15) == 0) prefetch_l2(a.A.ptr + r + 32u + (lane ? 16u : 0u));
            if (epi.needs_old() && (lane & 7) == 0) prefetch_l2(a.out + r + 32u);
        }
#endif
    }
};

// One lane's share of a chunk: 4 consecutive stored entries.
struct ChunkRegs {
    double v[4];
    int32_t c[4];
};

// Loads one lane's share of the chunk that starts at entry `cb` of the tile [e0, e1): entries [cb+4 lane, +4).
// Entries outside the tile (they belong to the neighbouring tiles or to the allocation slack) get the value 0
// and so add nothing to any row; their column index stays a valid one.  Lanes entirely past the tile read
// nothing (index 0: a harmless in-bounds gather).
__device__ __forceinline__ void load_chunk(const CsrView &A, uint32_t cb, int lane, uint32_t e0, uint32_t e1,
                                           uint64_t pol_stream, ChunkRegs &r)
{
    const uint32_t q = cb + 4u * (uint32_t)lane;
    if (cb >= e0 && cb + kChunk <= e1) {   // steady state: the whole chunk lies inside the tile
        ldg_stream_f64x4(A.val + q, r.v);
        ldg_stream_s32x4(A.idx + q, r.c, pol_stream);
        return;
    }
    if (q < e1) {
        ldg_stream_f64x4(A.val + q, r.v);
        ldg_stream_s32x4(A.idx + q, r.c, pol_stream);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q + k < e0 || q + k >= e1) r.v[k] = 0.0;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { r.v[k] = 0.0; r.c[k] = 0; }
    }
}

// L2 prefetch of the chunk that starts at entry cb (whole chunks inside the tile only)
__device__ __forceinline__ void prefetch_chunk(const CsrView &A, uint32_t cb, uint32_t e1, int lane)
{
#if LSQRB_L2_PREFETCH_CHUNKS > 0
    if (lane == 0 && cb + kChunk <= e1) {
        prefetch_l2_bulk(A.val + cb, kChunk * 8u);
        prefetch_l2_bulk(A.idx + cb, kChunk * 4u);
    }
#endif
}

template <int EPI>
struct WarpTileState {
    uint32_t wb, woff;     // window base row, lanes below woff are finished rows
    RowWindow<EPI> win;
    double carry;          // running sum of the row that is open at the chunk boundary
};

// Reduces the chunk [base, base+128): `cv` = the lane's 4 stored values, x0..x3 = the gathered vector entries.
template <int EPI>
__device__ __forceinline__ void warp_chunk_core(const StreamArgs &a, RowEpilogue<EPI, true> &epi, double *su, WarpTileState<EPI> &ts,
                                                const double (&cv)[4], double x0, double x1, double x2, double x3,
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
            p = r < r1 ? a.A.ptr[r] : kPtrSentinel;
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
            if (EPI == SEPI_ATPROD_UPD) epi.apply(a, (int64_t)ts.wb + lane, s, ts.win.O, true, ts.win.W, ts.win.X);
            else if (EPI == SEPI_APROD_ACC) epi.apply(a, (int64_t)ts.wb + lane, s, ts.win.O, true, ts.win.W);
            else epi.apply(a, (int64_t)ts.wb + lane, s, ts.win.O);
        }
        ts.woff += (uint32_t)__popc(__ballot_sync(0xffffffffu, ends));
        if (ts.woff >= 32u && ts.wb + 32u < r1) {          // window exhausted: more rows may end here
            ts.wb += 32u;
            ts.woff = 0;
            ts.win.load(a, epi, ts.wb, r1, lane);
            continue;
        }
        break;
    }
    __syncwarp();
    // keep the window ahead of the stream: the reload is in flight during the next chunk
    if (endp < e1 && ts.woff >= 16u && ts.wb + ts.woff < r1) {
        ts.wb += ts.woff;
        ts.woff = 0;
        ts.win.load(a, epi, ts.wb, r1, lane);
    }
}

// Processes the chunk [base, base+128) held in `cur`; `nxt` receives the following chunk, whose loads
// stay in flight while this one is reduced.  (LSQRB_GATHER_AHEAD = 0: gathers are issued and consumed in the
// same step.)
template <int EPI>
__device__ __forceinline__ void warp_chunk(const StreamArgs &a, RowEpilogue<EPI, true> &epi, double *su, WarpTileState<EPI> &ts,
                                           const ChunkRegs &cur, ChunkRegs &nxt, uint32_t base, uint32_t r1, uint32_t e0, uint32_t e1,
                                           int lane, uint64_t pol_stream, uint64_t pol_keep)
{
    // ---- gathers of this chunk, then the stream of the next one
    const double x0 = ldg_keep_f64(a.x + cur.c[0], pol_keep);
    const double x1 = ldg_keep_f64(a.x + cur.c[1], pol_keep);
    const double x2 = ldg_keep_f64(a.x + cur.c[2], pol_keep);
    const double x3 = ldg_keep_f64(a.x + cur.c[3], pol_keep);
    load_chunk(a.A, base + kChunk, lane, e0, e1, pol_stream, nxt);
    prefetch_chunk(a.A, base + (1u + LSQRB_L2_PREFETCH_CHUNKS) * kChunk, e1, lane);
    warp_chunk_core<EPI>(a, epi, su, ts, cur.v, x0, x1, x2, x3, base, r1, e1, lane);
}

#if LSQRB_GATHER_AHEAD
// Gather-ahead pipeline: three chunks are in flight per warp.  While chunk c is reduced, the gathers x[idx] of
// chunk c+1 (issued at the start of the step, from indices that arrived during the previous step), the values of
// chunk c+1 and the indices of chunk c+2 are outstanding, so a warp never waits for the L2 round trip of a gather
// it has just issued.  The index registers are reused as soon as the gathers that read them have been issued.
struct ChunkStage {
    double v[4];
    double x[4];
};

__device__ __forceinline__ void load_chunk_idx(const CsrView &A, uint32_t cb, int lane, uint32_t e1, uint64_t pol_stream, int32_t (&c)[4])
{
    const uint32_t q = cb + 4u * (uint32_t)lane;
    if (q < e1) {
        ldg_stream_s32x4(A.idx + q, c, pol_stream);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = 0;          // a harmless in-bounds gather
    }
}

// values outside the tile [e0, e1) become 0 and so add nothing to any row
__device__ __forceinline__ void load_chunk_val(const CsrView &A, uint32_t cb, int lane, uint32_t e0, uint32_t e1, double (&v)[4])
{
    const uint32_t q = cb + 4u * (uint32_t)lane;
    if (cb >= e0 && cb + kChunk <= e1) {               // steady state: the whole chunk lies inside the tile
        ldg_stream_f64x4(A.val + q, v);
        return;
    }
    if (q < e1) {
        ldg_stream_f64x4(A.val + q, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q + k < e0 || q + k >= e1) v[k] = 0.0;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = 0.0;
    }
}

// issue the loads of the two following chunks: gathers + values of [nb, nb+128), indices of [nb+128, nb+256)
__device__ __forceinline__ void stage_issue(const StreamArgs &a, ChunkStage &s, int32_t (&c)[4], uint32_t nb, int lane,
                                            uint32_t e0, uint32_t e1, uint64_t pol_stream, uint64_t pol_keep)
{
    if (nb >= e1) return;                              // (warp-uniform) no following chunk
    s.x[0] = ldg_keep_f64(a.x + c[0], pol_keep);
    s.x[1] = ldg_keep_f64(a.x + c[1], pol_keep);
    s.x[2] = ldg_keep_f64(a.x + c[2], pol_keep);
    s.x[3] = ldg_keep_f64(a.x + c[3], pol_keep);
    load_chunk_val(a.A, nb, lane, e0, e1, s.v);
    load_chunk_idx(a.A, nb + kChunk, lane, e1, pol_stream, c);
}

template <int EPI>
__device__ __forceinline__ void warp_tile(const StreamArgs &a, RowEpilogue<EPI, true> &epi, double *su /* 128 doubles, this warp's */,
                                          uint32_t r0, uint32_t r1, uint32_t e0, uint32_t e1, int lane,
                                          uint64_t pol_stream, uint64_t pol_keep)
{
    const uint32_t a0 = e0 & ~3u;
    int32_t c[4];
    ChunkStage sa, sb;
    load_chunk_idx(a.A, a0, lane, e1, pol_stream, c);
    load_chunk_val(a.A, a0, lane, e0, e1, sa.v);
    WarpTileState<EPI> ts;
    ts.wb = r0;
    ts.woff = 0;
    ts.carry = 0.0;
    ts.win.load(a, epi, r0, r1, lane);
    sa.x[0] = ldg_keep_f64(a.x + c[0], pol_keep);
    sa.x[1] = ldg_keep_f64(a.x + c[1], pol_keep);
    sa.x[2] = ldg_keep_f64(a.x + c[2], pol_keep);
    sa.x[3] = ldg_keep_f64(a.x + c[3], pol_keep);
    load_chunk_idx(a.A, a0 + kChunk, lane, e1, pol_stream, c);
    // two chunks per trip so that the stage registers need no copies; at least one chunk is processed even for a
    // tile without entries, so that its (empty) rows still get their epilogue
    for (uint32_t base = a0;;) {
        stage_issue(a, sb, c, base + kChunk, lane, e0, e1, pol_stream, pol_keep);
        warp_chunk_core<EPI>(a, epi, su, ts, sa.v, sa.x[0], sa.x[1], sa.x[2], sa.x[3], base, r1, e1, lane);
        base += kChunk;
        if (base >= e1) break;
        stage_issue(a, sa, c, base + kChunk, lane, e0, e1, pol_stream, pol_keep);
        warp_chunk_core<EPI>(a, epi, su, ts, sb.v, sb.x[0], sb.x[1], sb.x[2], sb.x[3], base, r1, e1, lane);
        base += kChunk;
        if (base >= e1) break;
    }
}
#else
template <int EPI>
__device__ __forceinline__ void warp_tile(const StreamArgs &a, RowEpilogue<EPI, true> &epi, double *su /* 128 doubles, this warp's */,
                                          uint32_t r0, uint32_t r1, uint32_t e0, uint32_t e1, int lane,
                                          uint64_t pol_stream, uint64_t pol_keep)
{
    const uint32_t a0 = e0 & ~3u;
    WarpTileState<EPI> ts;
    ts.wb = r0;
    ts.woff = 0;
    ts.carry = 0.0;
    ts.win.load(a, epi, r0, r1, lane);
    ChunkRegs ra, rb;
    load_chunk(a.A, a0, lane, e0, e1, pol_stream, ra);
#pragma unroll
    for (uint32_t d = 1; d <= LSQRB_L2_PREFETCH_CHUNKS; ++d) prefetch_chunk(a.A, a0 + d * kChunk, e1, lane);
    // two chunks per trip so that the register double buffer needs no copies; at least one chunk is processed
    // even for a tile without entries, so that its (empty) rows still get their epilogue
    for (uint32_t base = a0;;) {
        warp_chunk<EPI>(a, epi, su, ts, ra, rb, base, r1, e0, e1, lane, pol_stream, pol_keep);
        base += kChunk;
        if (base >= e1) break;
        warp_chunk<EPI>(a, epi, su, ts, rb, ra, base, r1, e0, e1, lane, pol_stream, pol_keep);
        base += kChunk;
        if (base >= e1) break;
    }
}

#endif

template <int EPI>
__global__ void __launch_bounds__(kWThreads, kWMinBlocks)
spmv_warp_kernel(StreamArgs a)
{
    constexpr bool kFused = (sepi_is_aprod(EPI) || EPI == SEPI_ATPROD || EPI == SEPI_INIT_ATPROD || EPI == SEPI_ATPROD_UPD);
    __shared__ double s_red[kWWarps];
    __shared__ __align__(16) double s_u[kWWarps][kChunk];

    DevState *st = a.st;
    RowEpilogue<EPI, true> epi;
    int mode = MODE_FULL;
    const int tid = threadIdx.x;
    bool tracing = false;
    if (kFused) {
        if (st->done) return;
        if (sepi_is_aprod(EPI) && st->istop != 0) return;   // stop already decided: only the deferred update is left
        if (EPI == SEPI_ATPROD && st->beta == 0.0) {
            // beta = 0: the reference skips the A' half and keeps alpha (src/lsqr.f90:691-699)
            if (blockIdx.x == 0 && threadIdx.x == 0) step_after_atprod(*st, 0.0, false);
            return;
        }
        if (EPI == SEPI_ATPROD_UPD) {
            if (st->istop != 0) mode = MODE_FLUSH;
            else if (st->beta == 0.0) mode = MODE_UPDATE_ONLY;
        }
        epi.load(st);
        tracing = st->tr_on != 0;
        if (tracing && blockIdx.x == 0 && tid == 0) st->trace[0][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
    } else if (a.check_done && st->done) {
        return;
    }

    if (mode != MODE_FULL) {
        // elementwise part only (n-vectors): x += t1 w [; w' = v/alpha + t2 w]
        if (epi.upd) {
            epi.load_update_coefficients(st);
            const int64_t n = a.A.nrows;
            for (int64_t i = (int64_t)blockIdx.x * kWThreads + tid; i < n; i += (int64_t)gridDim.x * kWThreads) {
                const double wo = a.uw[i];
                a.ux[i] = epi.t1 * wo + a.ux[i];
                if (epi.wantse) a.use[i] += (epi.t3 * wo) * (epi.t3 * wo);
                if (mode == MODE_UPDATE_ONLY) {
                    const double wn = epi.t2 * wo + epi.ia * a.out[i];
                    a.uw[i] = wn;
                    epi.sq2 += wn * wn;
                }
            }
        }
    } else {
        const uint64_t pol_stream = l2_policy_evict_first();
        const uint64_t pol_keep = l2_policy_evict_last();
        const int lane = tid & 31, wib = tid >> 5;
        const int nw = (int)gridDim.x * kWWarps;
        const uint2 *__restrict__ tiles = a.map.tiles;
        const uint32_t *__restrict__ order = a.map.order;
        const int nslots = order ? a.map.nslots : a.map.ntiles;
        for (int s = (int)blockIdx.x * kWWarps + wib; s < nslots; s += nw) {
            const uint32_t t = order ? order[s] : (uint32_t)s;
            if (t == kNoTile) break;                        // this warp's list is exhausted
            const uint2 d0 = tiles[t], d1 = tiles[t + 1];
            if (d0.x == d1.x) continue;                     // no row starts in this tile (inside a long row)
            warp_tile<EPI>(a, epi, s_u[wib], d0.x, d1.x, d0.y, d1.y, lane, pol_stream, pol_keep);
        }
    }

    if (kFused) {
        double total, total_w;
        if (EPI == SEPI_ATPROD_UPD) {
            // sum(w'^2) of the deferred update closes iteration k; sum(v'^2) drives the step of iteration k+1
            if (finish_reduction2<kWThreads>(st, 0, epi.sq, epi.sq2, s_red, &total, &total_w)) {
                if (tracing) st->trace[1][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
                if (st->upd_pending) {
                    __threadfence();
                    step_after_update(*st, total_w, __ldcg(a.ux), a.ring);
                    st->upd_pending = 0;
                }
                if (tracing) st->trace[3][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
                if (!st->done) {
                    step_after_atprod(*st, total, mode == MODE_FULL);
                    st->upd_pending = 1;
                }
                if (tracing) { st->trace[2][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns(); st->tr_n += 1; }
            }
        } else if (finish_reduction<kWThreads>(st, 0, epi.sq, s_red, &total)) {
            if (tracing) st->trace[1][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
            if (sepi_is_aprod(EPI)) {
                if (a.aux) *a.aux = total; else step_after_aprod(*st, total);
            }
            else if (EPI == SEPI_ATPROD) step_after_atprod(*st, total, true);
            else step_init_alpha(*st, total);
            if (tracing) { st->trace[2][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns(); st->tr_n += 1; }
        }
    }
}

}  // namespace lsqrb
