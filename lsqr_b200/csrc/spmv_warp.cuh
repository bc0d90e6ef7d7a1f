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
constexpr int kWMinBlocks = 4;               // 32 warps per SM, <= 64 registers per thread
constexpr uint32_t kChunk = 128;             // stored entries per warp step (4 per lane)
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
    __device__ __forceinline__ void load(const StreamArgs &a, const RowEpilogue<EPI> &epi, uint32_t wb, uint32_t r1, int lane)
    {
        const uint32_t r = wb + (uint32_t)lane;
        P = PE = kPtrSentinel;
        O = 0.0;
        if (r < r1) {
            P = a.A.ptr[r];
            PE = a.A.ptr[r + 1];
            if (epi.needs_old()) O = a.out[r];
        }
    }
};

template <int EPI>
__device__ __forceinline__ void warp_tile(const StreamArgs &a, RowEpilogue<EPI> &epi, double *su /* 128 doubles, this warp's */,
                                          uint32_t r0, uint32_t r1, uint32_t e0, uint32_t e1, int lane,
                                          uint64_t pol_stream, uint64_t pol_keep)
{
    const uint32_t a0 = e0 & ~3u;
    const uint32_t *__restrict__ ptr = a.A.ptr;
    uint32_t wb = r0, woff = 0;
    RowWindow<EPI> win;
    win.load(a, epi, wb, r1, lane);

    double v[4] = {0.0, 0.0, 0.0, 0.0};
    int32_t c[4] = {0, 0, 0, 0};
    {
        const uint32_t q = a0 + 4u * (uint32_t)lane;
        if (q < e1) {
            ldg_stream_f64x4(a.A.val + q, v);
            ldg_stream_s32x4(a.A.idx + q, c, pol_stream);
        }
    }
    double carry = 0.0;

    // (at least one pass, so that a tile of empty rows still gets its epilogue)
    for (uint32_t base = a0;; base += kChunk) {
        const uint32_t endp = base + kChunk;
        const uint32_t q = base + 4u * (uint32_t)lane;

        // ---- gathers of this chunk (invalid lanes hold index 0: a harmless in-bounds read)
        const double x0 = ldg_keep_f64(a.x + c[0], pol_keep);
        const double x1 = ldg_keep_f64(a.x + c[1], pol_keep);
        const double x2 = ldg_keep_f64(a.x + c[2], pol_keep);
        const double x3 = ldg_keep_f64(a.x + c[3], pol_keep);
        const double cv0 = v[0], cv1 = v[1], cv2 = v[2], cv3 = v[3];

        // ---- stream of the next chunk, in flight while this one is reduced
        {
            const uint32_t qn = q + kChunk;
            v[0] = v[1] = v[2] = v[3] = 0.0;
            c[0] = c[1] = c[2] = c[3] = 0;
            if (qn < e1) {
                ldg_stream_f64x4(a.A.val + qn, v);
                ldg_stream_s32x4(a.A.idx + qn, c, pol_stream);
            }
        }

        // ---- head mask of the chunk: bit i set <=> a row starts at entry base+i
        uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        {
            uint32_t p = win.P;
            uint32_t wb2 = wb;
            for (;;) {
                const uint32_t rel = p - base;
                const bool hv = rel < kChunk;
                const uint32_t bit = hv ? (1u << (rel & 31u)) : 0u;
                const uint32_t w = rel >> 5;
                m0 |= __reduce_or_sync(0xffffffffu, w == 0u ? bit : 0u);
                m1 |= __reduce_or_sync(0xffffffffu, w == 1u ? bit : 0u);
                m2 |= __reduce_or_sync(0xffffffffu, w == 2u ? bit : 0u);
                m3 |= __reduce_or_sync(0xffffffffu, w == 3u ? bit : 0u);
                const uint32_t p31 = __shfl_sync(0xffffffffu, p, 31);
                if (p31 >= endp || wb2 + 32u >= r1) break;     // (the sentinel ends the walk too)
                wb2 += 32u;                                    // more than a window of rows starts in this chunk
                const uint32_t r = wb2 + (uint32_t)lane;
                p = r < r1 ? ptr[r] : kPtrSentinel;
            }
        }
        const uint32_t mw = (lane < 8) ? m0 : (lane < 16) ? m1 : (lane < 24) ? m2 : m3;
        const uint32_t f = (mw >> ((lane & 7) * 4)) & 0xFu;

        // ---- products (entries outside [e0, e1) belong to other tiles), carry of a row that began earlier
        double p0 = (q + 0u >= e0 && q + 0u < e1) ? cv0 * x0 : 0.0;
        const double p1 = (q + 1u >= e0 && q + 1u < e1) ? cv1 * x1 : 0.0;
        const double p2 = (q + 2u >= e0 && q + 2u < e1) ? cv2 * x2 : 0.0;
        const double p3 = (q + 3u >= e0 && q + 3u < e1) ? cv3 * x3 : 0.0;
        if (lane == 0 && !(f & 1u)) p0 = carry + p0;

        // ---- segmented running sums inside the lane ...
        const double t0 = p0;
        const double t1 = (f & 2u) ? p1 : t0 + p1;
        const double t2 = (f & 4u) ? p2 : t1 + p2;
        const double t3 = (f & 8u) ? p3 : t2 + p3;
        // ---- ... and across lanes (Kogge-Stone; lane l takes lane l-d iff no head lies in lanes (l-d, l])
        const uint32_t hb = __ballot_sync(0xffffffffu, f != 0u);
        const uint32_t below = hb & (0xffffffffu >> (31 - lane));
        const int ss = below ? 31 - __clz(below) : -1;          // last lane <= this one that holds a head
        double vs = t3;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, vs, d);
            if (lane - d >= ss && lane >= d) vs += y;
        }
        double cin = __shfl_up_sync(0xffffffffu, vs, 1);
        if (lane == 0) cin = 0.0;
        double2 u01, u23;
        u01.x = (f & 1u) ? t0 : t0 + cin;
        u01.y = (f & 3u) ? t1 : t1 + cin;
        u23.x = (f & 7u) ? t2 : t2 + cin;
        u23.y = (f & 15u) ? t3 : t3 + cin;
        carry = __shfl_sync(0xffffffffu, u23.y, 31);
        *reinterpret_cast<double2 *>(su + 4 * lane) = u01;
        *reinterpret_cast<double2 *>(su + 4 * lane + 2) = u23;
        __syncwarp();

        // ---- rows that end in this chunk: lane j finishes row wb+j
        for (;;) {
            const bool ends = (uint32_t)lane >= woff && win.PE != kPtrSentinel && win.PE <= endp;
            if (ends) {
                const double s = (win.PE != win.P) ? su[win.PE - 1u - base] : 0.0;
                epi.apply(a, (int64_t)wb + lane, s, win.O);
            }
            woff += (uint32_t)__popc(__ballot_sync(0xffffffffu, ends));
            if (woff >= 32u && wb + 32u < r1) {                // window exhausted: more rows may end here
                wb += 32u;
                woff = 0;
                win.load(a, epi, wb, r1, lane);
                continue;
            }
            break;
        }
        __syncwarp();
        // keep the window ahead of the stream: the reload is in flight during the next chunk
        if (endp >= e1) break;
        if (woff >= 16u && wb + woff < r1) {
            wb += woff;
            woff = 0;
            win.load(a, epi, wb, r1, lane);
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(kWThreads, kWMinBlocks)
spmv_warp_kernel(StreamArgs a)
{
    constexpr bool kFused = (EPI == SEPI_APROD || EPI == SEPI_ATPROD || EPI == SEPI_INIT_ATPROD || EPI == SEPI_ATPROD_UPD);
    __shared__ double s_red[kWWarps];
    __shared__ __align__(16) double s_u[kWWarps][kChunk];

    DevState *st = a.st;
    RowEpilogue<EPI> epi;
    int mode = MODE_FULL;
    if (kFused) {
        if (st->done) return;
        if (EPI == SEPI_APROD && st->istop != 0) return;   // stop already decided: only the deferred update is left
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
    }
    const int tid = threadIdx.x;

    if (mode != MODE_FULL) {
        // elementwise part only (n-vectors): x += t1 w [; w' = v/alpha + t2 w]
        if (epi.upd) {
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
        for (int t = (int)blockIdx.x * kWWarps + wib; t < a.map.ntiles; t += nw) {
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
                if (st->upd_pending) {
                    __threadfence();
                    step_after_update(*st, total_w, __ldcg(a.ux), a.ring);
                    st->upd_pending = 0;
                }
                if (!st->done) {
                    step_after_atprod(*st, total, mode == MODE_FULL);
                    st->upd_pending = 1;
                }
            }
        } else if (finish_reduction<kWThreads>(st, 0, epi.sq, s_red, &total)) {
            if (EPI == SEPI_APROD) {
                if (a.aux) *a.aux = total; else step_after_aprod(*st, total);
            }
            else if (EPI == SEPI_ATPROD) step_after_atprod(*st, total, true);
            else step_init_alpha(*st, total);
        }
    }
}

}  // namespace lsqrb
