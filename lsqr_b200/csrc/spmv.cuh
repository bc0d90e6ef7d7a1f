// spmv.cuh -- K3/K4: the warp-autonomous segmented CSR SpMV of the engine, hand-written for sm_100a.
//
// ONE persistent launch computes one whole product, A v or A'u, even when the matrix is stored blocked along the
// gathered coordinate (column blocks of A / row blocks of A', so that the gathered slice of the dense vector stays
// in L2).  The rows are cut into row-aligned tiles, the SAME row ranges in every block; a warp owns the same tiles
// in every block and walks the blocks in order, so "block b+1 accumulates onto what block b stored" is program
// order inside one warp and needs no atomics.  A guard on the GATHERED side makes block b start when every CTA has
// finished block b-1 (one thread per CTA polls a counter), so exactly one L2-sized slice of the dense vector is live;
// the grid is proven co-resident at initialize (probe mode below, plan.cuh).
//
// Three flavours of the same kernel (template parameter FLAV, chosen per plan): FLAV_GATHER for random columns
// (first use of the gathered values behind the head mask, scan depth adapted to the chunk), FLAV_LOCAL for banded
// matrices whose gathers hit L1 (multiply early, full scan), FLAV_WINDOW for the opt-in staged gather windows.
//
// A tile of one block (a "piece") is consumed in chunks of 128 stored entries, 4 per lane:
//   * val[] / idx[] of chunk c+1 are fetched with one 256-bit and one 128-bit streaming load per lane (evict-first)
//     while chunk c is reduced (register double buffer);
//   * x[idx]: 4 independent global gathers per lane through L1/L2 (evict-last); FLAV_WINDOW only: if the piece's
//     gathered indices span a window that fits the warp's shared-memory buffer (the span is recorded per piece at
//     initialize), the window is staged with coalesced loads and the gathers are shared-memory reads (measured
//     bank-conflict-bound and no faster on the banded family: opt-in);
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
    int pdl;                   // launched as a programmatic dependent of the previous kernel: see spmv_kernel
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
    template <int MODE>
    __device__ __forceinline__ void load(const SpmvArgs &a, const BlockCtx &bc, uint32_t wb, uint32_t r1, int lane)
    {
        const uint32_t r = wb + (uint32_t)lane;
        P = PE = kPtrSentinel;
        O = 0.0;
        G = 0.0;
        if (r < r1) {
            P = bc.ptr[r];
            PE = bc.ptr[r + 1];
            if (MODE == BM_ACC) O = a.part[r];
            if (FIN != FIN_NONE && MODE == BM_FINAL) {
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
    template <int MODE>
    __device__ __forceinline__ void apply(const SpmvArgs &a, uint32_t row, double s, double old, double g)
    {
        if (MODE == BM_STORE) { a.part[row] = s; return; }
        if (MODE == BM_ACC) { a.part[row] = old + s; return; }
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

// One lane's share of a chunk: EPL consecutive stored entries (a chunk is 32 * EPL entries).  Only EPL = 4 is
// instantiated (plan.cuh: kEpl): 24 registers of double buffer, 4 CTAs / 32 warps per SM -- the L1TEX pipe is the
// limit and warps are what hides its latency.  An EPL = 8 build (the per-chunk overhead once per 256 entries, 2 CTAs /
// 16 warps per SM, <= 128 registers) was measured on all four families and was slower or equal everywhere
// (profiles/r02/run5/spmv_bench_epl_window_ab.jsonl); the template stays generic in EPL.
template <int EPL>
struct ChunkRegs {
    double v[EPL];
    int32_t c[EPL];
};

template <int EPL>
__device__ __forceinline__ void ldg_chunk(const SpmvArgs &a, uint32_t q, uint64_t pol_stream, ChunkRegs<EPL> &r)
{
    if constexpr (EPL == 4) {
        ldg_stream_f64x4(a.val + q, *reinterpret_cast<double (*)[4]>(&r.v[0]));
        ldg_stream_s32x4(a.idx + q, *reinterpret_cast<int32_t (*)[4]>(&r.c[0]), pol_stream);
    } else {
        ldg_stream_f64x4(a.val + q, *reinterpret_cast<double (*)[4]>(&r.v[0]));
        ldg_stream_f64x4(a.val + q + 4, *reinterpret_cast<double (*)[4]>(&r.v[EPL - 4]));
        ldg_stream_s32x8(a.idx + q, *reinterpret_cast<int32_t (*)[8]>(&r.c[0]));
    }
}

// Loads one lane's share of the chunk that starts at entry `cb` of the piece [e0, e1): entries [cb + EPL lane, + EPL).
// Entries outside the piece (they belong to the neighbouring pieces or to the allocation slack) get the value 0 and
// the index `safe` (inside the piece's gather window), so they add nothing to any row and gather harmlessly.
template <int EPL>
__device__ __forceinline__ void load_chunk(const SpmvArgs &a, uint32_t cb, int lane, uint32_t e0, uint32_t e1,
                                           int32_t safe, uint64_t pol_stream, ChunkRegs<EPL> &r)
{
    const uint32_t q = cb + (uint32_t)EPL * (uint32_t)lane;
    if (cb >= e0 && cb + 32u * EPL <= e1) {   // steady state: the whole chunk lies inside the piece
        ldg_chunk<EPL>(a, q, pol_stream, r);
        return;
    }
    if (q < e1) {
        ldg_chunk<EPL>(a, q, pol_stream, r);
#pragma unroll
        for (int k = 0; k < EPL; ++k)
            if (q + k < e0 || q + k >= e1) { r.v[k] = 0.0; r.c[k] = safe; }
    } else {
#pragma unroll
        for (int k = 0; k < EPL; ++k) { r.v[k] = 0.0; r.c[k] = safe; }
    }
}

// The per-warp row-sum buffer is addressed in the shared state space with a 32-bit address computed once per kernel:
// with a generic pointer the compiler re-derives the shared window (S2UR SR_CgaCtaId + ULEA) in front of every access
// inside the chunk loop, two MIO round trips per chunk on the critical path of the row-end pass.
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double x, double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// Byte offset of the running sum of chunk entry e (= EPL * lane + k) in the warp's buffer.  The buffer is laid out in
// EPL/2 planes of 32 x 16 bytes -- plane k/2 holds the pair (t[k], t[k+1]) of every lane, lane-contiguous -- so that each
// warp-wide STS.128 writes 512 CONTIGUOUS bytes: 4 conflict-free wavefronts.  The natural layout (entry e at 8 e) puts
// the lanes' 16-byte pieces 32 bytes apart, which ncu counted as 9.4 wavefronts per STS.128 (profiles/r02/final/
// c5q_gather_counters.txt: 18.7 store wavefronts per chunk), on the L1TEX data stage this kernel saturates.
template <int EPL>
__device__ __forceinline__ uint32_t row_sum_slot(uint32_t e)
{
    const uint32_t lane = e / (uint32_t)EPL, k = e % (uint32_t)EPL;
    return (k >> 1) * 512u + lane * 16u + (k & 1u) * 8u;
}

template <int FIN>
struct WarpTileState {
    uint32_t wb, woff;     // window base row, lanes below woff are finished rows
    RowWindow<FIN> win;
    double carry;          // running sum of the row that is open at the chunk boundary
};

// Reduces the chunk [base, base + 32 EPL): t[] = the lane's EPL gathered vector entries, cv[] = its stored values.
template <int FIN, int EPL, int MODE, bool LATE>
__device__ __forceinline__ void warp_chunk_core(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, uint32_t su,
                                                WarpTileState<FIN> &ts, double (&t)[EPL], const double (&cv)[EPL],
                                                uint32_t base, uint32_t r1, uint32_t e1, int lane)
{
    constexpr uint32_t kC = 32u * EPL;
    const uint32_t endp = base + kC;
    // ---- head mask of the chunk: bit i set <=> a row starts at entry base+i  (EPL words of 32 bits)
    uint32_t m[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) m[k] = 0u;
    {
        uint32_t p = ts.win.P;
        uint32_t wb2 = ts.wb;
        for (;;) {
            const uint32_t rel = p - base;
            const uint32_t bit = rel < kC ? (1u << (rel & 31u)) : 0u;
            const uint32_t w = rel >> 5;
#pragma unroll
            for (int k = 0; k < EPL; ++k) m[k] |= __reduce_or_sync(0xffffffffu, w == (uint32_t)k ? bit : 0u);
            const uint32_t p31 = __shfl_sync(0xffffffffu, p, 31);
            if (p31 >= endp || wb2 + 32u >= r1) break;     // (the sentinel ends the walk too)
            wb2 += 32u;                                    // more than a window of rows starts in this chunk
            const uint32_t r = wb2 + (uint32_t)lane;
            p = r < r1 ? bc.ptr[r] : kPtrSentinel;
        }
    }
    // the lane's EPL head flags: word (lane * EPL) / 32, bit offset (lane * EPL) % 32
    uint32_t mw = m[0];
#pragma unroll
    for (int k = 1; k < EPL; ++k) mw = ((uint32_t)lane * EPL) / 32u == (uint32_t)k ? m[k] : mw;
    const uint32_t f = (mw >> (((uint32_t)lane * EPL) & 31u)) & ((1u << EPL) - 1u);

    // ---- products, gather-bound flavour (LATE).  First use of the gathered values: it comes AFTER the head mask so
    // that the mask's ~70 instructions and MIO round trips overlap the gather latency (with the multiply in front of
    // the mask every warp stalled there on a cold gather: 4 % on the random-column families).  Matrices whose gathers
    // hit L1 (banded) want the opposite -- the multiplies in flight while the mask's REDUX results travel -- and
    // measured 12 % slower with the late multiply, so warp_chunk multiplies for them.
    if (LATE) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) t[k] *= cv[k];
    }

    // ---- segmented running sums inside the lane (t_k = sum of the lane's entries of the segment that entry k
    // belongs to, up to and including k)
    if (lane == 0 && !(f & 1u)) t[0] = ts.carry + t[0];    // row that began in an earlier chunk
#pragma unroll
    for (int k = 1; k < EPL; ++k)
        if (!(f & (1u << k))) t[k] += t[k - 1];
    // ---- ... and across lanes (Kogge-Stone; lane l takes lane l-d iff no head lies in lanes (l-d, l])
    const uint32_t hb = __ballot_sync(0xffffffffu, f != 0u);
    const uint32_t below = hb & (0xffffffffu >> (31 - lane));
    const int reach = lane - (below ? 31 - __clz(below) : 0);   // how far down this lane's open segment extends
    // A step of distance d adds nothing when no lane's segment reaches d lanes down: short rows (20 entries = 5
    // lanes) skip the steps of 8 and 16 lanes.  Shuffles share the L1TEX data stage with the gathers, which is the
    // saturated unit of this kernel, so the skipped steps are throughput, not only latency.
    // (gather-bound flavour only: on the banded family the extra reduction and branches cost 3-5 %)
    const int span = LATE ? (int)__reduce_max_sync(0xffffffffu, (unsigned)reach) : 31;
    double vs = t[EPL - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        if (d <= span) {                                   // warp-uniform
            const double y = __shfl_up_sync(0xffffffffu, vs, d);
            if (d <= reach) vs += y;
        }
    }
    double cin = __shfl_up_sync(0xffffffffu, vs, 1);
    if (lane == 0) cin = 0.0;
    // entries that precede the lane's first head continue the segment of the lanes below
#pragma unroll
    for (int k = 0; k < EPL; ++k)
        if (!(f & ((2u << k) - 1u))) t[k] += cin;
    ts.carry = __shfl_sync(0xffffffffu, t[EPL - 1], 31);
    // ---- rows that end in this chunk: lane j finishes row wb+j (the sentinel of absent rows never ends).  A chunk in
    // which no row ends -- a third to a half of the chunks of a product with 200-250 entries per row -- skips the
    // shared-memory round trip altogether (two STS.128, two warp barriers).
    if (__any_sync(0xffffffffu, (uint32_t)lane >= ts.woff && ts.win.PE <= endp)) {
#pragma unroll
        for (int k = 0; k < EPL; k += 2) sts_f64x2(su + row_sum_slot<EPL>((uint32_t)(EPL * lane + k)), t[k], t[k + 1]);
        __syncwarp();
        for (;;) {
            const bool ends = (uint32_t)lane >= ts.woff && ts.win.PE <= endp;
            if (ends) {
                const double s = (ts.win.PE != ts.win.P) ? lds_f64(su + row_sum_slot<EPL>(ts.win.PE - 1u - base)) : 0.0;
                epi.template apply<MODE>(a, ts.wb + (uint32_t)lane, s, ts.win.O, ts.win.G);
            }
            ts.woff += (uint32_t)__popc(__ballot_sync(0xffffffffu, ends));
            if (ts.woff >= 32u && ts.wb + 32u < r1) {          // window exhausted: more rows may end here
                ts.wb += 32u;
                ts.woff = 0;
                ts.win.template load<MODE>(a, bc, ts.wb, r1, lane);
                continue;
            }
            break;
        }
        __syncwarp();                                      // the next chunk rewrites the buffer
    }
    // keep the window ahead of the stream: the reload is in flight during the next chunk
    if (endp < e1 && ts.woff >= 16u && ts.wb + ts.woff < r1) {
        ts.wb += ts.woff;
        ts.woff = 0;
        ts.win.template load<MODE>(a, bc, ts.wb, r1, lane);
    }
}

// Processes the chunk [base, base + 32 EPL) held in `cur`; `nxt` receives the following chunk, whose loads stay in
// flight while this one is reduced.  WIN: the gathers are reads of the warp's staged window.
template <int FIN, bool WIN, int EPL, int MODE, bool LATE>
__device__ __forceinline__ void warp_chunk(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, uint32_t su,
                                           const double *xw /* WIN: window - win_lo */, WarpTileState<FIN> &ts,
                                           const ChunkRegs<EPL> &cur, ChunkRegs<EPL> &nxt, uint32_t base, uint32_t r1,
                                           uint32_t e0, uint32_t e1, int32_t safe, int lane,
                                           uint64_t pol_stream, uint64_t pol_keep)
{
    double t[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) t[k] = WIN ? xw[cur.c[k]] : ldg_keep_f64(a.x + cur.c[k], pol_keep);
    load_chunk<EPL>(a, base + 32u * EPL, lane, e0, e1, safe, pol_stream, nxt);
    if (!LATE) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) t[k] *= cur.v[k];
    }
    warp_chunk_core<FIN, EPL, MODE, LATE>(a, bc, epi, su, ts, t, cur.v, base, r1, e1, lane);
}

template <int FIN, bool WIN, int EPL, int MODE, bool LATE>
__device__ __forceinline__ void warp_tile_loop(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, uint32_t su,
                                               const double *xw, uint32_t r0, uint32_t r1, uint32_t e0, uint32_t e1,
                                               int32_t safe, int lane, uint64_t pol_stream, uint64_t pol_keep)
{
    const uint32_t a0 = e0 & ~(uint32_t)(EPL - 1);
    WarpTileState<FIN> ts;
    ts.wb = r0;
    ts.woff = 0;
    ts.carry = 0.0;
    ts.win.template load<MODE>(a, bc, r0, r1, lane);
    ChunkRegs<EPL> ra, rb;
    load_chunk<EPL>(a, a0, lane, e0, e1, safe, pol_stream, ra);
    // two chunks per trip so that the register double buffer needs no copies; at least one chunk is processed
    // even for a piece without entries, so that its (empty) rows still get their epilogue
    for (uint32_t base = a0;;) {
        warp_chunk<FIN, WIN, EPL, MODE, LATE>(a, bc, epi, su, xw, ts, ra, rb, base, r1, e0, e1, safe, lane, pol_stream, pol_keep);
        base += 32u * EPL;
        if (base >= e1) break;
        warp_chunk<FIN, WIN, EPL, MODE, LATE>(a, bc, epi, su, xw, ts, rb, ra, base, r1, e0, e1, safe, lane, pol_stream, pol_keep);
        base += 32u * EPL;
        if (base >= e1) break;
    }
}

// One piece: stage the gather window if the piece has one, then stream its chunks.
template <int FIN, int EPL, bool WINS, bool LATE>
__device__ __forceinline__ void warp_tile(const SpmvArgs &a, const BlockCtx &bc, Epilogue<FIN> &epi, uint32_t su, double *wbuf,
                                          const TileDesc &d0, const TileDesc &d1, int lane,
                                          uint64_t pol_stream, uint64_t pol_keep)
{
    // The block mode is a template parameter of the chunk loop (one clone per mode, only one of them hot at a time):
    // with the mode tested per finished row the loop ran 20 % more instructions, 3.6 % more time on C5.
    if (WINS && d0.win_len != 0u) {
        // coalesced 8-byte loads: no alignment requirement on x, which may be the caller's own array (aprod)
        const double *src = a.x + d0.win_lo;
        for (uint32_t i = (uint32_t)lane; i < d0.win_len; i += 32u) wbuf[i] = ldg_keep_f64(src + i, pol_keep);
        __syncwarp();
        const double *xw = wbuf - d0.win_lo;
        const int32_t safe = (int32_t)d0.win_lo;
        if (FIN != FIN_NONE && bc.mode == BM_FINAL)
            warp_tile_loop<FIN, true, EPL, BM_FINAL, LATE>(a, bc, epi, su, xw, d0.row, d1.row, d0.entry, d1.entry, safe, lane, pol_stream, pol_keep);
        else if (bc.mode == BM_ACC)
            warp_tile_loop<FIN, true, EPL, BM_ACC, LATE>(a, bc, epi, su, xw, d0.row, d1.row, d0.entry, d1.entry, safe, lane, pol_stream, pol_keep);
        else
            warp_tile_loop<FIN, true, EPL, BM_STORE, LATE>(a, bc, epi, su, xw, d0.row, d1.row, d0.entry, d1.entry, safe, lane, pol_stream, pol_keep);
    } else {
        if (FIN != FIN_NONE && bc.mode == BM_FINAL)
            warp_tile_loop<FIN, false, EPL, BM_FINAL, LATE>(a, bc, epi, su, nullptr, d0.row, d1.row, d0.entry, d1.entry, 0, lane, pol_stream, pol_keep);
        else if (bc.mode == BM_ACC)
            warp_tile_loop<FIN, false, EPL, BM_ACC, LATE>(a, bc, epi, su, nullptr, d0.row, d1.row, d0.entry, d1.entry, 0, lane, pol_stream, pol_keep);
        else
            warp_tile_loop<FIN, false, EPL, BM_STORE, LATE>(a, bc, epi, su, nullptr, d0.row, d1.row, d0.entry, d1.entry, 0, lane, pol_stream, pol_keep);
    }
}

// EPL: stored entries per lane and chunk (see ChunkRegs).
// FLAV: kernel flavour, chosen per plan (plan.cuh): FLAV_LOCAL = gathers that mostly hit L1 (banded matrices),
// FLAV_WINDOW = staged gather windows (opt-in), FLAV_GATHER = random columns (late multiply, adaptive scan depth).
enum SpmvFlavour { FLAV_LOCAL = 0, FLAV_WINDOW = 1, FLAV_GATHER = 2 };

template <int FIN, int EPL, int FLAV>
__global__ void __launch_bounds__(kWThreads, EPL == 4 ? 4 : 2)
spmv_kernel(SpmvArgs a)
{
    constexpr bool WINS = FLAV == FLAV_WINDOW;
    constexpr bool LATE = FLAV == FLAV_GATHER;
    constexpr bool kFused = (FIN == FIN_APROD || FIN == FIN_ATPROD || FIN == FIN_INIT_ATPROD);
    extern __shared__ __align__(128) double s_dyn[];         // per warp: su[32 EPL] | gather window[win_cap]; 128-byte aligned: a warp-wide STS.128 then touches 4 lines, not 5
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
    if (a.pdl) {
        // Programmatic dependent launch: this grid was allowed to start while the previous kernel of the stream drains
        // (its last CTAs, its grid reduction and the scalar step).  Nothing the previous kernel writes may be read
        // before griddepcontrol.wait; the matrix is constant, so the warp's first chunk is pulled towards L2 meanwhile.
        asm volatile("griddepcontrol.launch_dependents;");
        const int s0 = (int)blockIdx.x * kWWarps + (tid >> 5);
        const int ns = a.order ? a.nslots : a.ntiles;
        if (s0 < ns) {
            const uint32_t t = a.order ? a.order[s0] : (uint32_t)s0;
            if (t != kNoTile) {
                const TileDesc d0 = a.tiles[t];
                const uint32_t q = (d0.entry & ~3u) + 4u * (uint32_t)(tid & 31);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.val + q));
                if ((tid & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.idx + q));
                if ((tid & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ptr + d0.row));
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
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
        double *su_g = s_dyn + (size_t)wib * (32u * EPL + (WINS ? (size_t)a.win_cap : (size_t)0));
        double *wbuf = su_g + 32 * EPL;
        uint32_t su = (uint32_t)__cvta_generic_to_shared(su_g);
        asm volatile("" : "+r"(su));     // opaque from here on: kept in a register, not re-derived from %tid in the loop
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
                // they were waiting for: 6 % on C5/4); the grid was proven co-resident by the probe at initialize (plan.cuh).
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
                warp_tile<FIN, EPL, WINS, LATE>(a, bc, epi, su, wbuf, d0, d1, lane, pol_stream, pol_keep);
            }
            if (a.guard && b + 1 < a.nblocks) {
                __syncthreads();                                // every warp of this CTA has finished block b
                if (tid == 0) atomicAdd(&st->blk_done[b], 1u);
            }
        }
    }

    if (kFused) {
        Ssq total;
        // the step after A'u (rotations, estimates, stopping tests) is run by a whole warp: step_after_atprod_impl
        if (finish_ssq<kWThreads, FIN == FIN_ATPROD && kWarpStep>(st, 0, st->partial, epi.sq, s_exc, s_red, &total)) {
            if (tid == 0) {
                if (a.guard) for (int b = 0; b + 1 < a.nblocks; ++b) st->blk_done[b] = 0;
                if (tracing) st->trace[1][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns();
            }
            if (FIN == FIN_APROD) {
                if (a.aux) *a.aux = total; else step_after_aprod(*st, ssq_norm(total));
            }
            else if (FIN == FIN_ATPROD) {
                if (kWarpStep) step_after_atprod_warp(*st, ssq_norm(total), true, tid & 31);
                else           step_after_atprod(*st, ssq_norm(total), true);
            }
            else step_init_alpha(*st, ssq_norm(total));
            if (tracing && tid == 0) { st->trace[2][st->tr_n & (kTraceSlots - 1)] = globaltimer_ns(); st->tr_n += 1; }
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
