// spmv_stream.cuh -- K3/K4, variant 2: tile-streamed CSR SpMV for sm_100a.
//
// The matrix is cut into row-aligned tiles of ~kTile stored entries.  A persistent CTA walks its
// tiles; for each tile ONE elected thread issues two 1-D TMA bulk copies (cp.async.bulk, SASS
// UBLKCP) that stream the tile's val[] and idx[] segments from HBM into a 3-stage shared-memory
// ring guarded by mbarriers, so the HBM stream never waits on the row structure and needs no
// registers.  All threads then (1) gather x[idx] from L2 and overwrite val with the products,
// (2) reduce the products row by row out of shared memory with G lanes per row (G chosen per tile
// from its mean row length) and apply the fused epilogue.  Tiles holding a row longer than the
// ring slot are processed straight from global memory by the whole CTA.
//
// Everything about the summation order is fixed by the tile map, never by scheduling, so results
// are run-to-run reproducible.  HBM traffic = val + idx once, ptr once, out once (+old once).
#pragma once

#include "kernels.cuh"

namespace lsqrb {

constexpr int kTile = 3072;      // nominal stored entries per tile
constexpr int kCap = 4096;       // ring-slot capacity (entries); a tile fits iff its rows are <= kCap - kTile + 1 long
constexpr int kCapS = kCap + 8;  // + alignment slack on both sides
constexpr int kStages = 2;
constexpr int kStreamThreads = 512;
constexpr int kPrefRounds = 2;   // thread-per-row mode: rounds whose ptr[] / out[] loads are issued before the gather phase
// one ring slot: val[kCapS] f64 | idx[kCapS] i32   (both multiples of 16 bytes)
constexpr size_t kSlotBytes = (size_t)kCapS * 8 + (size_t)kCapS * 4;
constexpr size_t kStreamSmem = (size_t)kStages * kSlotBytes + 64;
static_assert(kSlotBytes % 16 == 0 && (kCapS * 8) % 16 == 0, "TMA alignment");

struct TileMap {
    const uint2 *tiles;   // [ntiles+1]  {first row, first stored entry}; tiles[ntiles] = {nrows, nnz}
    int ntiles;
    // Warp kernel only: balanced schedule for matrices with very uneven rows.  Slot s = k * (warps of the grid) + w is
    // the k-th tile of warp w (kNoTile = none); nullptr = round robin (tile t belongs to warp t mod warps).
    const uint32_t *order;
    int nslots;
};
constexpr uint32_t kNoTile = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends the thread in hardware for a bounded time; a copy that never lands would
    // otherwise hang the GPU, so give up loudly after ~2 s instead.
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier; matrix data is read once,
// so it carries an L2 evict-first policy.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Tile map construction: tile t starts at the first row whose work coordinate is at or after t*tile.
// ---------------------------------------------------------------------------------------------
// ptr may be the pointer array of one block of a row-blocked transpose: its entries then start at ptr[0] != 0.
// Work coordinate of row r: W(r) = (ptr[r] - ptr[0]) + row_w * r  (stored entries before the row, plus row_w units
// per row: a run of EMPTY rows is work too -- every row gets its epilogue -- and must not land in one tile).
__global__ void build_tiles_kernel(const uint32_t *__restrict__ ptr, int64_t nrows, int ntiles, uint32_t tile, uint32_t row_w,
                                   uint2 *tiles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) { tiles[t] = make_uint2((uint32_t)nrows, ptr[nrows]); return; }
    const uint64_t base = ptr[0];
    const uint64_t target = (uint64_t)t * tile;
    int64_t lo = 0, hi = nrows;   // first r in [0, nrows] with W(r) >= target
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (((uint64_t)ptr[mid] - base) + (uint64_t)row_w * (uint64_t)mid >= target) hi = mid; else lo = mid + 1;
    }
    tiles[t] = make_uint2((uint32_t)lo, ptr[lo]);
}

// ---------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------
enum StreamEpilogue {
    SEPI_APROD = 0,        // u' = ca_mat*s + ca_vec*u ; sum u'^2 ; step_after_aprod (or partial to aux)
    SEPI_ATPROD = 1,       // v' = ct_mat*s + ct_vec*v ; sum v'^2 ; step_after_atprod
    SEPI_INIT_ATPROD = 2,  // v  = ct_mat*s            ; sum v^2  ; step_init_alpha
    SEPI_ACC = 3,          // out += s
    SEPI_STORE = 4,        // out  = s
    SEPI_ATPROD_UPD = 5,   // SEPI_ATPROD fused with the DEFERRED x/w update of the previous iteration:
                           //   x += t1 w ; w' = v/alpha + t2 w ; sum w'^2   (v = old v, read anyway)
    SEPI_APROD_ACC = 6     // last block of a column-blocked A: u' = ca_mat*(gu + s) + ca_vec*u ; sum u'^2 ; step_after_aprod
                           //   (gu = the partial A v of the earlier blocks, passed in uw)
};
constexpr bool sepi_is_aprod(int epi) { return epi == SEPI_APROD || epi == SEPI_APROD_ACC; }

struct StreamArgs {
    CsrView A;
    TileMap map;
    const double *x;       // gathered dense vector
    double *out;           // result vector (rows of A)
    DevState *st;
    double *aux;           // multi-GPU: where SEPI_APROD puts its partial sum(u'^2)
    double *ux, *uw, *use; // SEPI_ATPROD_UPD: solution, search direction, standard errors
    int out_aligned16;     // out[] may be the source of 16-byte aligned bulk copies
    int check_done;        // unfused STORE / ACC launched from the solve loop: nothing to do once the solver has stopped
    volatile lsqr_b200_iter_record *ring;
};

// LAZY: the two epilogue coefficients are re-read from the (kernel-invariant, L1-resident) device state at every
// use instead of living in registers for the whole kernel (the warp kernel is register-bound).
template <int EPI, bool LAZY = false>
struct RowEpilogue {
    double cm = 1.0, cv = 0.0;
    bool upd = false, wantse = false;
    double sq = 0.0, sq2 = 0.0;
    // coefficients of the deferred x/w update; elementwise fallback modes read them once, the row epilogue
    // re-reads them from the (L1-resident, kernel-invariant) device state to keep registers free
    double t1 = 0.0, t2 = 0.0, t3 = 0.0, ia = 1.0;

    __device__ __forceinline__ void load(const DevState *st)
    {
        if (!LAZY) {
            if (sepi_is_aprod(EPI)) { cm = st->ca_mat; cv = st->ca_vec; }
            if (EPI == SEPI_ATPROD || EPI == SEPI_INIT_ATPROD || EPI == SEPI_ATPROD_UPD) { cm = st->ct_mat; cv = st->ct_vec; }
        }
        if (EPI == SEPI_ATPROD_UPD) { upd = st->upd_pending != 0; wantse = st->wantse != 0; }
    }
    __device__ __forceinline__ double coef_mat(const DevState *st) const
    {
        if (!LAZY) return cm;
        return sepi_is_aprod(EPI) ? __ldg(&st->ca_mat) : __ldg(&st->ct_mat);
    }
    __device__ __forceinline__ double coef_vec(const DevState *st) const
    {
        if (!LAZY) return cv;
        return sepi_is_aprod(EPI) ? __ldg(&st->ca_vec) : __ldg(&st->ct_vec);
    }
    __device__ __forceinline__ void load_update_coefficients(const DevState *st)
    {
        t1 = st->t1; t2 = st->t2; t3 = st->t3; ia = st->inv_alpha;
    }
    // `old` = out[row] fetched before the row sum (hides the DRAM latency behind the reduction)
    __device__ __forceinline__ bool needs_old() const { return sepi_is_aprod(EPI) || EPI == SEPI_ATPROD || EPI == SEPI_ACC || EPI == SEPI_ATPROD_UPD; }
    // wo / xo: w[row], x[row] of the deferred update when the caller prefetched them (have_wx), else read here
    __device__ __forceinline__ void apply(const StreamArgs &a, int64_t row, double s, double old,
                                          bool have_wx = false, double wo = 0.0, double xo = 0.0)
    {
        if (EPI == SEPI_ACC) { a.out[row] = old + s; return; }
        if (EPI == SEPI_STORE) { a.out[row] = s; return; }
        if (EPI == SEPI_INIT_ATPROD) { const double r = coef_mat(a.st) * s; a.out[row] = r; sq += r * r; return; }
        if (EPI == SEPI_APROD_ACC) s += have_wx ? wo : a.uw[row];   // + the partial sum of the earlier column blocks
        const double r = coef_mat(a.st) * s + coef_vec(a.st) * old;
        a.out[row] = r;
        sq += r * r;
        if (EPI == SEPI_ATPROD_UPD && upd) {
            const DevState *st = a.st;
            const double c1 = __ldg(&st->t1), c2 = __ldg(&st->t2), cia = __ldg(&st->inv_alpha);
            if (!have_wx) { wo = a.uw[row]; xo = a.ux[row]; }
            a.ux[row] = c1 * wo + xo;
            const double wn = c2 * wo + cia * old;
            a.uw[row] = wn;
            sq2 += wn * wn;
            if (wantse) { const double c3 = __ldg(&st->t3); a.use[row] += (c3 * wo) * (c3 * wo); }
        }
    }
};

// Row reduction with G lanes per row (G = 8 or 32) out of the products in sval[]; entry e of the matrix
// sits at sval[e - a0].
template <int G, int EPI>
__device__ __forceinline__ void reduce_rows_group(const StreamArgs &a, RowEpilogue<EPI> &epi, const double *sval,
                                                  uint32_t r0, uint32_t r1, uint32_t a0)
{
    constexpr int kGroups = kStreamThreads / G;
    const int lane = threadIdx.x % G;
    const int grp = threadIdx.x / G;
    for (uint32_t rb = r0; rb < r1; rb += kGroups) {
        const uint32_t r = rb + grp;
        const bool valid = r < r1;
        double s = 0.0, old = 0.0;
        if (valid) {
            const uint32_t p0 = a.A.ptr[r] - a0, p1 = a.A.ptr[r + 1] - a0;
            if (lane == 0 && epi.needs_old()) old = a.out[r];
            for (uint32_t k = p0 + lane; k < p1; k += G) s += sval[k];
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (valid && lane == 0) epi.apply(a, r, s, old);
    }
}

// What a fused launch has to do, decided from the device state (uniform over the grid)
enum StreamMode {
    MODE_FULL = 0,         // stream the matrix, reduce rows, epilogue
    MODE_UPDATE_ONLY = 1,  // beta = 0: no A' product (src/lsqr.f90:691), but the deferred x/w update is due
    MODE_FLUSH = 2         // the stopping iteration is known: only x += t1 w of that iteration is left
};

template <int EPI>
__global__ void __launch_bounds__(kStreamThreads)
spmv_stream_kernel(StreamArgs a)
{
    constexpr bool kFused = (sepi_is_aprod(EPI) || EPI == SEPI_ATPROD || EPI == SEPI_INIT_ATPROD || EPI == SEPI_ATPROD_UPD);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kStages * kSlotBytes);
    auto slot_val = [&](int s) { return reinterpret_cast<double *>(smem_raw + (size_t)s * kSlotBytes); };
    auto slot_idx = [&](int s) { return reinterpret_cast<int32_t *>(slot_val(s) + kCapS); };
    __shared__ double s_red[kStreamThreads / 32];

    DevState *st = a.st;
    RowEpilogue<EPI> epi;
    int mode = MODE_FULL;
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
    } else if (a.check_done && st->done) {
        return;
    }

    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    const int tid = threadIdx.x;

    if (mode != MODE_FULL) {
        // elementwise part only (n-vectors): x += t1 w [; w' = v/alpha + t2 w]
        if (epi.upd) {
            epi.load_update_coefficients(st);
            const int64_t n = a.A.nrows;
            for (int64_t i = (int64_t)blockIdx.x * kStreamThreads + tid; i < n; i += (int64_t)gridDim.x * kStreamThreads) {
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
        if (tid == 0) {
            for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();

        const uint2 *tiles = a.map.tiles;
        const int ntiles = a.map.ntiles;
        const uint32_t *ptr = a.A.ptr;

        // producer (thread 0): stage tile t into ring slot `slot`; false if the tile needs no staging
        auto stage_tile = [&](int t, int slot) {
            const uint2 t0 = tiles[t], t1 = tiles[t + 1];
            const uint32_t len = t1.y - t0.y;
            if (t0.x == t1.x || len == 0 || len > (uint32_t)kCap) return false;   // rowless, empty or long tile
            const uint32_t a0 = t0.y & ~3u;
            const uint32_t cnt = ((t1.y + 3u) & ~3u) - a0;
            mbar_expect_tx(full + slot, cnt * 12u);
            tma_load_1d(slot_val(slot), a.A.val + a0, cnt * 8u, full + slot, pol_stream);
            tma_load_1d(slot_idx(slot), a.A.idx + a0, cnt * 4u, full + slot, pol_stream);
            return true;
        };

        int prod_t = blockIdx.x;          // producer's next tile
        int prod_n = 0;                   // tiles staged so far
        int cons_n = 0;                   // staged tiles consumed so far
        if (tid == 0) {
            while (prod_n < kStages && prod_t < ntiles) {
                if (stage_tile(prod_t, prod_n % kStages)) ++prod_n;
                prod_t += gridDim.x;
            }
        }

        uint2 d0 = make_uint2(0, 0), d1 = make_uint2(0, 0);
        if ((int)blockIdx.x < ntiles) { d0 = tiles[blockIdx.x]; d1 = tiles[blockIdx.x + 1]; }
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const uint32_t r0 = d0.x, r1 = d1.x, e0 = d0.y, e1 = d1.y;
            {   // descriptor of the next tile: in flight while this one is processed
                const int tn = t + gridDim.x;
                if (tn < ntiles) { d0 = tiles[tn]; d1 = tiles[tn + 1]; }
            }
            const uint32_t len = e1 - e0;
            if (r0 == r1) continue;                       // no row starts in this tile (inside a long row)
            if (len > (uint32_t)kCap) {
                // ---- long tile: every row straight from global memory, whole CTA per row
                for (uint32_t r = r0; r < r1; ++r) {
                    const uint32_t p0 = ptr[r], p1 = ptr[r + 1];
                    double old = 0.0;
                    if (tid == 0 && epi.needs_old()) old = a.out[r];
                    double s = 0.0;
#pragma unroll 4
                    for (uint32_t k = p0 + tid; k < p1; k += kStreamThreads) {
                        const double v = ldg_stream_f64(a.A.val + k, pol_stream);
                        const int32_t c = ldg_stream_s32(a.A.idx + k, pol_stream);
                        s += v * ldg_keep_f64(a.x + c, pol_keep);
                    }
                    s = block_sum<kStreamThreads>(s, s_red);
                    if (tid == 0) epi.apply(a, r, s, old);
                }
                continue;
            }
            const uint32_t a0 = e0 & ~3u;
            const uint32_t nrows = r1 - r0;
            const bool thread_per_row = len <= 24u * nrows;          // mean row length <= 24
            // thread-per-row mode: issue the row-pointer / old-value loads of the first rounds now, so their
            // DRAM latency hides behind the gather phase
            uint32_t pp0[kPrefRounds], pp1[kPrefRounds];
            double pold[kPrefRounds];
            if (thread_per_row) {
#pragma unroll
                for (int j = 0; j < kPrefRounds; ++j) {
                    const uint32_t r = r0 + tid + j * kStreamThreads;
                    pp0[j] = pp1[j] = 0u; pold[j] = 0.0;
                    if (r < r1) {
                        pp0[j] = ptr[r]; pp1[j] = ptr[r + 1];
                        if (epi.needs_old()) pold[j] = a.out[r];
                    }
                }
            }
            const int slot = cons_n % kStages;
            double *sv = slot_val(slot);
            if (len > 0) {
                const int32_t *si = slot_idx(slot);
                mbar_wait(full + slot, (uint32_t)(cons_n / kStages) & 1u);
                // ---- gather + multiply in place, four consecutive entries per thread and step.  The aligned
                // superset [a0, a1) may hold a few entries of the neighbouring tiles: harmless extra products.
                const uint32_t ngroups = (((e1 + 3u) & ~3u) - a0) >> 2;
#pragma unroll 2
                for (uint32_t gq = tid; gq < ngroups; gq += kStreamThreads) {
                    const int4 c = *reinterpret_cast<const int4 *>(si + 4 * gq);
                    double2 v01 = *reinterpret_cast<const double2 *>(sv + 4 * gq);
                    double2 v23 = *reinterpret_cast<const double2 *>(sv + 4 * gq + 2);
                    const double x0 = ldg_keep_f64(a.x + c.x, pol_keep);
                    const double x1 = ldg_keep_f64(a.x + c.y, pol_keep);
                    const double x2 = ldg_keep_f64(a.x + c.z, pol_keep);
                    const double x3 = ldg_keep_f64(a.x + c.w, pol_keep);
                    v01.x *= x0; v01.y *= x1; v23.x *= x2; v23.y *= x3;
                    *reinterpret_cast<double2 *>(sv + 4 * gq) = v01;
                    *reinterpret_cast<double2 *>(sv + 4 * gq + 2) = v23;
                }
                __syncthreads();
            }
            // ---- row reduction out of shared memory
            if (thread_per_row) {
#pragma unroll
                for (int j = 0; j < kPrefRounds; ++j) {
                    const uint32_t r = r0 + tid + j * kStreamThreads;
                    if (r < r1) {
                        double s = 0.0;
                        for (uint32_t k = pp0[j] - a0; k < pp1[j] - a0; ++k) s += sv[k];
                        epi.apply(a, r, s, pold[j]);
                    }
                }
                for (uint32_t r = r0 + tid + kPrefRounds * kStreamThreads; r < r1; r += kStreamThreads) {
                    const uint32_t p0 = ptr[r] - a0, p1 = ptr[r + 1] - a0;
                    const double old = epi.needs_old() ? a.out[r] : 0.0;
                    double s = 0.0;
                    for (uint32_t k = p0; k < p1; ++k) s += sv[k];
                    epi.apply(a, r, s, old);
                }
            } else if (len <= 96u * nrows) {
                reduce_rows_group<8, EPI>(a, epi, sv, r0, r1, a0);
            } else {
                reduce_rows_group<32, EPI>(a, epi, sv, r0, r1, a0);
            }
            if (len > 0) {
                fence_proxy_async_smem();                 // order this thread's generic-proxy writes (products) before the async refill
                __syncthreads();                          // everyone is done with this slot
                ++cons_n;
                if (tid == 0) {
                    while (prod_t < ntiles) {
                        const bool staged = stage_tile(prod_t, prod_n % kStages);
                        prod_t += gridDim.x;
                        if (staged) { ++prod_n; break; }
                    }
                }
            }
        }
    }

    if (kFused) {
        double total, total_w;
        if (EPI == SEPI_ATPROD_UPD) {
            // sum(w'^2) of the deferred update closes iteration k; sum(v'^2) drives the step of iteration k+1
            if (finish_reduction2<kStreamThreads>(st, 0, epi.sq, epi.sq2, s_red, &total, &total_w)) {
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
        } else if (finish_reduction<kStreamThreads>(st, 0, epi.sq, s_red, &total)) {
            if (sepi_is_aprod(EPI)) {
                if (a.aux) *a.aux = total; else step_after_aprod(*st, total);
            }
            else if (EPI == SEPI_ATPROD) step_after_atprod(*st, total, true);
            else step_init_alpha(*st, total);
        }
    }
}

}  // namespace lsqrb
