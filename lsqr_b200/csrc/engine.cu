// engine.cu -- host driver + C ABI of the B200 LSQR engine (include/lsqr_b200.h).
//
// One handle owns: CSR(A) and CSR(A') in HBM, the work vectors u(m), v,w,x,(se)(n), a
// device-resident scalar state (DevState), and a ring of per-iteration records in pinned mapped
// host memory.  The host never computes a scalar of the recurrence: it enqueues batches of
// iterations (CUDA graph), and reads the records to learn when the device decided to stop.
#include <dlfcn.h>
#include <nccl.h>   // types only; the library is dlopen'ed so single-GPU use has no NCCL dependency

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <queue>
#include <vector>

#include "build_csr.h"
#include "kernels.cuh"
#include "spmv_stream.cuh"
#include "spmv_warp.cuh"

namespace lsqrb {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// ---------------------------------------------------------------------------------------------
// NCCL, loaded at run time
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // Prefer a copy that is already in the process (torch loads its bundled libnccl.so.2).
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        if (api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy) api.lib = h;
    });
    return api.lib ? &api : nullptr;
}

#define LSQRB_NCCL(call)                                                                          \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            NcclApi *a__ = nccl_api();                                                            \
            set_last_error(std::string(#call) + ": " +                                            \
                           ((a__ && a__->GetErrorString) ? a__->GetErrorString(r__) : "nccl error")); \
            return LSQR_B200_ERR_NCCL;                                                            \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Fortran-style number formatting for the nout log (1PEw.d)
// ---------------------------------------------------------------------------------------------
static std::string fe(int w, int d, double v)
{
    char tmp[64];
    snprintf(tmp, sizeof tmp, "%.*E", d, v);
    std::string s(tmp);
    size_t e = s.find('E');
    if (e != std::string::npos && s.size() - (e + 2) >= 3) s.erase(e, 1);   // E+100 -> +100
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)w - s.size(), ' ') + s;
}

// ---------------------------------------------------------------------------------------------
// Work: stream + device state + record ring, shared by the ez path and the operator-hook path
// ---------------------------------------------------------------------------------------------
struct Work {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    DevState *st = nullptr;                       // device
    DevState h;                                   // host mirror (header part only is copied)
    lsqr_b200_iter_record *ring_h = nullptr;      // pinned, mapped
    lsqr_b200_iter_record *ring_d = nullptr;      // device alias of ring_h
    int max_grid = kNumSMs * 8;
    int stream_grid = kNumSMs * 3;   // cap on the persistent CTAs of the tile-streamed kernels
    int sms = kNumSMs;
    int64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    int init(int dev, void *user_stream)
    {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            set_last_error("no CUDA device is visible; this engine has no CPU path");
            return LSQR_B200_ERR_NO_DEVICE;
        }
        if (dev < 0) LSQRB_CUDA(cudaGetDevice(&dev));
        if (dev >= count) { set_last_error("device ordinal out of range"); return LSQR_B200_ERR_ARG; }
        device = dev;
        LSQRB_CUDA(cudaSetDevice(device));
        if (user_stream) {
            stream = (cudaStream_t)user_stream;
        } else {
            LSQRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            own_stream = true;
        }
        LSQRB_CUDA(cudaMalloc(&st, sizeof(DevState)));
        LSQRB_CUDA(cudaMemsetAsync(st, 0, sizeof(DevState), stream));
        LSQRB_CUDA(cudaHostAlloc(&ring_h, sizeof(lsqr_b200_iter_record) * kRingSize, cudaHostAllocMapped));
        LSQRB_CUDA(cudaHostGetDevicePointer(&ring_d, ring_h, 0));
        for (auto &e : ev) LSQRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDefault));
        LSQRB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        sms = std::max(1, sms);
        max_grid = std::min(kMaxPartials, std::max(1, sms) * env_int("LSQR_B200_BLOCKS_PER_SM", 8));
        stream_grid = std::min(kMaxPartials, std::max(1, sms) * env_int("LSQR_B200_STREAM_CTAS_PER_SM", 3));
        return LSQR_B200_OK;
    }

    void destroy()
    {
        if (st) cudaFree(st);
        if (ring_h) cudaFreeHost(ring_h);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        if (own_stream && stream) cudaStreamDestroy(stream);
        st = nullptr; ring_h = nullptr; stream = nullptr;
    }

    int grid_for(int64_t items, int per_block) const
    {
        int64_t b = (items + per_block - 1) / per_block;
        if (b < 1) b = 1;
        return (int)std::min<int64_t>(b, max_grid);
    }

    // reset the scalar state for a new solve (src/lsqr.f90:597-617)
    int reset_state(double damp, double atol, double btol, double conlim, int itnlim, int wantse, int dist)
    {
        memset(&h, 0, offsetof(DevState, partial));
        h.damp = damp; h.atol = atol; h.btol = btol;
        h.ctol = conlim > 0.0 ? 1.0 / conlim : 0.0;
        h.itnlim = itnlim; h.wantse = wantse; h.damped = damp > 0.0; h.dist = dist;
        h.tr_on = env_int("LSQR_B200_TRACE", 0); h.tr_n = 0;
        h.cs2 = -1.0;
        h.inv_alpha = h.inv_beta = 1.0;
        h.g_c0 = h.g_c1 = h.g_c2 = h.g_c3 = 1.0;
        LSQRB_CUDA(cudaMemcpyAsync(st, &h, offsetof(DevState, partial), cudaMemcpyHostToDevice, stream));
        for (int i = 0; i < kRingSize; ++i) ring_h[i].itn = -1.0;
        return LSQR_B200_OK;
    }

    int fetch_state()
    {
        LSQRB_CUDA(cudaMemcpyAsync(&h, st, offsetof(DevState, partial), cudaMemcpyDeviceToHost, stream));
        LSQRB_CUDA(cudaStreamSynchronize(stream));
        if (h.tr_on) dump_trace();
        return LSQR_B200_OK;
    }

    // LSQR_B200_TRACE=1: timeline of the fused kernels of the last solve, to stderr (microseconds)
    void dump_trace()
    {
        const int n = std::min(h.tr_n, kTraceSlots);
        std::vector<unsigned long long> t((size_t)4 * kTraceSlots);
        if (cudaMemcpy(t.data(), (const char *)st + offsetof(DevState, trace), sizeof(unsigned long long) * t.size(),
                       cudaMemcpyDeviceToHost) != cudaSuccess) return;
        fprintf(stderr, "[lsqr_b200 trace] %d fused kernels: idx start_us busy_us step_us gap_to_next_us upd_step_us\n", n);
        for (int k = 0; k < n; ++k) {
            const double t0 = (double)(t[k] - t[0]) * 1e-3;
            const double busy = (double)(t[kTraceSlots + k] - t[k]) * 1e-3;
            const double step = (double)(t[2 * kTraceSlots + k] - t[kTraceSlots + k]) * 1e-3;
            const double gap = k + 1 < n ? (double)(t[k + 1] - t[2 * kTraceSlots + k]) * 1e-3 : 0.0;
            const unsigned long long tm = t[3 * kTraceSlots + k];
            const double ustep = tm > t[kTraceSlots + k] && tm <= t[2 * kTraceSlots + k] ? (double)(tm - t[kTraceSlots + k]) * 1e-3 : 0.0;
            if (k < 40 || k >= n - 4) fprintf(stderr, "[lsqr_b200 trace] %4d %10.2f %8.2f %6.2f %6.2f %6.2f\n", k, t0, busy, step, gap, ustep);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// SpMV launch dispatch
// ---------------------------------------------------------------------------------------------
static int pick_lanes(const Csr &M, const char *env_name)
{
    int forced = env_int(env_name, 0);
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return forced;
    const double mean = M.nrows > 0 ? (double)M.nnz / (double)M.nrows : 1.0;
    int lanes = 2;
    while (lanes < 32 && lanes * 3 < mean) lanes <<= 1;   // about three passes over a mean row
    return lanes;
}

static inline CsrView view_of(const Csr &M) { return CsrView{M.ptr, M.idx, M.val, M.nrows}; }
// block b of a row-blocked transpose (an ordinary CSR over nkeys rows whose entries start at ptr[b*nkeys])
static inline CsrView view_of_block(const Csr &M, int64_t b) { return CsrView{M.ptr + b * M.nkeys, M.idx, M.val, M.nkeys}; }

template <int EPI>
static int launch_spmv(Work &wk, const CsrView &V, int lanes, const double *x, double *out, double *aux, int check_done = 0)
{
    const int grid = wk.grid_for(V.nrows, kThreads / lanes);
    switch (lanes) {
    case 1:  spmv_rowgroup_kernel<1, EPI><<<grid, kThreads, 0, wk.stream>>>(V, x, out, wk.st, aux, check_done); break;
    case 2:  spmv_rowgroup_kernel<2, EPI><<<grid, kThreads, 0, wk.stream>>>(V, x, out, wk.st, aux, check_done); break;
    case 4:  spmv_rowgroup_kernel<4, EPI><<<grid, kThreads, 0, wk.stream>>>(V, x, out, wk.st, aux, check_done); break;
    case 8:  spmv_rowgroup_kernel<8, EPI><<<grid, kThreads, 0, wk.stream>>>(V, x, out, wk.st, aux, check_done); break;
    case 16: spmv_rowgroup_kernel<16, EPI><<<grid, kThreads, 0, wk.stream>>>(V, x, out, wk.st, aux, check_done); break;
    default: spmv_rowgroup_kernel<32, EPI><<<grid, kThreads, 0, wk.stream>>>(V, x, out, wk.st, aux, check_done); break;
    }
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

// ---- variant 2: tile-streamed kernels (spmv_stream.cuh) -----------------------------------------
struct TileMapOwner {
    uint2 *tiles = nullptr;
    int ntiles = 0;
    uint32_t tile = kTile;   // nominal work units per tile (stored entries + row_w per row)
    uint32_t row_w = 0;      // kind 3: weight of a row in the tile cut, so that long runs of empty rows are split too
    int kind = 2;            // 2 = CTA tiles of the TMA-streamed kernel, 3 = warp tiles of the segmented kernel
    // kind 3, very uneven rows only: balanced tile schedule (see TileMap::order); the grid is then the full persistent grid
    uint32_t *order = nullptr;
    int nslots = 0;
    int ctas = 0;            // kind 3: persistent grid this map was cut for (0 = sms * CTAs per SM); the chunk launches
                             // that overlap an all-reduce leave a few SMs to the collective's own kernels
    double imbalance = 1.0;  // most loaded warp / mean load under the schedule in use (diagnostic)
};

static void tile_map_free(TileMapOwner &mp)
{
    if (mp.tiles) cudaFree(mp.tiles);
    if (mp.order) cudaFree(mp.order);
    mp.tiles = nullptr; mp.order = nullptr;
}

// Variant 3 sizes its tiles so that every warp of the persistent grid gets the same number of them
// (k tiles of ~8K entries or less each): a small matrix is cut into exactly one tile per warp.
static uint32_t warp_tile_size(int ctas, int64_t nnz)
{
    const int forced = env_int("LSQR_B200_WARP_TILE", 0);
    if (forced >= 128) return (uint32_t)forced;
    const int64_t nwarps = (int64_t)ctas * kWWarps;
    const int64_t k = std::max<int64_t>(1, (nnz + nwarps * 8192 - 1) / (nwarps * 8192));
    const int64_t t = (nnz + nwarps * k - 1) / (nwarps * k);
    return (uint32_t)std::max<int64_t>(512, (t + 3) & ~(int64_t)3);
}

static inline int64_t tile_work(const TileMapOwner &mp, const CsrView &V, int64_t nnz) { return nnz + (int64_t)mp.row_w * V.nrows; }

static int build_tiles(Work &wk, const CsrView &V, int64_t nnz, uint32_t tile, TileMapOwner *out)
{
    if (out->tiles) { cudaFree(out->tiles); out->tiles = nullptr; }
    out->tile = tile;
    const int64_t nt = std::max<int64_t>(1, (tile_work(*out, V, nnz) + tile - 1) / tile);
    out->ntiles = (int)nt;
    LSQRB_CUDA(cudaMalloc(&out->tiles, sizeof(uint2) * (size_t)(nt + 1)));
    build_tiles_kernel<<<(int)((nt + 1 + 255) / 256), 256, 0, wk.stream>>>(V.ptr, V.nrows, (int)nt, tile, out->row_w, out->tiles);
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

// Cost of a warp tile in units of one 128-entry chunk: its chunks plus its row-window reloads.
static inline double tile_cost(const uint2 &a, const uint2 &b)
{
    if (a.x == b.x) return 0.0;   // no row starts here: skipped by the kernel
    return std::ceil((double)(b.y - a.y + 3u) / 128.0) + 0.25 * std::ceil((double)(b.x - a.x) / 32.0) + 1.0;
}

// Tiles are row-aligned, so a row of 10 000 entries makes a tile of more than 10 000: with round-robin assignment the
// most loaded warp of a power-law matrix (C4) carries ~1.5x the mean and the whole grid waits for it.  When that
// happens the matrix is re-cut into smaller tiles and the tiles are dealt to the warps by LPT (largest first, to the
// least loaded warp).  The schedule is a pure function of ptr[], so results stay reproducible run to run.
static int balance_tile_map(Work &wk, const CsrView &V, int64_t nnz, TileMapOwner *out)
{
    const int nw = out->ctas * kWWarps;
    std::vector<uint2> t;
    std::vector<double> cost;
    double total = 0.0;
    auto fetch = [&]() -> int {   // tile bounds and costs of the current cut
        t.resize((size_t)out->ntiles + 1);
        LSQRB_CUDA(cudaMemcpyAsync(t.data(), out->tiles, sizeof(uint2) * t.size(), cudaMemcpyDeviceToHost, wk.stream));
        LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
        cost.resize((size_t)out->ntiles);
        total = 0.0;
        for (int i = 0; i < out->ntiles; ++i) total += (cost[(size_t)i] = tile_cost(t[(size_t)i], t[(size_t)i + 1]));
        return LSQR_B200_OK;
    };
    LSQRB_TRY(fetch());
    {   // round robin: tile i belongs to warp i mod (warps of the grid)
        const int gw = std::max(1, std::min((out->ntiles + kWWarps - 1) / kWWarps, out->ctas)) * kWWarps;
        std::vector<double> load((size_t)gw, 0.0);
        for (int i = 0; i < out->ntiles; ++i) load[(size_t)(i % gw)] += cost[(size_t)i];
        out->imbalance = total > 0 ? *std::max_element(load.begin(), load.end()) / (total / nw) : 1.0;
    }
    if (env_int("LSQR_B200_BALANCE", 1) == 0 || tile_work(*out, V, nnz) <= (int64_t)nw * 512) return LSQR_B200_OK;   // nothing to deal out
    if (out->imbalance <= 1.0 + 1e-3 * env_int("LSQR_B200_BALANCE_PERMILLE", 60)) return LSQR_B200_OK;

    // finer tiles pack better (the long rows stay as long as they are)
    const uint32_t fine = (uint32_t)std::max(512, std::min<int>((int)out->tile / 4, env_int("LSQR_B200_BALANCE_TILE", 2048))) & ~3u;
    if (fine < out->tile) { LSQRB_TRY(build_tiles(wk, V, nnz, fine, out)); LSQRB_TRY(fetch()); }
    const int nt = out->ntiles;
    if (nt <= nw) return LSQR_B200_OK;
    std::vector<int> ids;
    ids.reserve((size_t)nt);
    for (int i = 0; i < nt; ++i) if (cost[(size_t)i] > 0.0) ids.push_back(i);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return cost[(size_t)a] > cost[(size_t)b]; });
    typedef std::pair<double, int> Slot;   // (load, warp): the least loaded warp first, ties by warp number
    std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> heap;
    for (int w = 0; w < nw; ++w) heap.push(Slot(0.0, w));
    std::vector<std::vector<uint32_t>> lists((size_t)nw);
    for (int id : ids) {
        Slot sl = heap.top(); heap.pop();
        lists[(size_t)sl.second].push_back((uint32_t)id);
        sl.first += cost[(size_t)id];
        heap.push(sl);
    }
    size_t depth = 0;
    double worst = 0.0;
    while (!heap.empty()) { worst = std::max(worst, heap.top().first); heap.pop(); }
    for (auto &l : lists) { std::sort(l.begin(), l.end()); depth = std::max(depth, l.size()); }   // each warp walks its tiles in matrix order
    std::vector<uint32_t> order(depth * (size_t)nw, kNoTile);
    for (int w = 0; w < nw; ++w)
        for (size_t k = 0; k < lists[(size_t)w].size(); ++k) order[k * (size_t)nw + (size_t)w] = lists[(size_t)w][k];
    out->nslots = (int)order.size();
    out->imbalance = total > 0 ? worst / (total / nw) : 1.0;
    LSQRB_CUDA(cudaMalloc(&out->order, sizeof(uint32_t) * std::max<size_t>(order.size(), 1)));
    LSQRB_CUDA(cudaMemcpyAsync(out->order, order.data(), sizeof(uint32_t) * order.size(), cudaMemcpyHostToDevice, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

// V: a whole CSR, or one block of a row-blocked transpose (nnz = its number of stored entries)
static int build_tile_map(Work &wk, const CsrView &V, int64_t nnz, int kind, TileMapOwner *out, int reserve_sms = 0)
{
    out->kind = kind;
    out->ctas = std::max(1, wk.sms - reserve_sms) * kWMinBlocks;
    // A row costs about as much as 4 stored entries (its share of a window reload and its epilogue).  Without the row
    // term a block of a row-blocked banded A' -- half of whose rows are empty -- puts a million empty rows into ONE
    // tile, i.e. one warp (measured: C3 full size, Atprod 55 ms instead of 1.4 ms).
    out->row_w = kind == 3 ? (uint32_t)std::max(0, env_int("LSQR_B200_TILE_ROW_WEIGHT", 4)) : 0u;
    LSQRB_TRY(build_tiles(wk, V, nnz, kind == 3 ? warp_tile_size(out->ctas, tile_work(*out, V, nnz)) : (uint32_t)kTile, out));
    if (kind == 3) LSQRB_TRY(balance_tile_map(wk, V, nnz, out));
    return LSQR_B200_OK;
}

// Opt in to the large dynamic shared memory window and the maximum carve-out, and measure how many
// CTAs of this instantiation are co-resident per SM: the persistent grid is sized to exactly that.
template <int EPI>
static int stream_kernel_prepare(int *ctas_per_sm)
{
    static int occ = 0;
    if (occ == 0) {
        LSQRB_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamSmem));
        LSQRB_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<EPI>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int n = 0;
        LSQRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmv_stream_kernel<EPI>, kStreamThreads, kStreamSmem));
        occ = std::max(1, n);
        if (env_int("LSQR_B200_VERBOSE", 0))
            fprintf(stderr, "[lsqr_b200] spmv_stream_kernel<%d>: %d CTAs/SM, %zu B dynamic smem\n", EPI, occ, kStreamSmem);
    }
    *ctas_per_sm = occ;
    return LSQR_B200_OK;
}

struct StreamExtra {   // operands of the fused deferred update
    double *ux = nullptr, *uw = nullptr, *use = nullptr;
    int check_done = 0;
};

template <int EPI>
static int launch_stream(Work &wk, const CsrView &V, const TileMapOwner &map, const double *x, double *out, double *aux,
                         const StreamExtra &ex = StreamExtra())
{
    StreamArgs a;
    a.A = V;
    a.map = TileMap{map.tiles, map.ntiles, map.order, map.nslots};
    a.x = x; a.out = out; a.st = wk.st; a.aux = aux;
    a.ux = ex.ux; a.uw = ex.uw; a.use = ex.use;
    a.ring = wk.ring_d;
    a.out_aligned16 = ((uintptr_t)out & 15u) == 0;
    a.check_done = ex.check_done;
    if (map.kind == 3) {
        const int ctas = map.ctas > 0 ? map.ctas : wk.sms * kWMinBlocks;
        const int grid = map.order ? ctas : std::max(1, std::min((map.ntiles + kWWarps - 1) / kWWarps, ctas));
        spmv_warp_kernel<EPI><<<grid, kWThreads, 0, wk.stream>>>(a);
    } else {
        int occ = 1;
        LSQRB_TRY(stream_kernel_prepare<EPI>(&occ));
        const int grid = std::max(1, std::min(map.ntiles, std::min(wk.stream_grid, wk.sms * occ)));
        spmv_stream_kernel<EPI><<<grid, kStreamThreads, kStreamSmem, wk.stream>>>(a);
    }
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

static inline int vec_ok(const void *a, const void *b, const void *c, const void *d)
{
    auto al = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
    return al(a) && al(b) && al(c) && al(d);
}

template <bool LAZY>
static int launch_update(Work &wk, int64_t n, double *x, double *w, const double *v, double *se, bool wantse)
{
    const int grid = wk.grid_for((n + 1) / 2, kThreads);
    const int vok = vec_ok(x, w, v, wantse ? se : nullptr);
    if (wantse) xw_update_kernel<true, LAZY><<<grid, kThreads, 0, wk.stream>>>(n, x, w, v, se, wk.st, wk.ring_d, vok);
    else        xw_update_kernel<false, LAZY><<<grid, kThreads, 0, wk.stream>>>(n, x, w, v, se, wk.st, wk.ring_d, vok);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

// record 0 of the log (src/lsqr.f90:666-671): written once the first alpha, beta are known
__global__ void record0_kernel(DevState *st, volatile lsqr_b200_iter_record *ring)
{
    volatile lsqr_b200_iter_record *r = ring;
    r->istop = (double)st->istop;
    r->x1 = 0.0;
    r->rnorm = st->rnorm;
    r->test1 = 1.0;
    r->test2 = st->beta > 0.0 ? st->alpha / st->beta : 0.0;
    r->anorm = 0.0; r->acond = 0.0; r->phi = 0.0; r->dknorm = 0.0; r->dxk = 0.0; r->alfopt = 0.0;
    r->alpha = st->alpha; r->beta = st->beta; r->xnorm = 0.0; r->arnorm = st->arnorm;
    __threadfence_system();
    r->itn = 0.0;
}

// ---------------------------------------------------------------------------------------------
// nout log reconstruction on the host (formats of src/lsqr.f90:589-595,655-671,827-829,872-880)
// ---------------------------------------------------------------------------------------------
struct LogCtx {
    lsqr_b200_log_fn log = nullptr;  void *log_user = nullptr;
    lsqr_b200_iter_fn iter = nullptr; void *iter_user = nullptr;
    int64_t m = 0, n = 0;
    double damp = 0, atol = 0, btol = 0, conlim = 0, ctol = 0;
    int itnlim = 0, wantse = 0;
    double bnorm = 0;

    void line(const std::string &s) const { if (log) log(log_user, s.c_str()); }

    void header() const
    {
        if (!log) return;
        char buf[160];
        line(""); line("");
        line(" Enter LSQR.       Least-squares solution of  Ax = b");
        snprintf(buf, sizeof buf, " The matrix  A  has%7lld rows   and%7lld columns", (long long)m, (long long)n);
        line(buf);
        line(" damp   =" + fe(22, 14, damp) + "   wantse =" + std::string(9, ' ') + (wantse ? "T" : "F"));
        line(" atol   =" + fe(10, 2, atol) + std::string(15, ' ') + "conlim =" + fe(10, 2, conlim));
        snprintf(buf, sizeof buf, "%10d", itnlim);
        line(" btol   =" + fe(10, 2, btol) + std::string(15, ' ') + "itnlim =" + buf);
    }

    void iter_line(const lsqr_b200_iter_record &r, int nvals) const
    {
        static const int w[10] = {17, 17, 10, 10, 10, 10, 9, 8, 8, 8};
        static const int d[10] = {9, 9, 2, 2, 2, 2, 1, 1, 1, 1};
        const double vals[10] = {r.x1, r.rnorm, r.test1, r.test2, r.anorm, r.acond, r.phi, r.dknorm, r.dxk, r.alfopt};
        char buf[16];
        snprintf(buf, sizeof buf, "%6d", (int)r.itn);
        std::string s(buf);
        for (int k = 0; k < nvals; ++k) s += fe(w[k], d[k], vals[k]);
        line(s);
    }

    void record(const lsqr_b200_iter_record &r)
    {
        if (iter) iter(iter_user, &r);
        if (!log) return;
        const int itn = (int)r.itn;
        if (itn == 0) {
            bnorm = r.beta;
            line(""); line("");
            if (damp > 0.0) line("   Itn       x(1)           Function     Compatible   LS     Norm Abar Cond Abar");
            else            line("   Itn       x(1)           Function     Compatible   LS        Norm A    Cond A");
            line(std::string(80, ' ') + "    phi    dknorm   dxk  alfa_opt");
            iter_line(r, 4);
            line("");
            return;
        }
        const double test3 = 1.0 / r.acond;
        const double rtol = btol + atol * r.anorm * r.xnorm / bnorm;
        const bool print_iter = (n <= 40) || (itn <= 10) || (itn >= itnlim - 10) || (itn % 10 == 0) ||
                                (test3 <= 2.0 * ctol) || (r.test2 <= 10.0 * atol) ||
                                (r.test1 <= 10.0 * rtol) || (r.istop != 0.0);
        if (print_iter) iter_line(r, 10);
    }

    void footer(int istop, const DevState &h) const
    {
        if (!log) return;
        static const char *const msg[6] = {
            "The exact solution is x = 0                          ",
            "A solution to Ax = b was found, given atol, btol     ",
            "A least-squares solution was found, given atol       ",
            "A damped least-squares solution was found, given atol",
            "Cond(Abar) seems to be too large, given conlim       ",
            "The iteration limit was reached                      "};
        char buf[160];
        const std::string ex = " Exit  LSQR.  ";
        line(""); line("");
        snprintf(buf, sizeof buf, "     istop  =%2d               itn    =%8d", istop, h.itn);
        line(ex + buf);
        line(ex + "     anorm  =" + fe(12, 5, h.anorm) + "     acond  =" + fe(12, 5, h.acond));
        line(ex + "     bnorm  =" + fe(12, 5, h.bnorm) + "     xnorm  =" + fe(12, 5, h.xnorm));
        line(ex + "     rnorm  =" + fe(12, 5, h.rnorm) + "     arnorm =" + fe(12, 5, h.arnorm));
        snprintf(buf, sizeof buf, " occurred at itn %8d", h.maxdx);
        line(ex + "     max dx =" + fe(8, 1, h.dxmax) + buf);
        line(ex + "            =" + fe(8, 1, h.dxmax / (h.xnorm + 1.0e-20)) + "*xnorm");
        line(ex + "     " + msg[istop]);
    }
};

// Consume finished records in order; returns true once a record carries istop != 0.
static bool drain_ring(Work &wk, LogCtx &lc, int &seen, int upto)
{
    bool stop = false;
    while (seen < upto) {
        const int k = seen + 1;
        volatile lsqr_b200_iter_record *r = wk.ring_h + (k % kRingSize);
        if (r->itn != (double)k) break;
        lsqr_b200_iter_record rec;
        memcpy(&rec, (const void *)r, sizeof rec);
        lc.record(rec);
        seen = k;
        if (rec.istop != 0.0) { stop = true; break; }
    }
    return stop;
}

}  // namespace lsqrb

using namespace lsqrb;

// =============================================================================================
// the ez handle
// =============================================================================================
struct lsqr_b200_ez {
    Work wk;
    int32_t m = 0, n = 0;
    int64_t nnz = 0;
    Csr A, AT;
    int lanes_a = 4, lanes_at = 32;
    std::vector<TileMapOwner> mapA;    // one per block of the column-blocked A
    std::vector<TileMapOwner> mapAT;   // one per block of the row-blocked transpose
    std::vector<int64_t> a_off;        // first stored entry of every block of A (nblocks + 1 values)
    // multi-GPU: the LAST row block of A' is cut into comm_chunks column ranges; the all-reduce of range c runs on
    // comm_stream while the SpMV launch of range c+1 computes (mapATc[c], chunk c = columns [cc[c], cc[c+1]))
    int comm_chunks = 1;
    std::vector<int64_t> cc;
    std::vector<TileMapOwner> mapATc;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_comm = nullptr;
    bool fuse_last = false;            // column-blocked A, tiled kernels: the last block's launch finishes the Aprod step
    bool a_blocked = false;            // A is column-blocked (v does not fit in L2): Aprod = one launch per block into gu
    double *gu = nullptr;              // [m] A v of the column-blocked Aprod
    std::vector<int64_t> at_off;       // first stored entry of every block of A' (nblocks + 1 values)
    bool stream = true;           // tiled kernels (variants 2, 3) vs sub-warp-per-row (variant 1)
    bool blocked = false;         // A' is row-blocked (u does not fit in L2): Atprod = one launch per block into g
    bool deferred = false;        // single GPU, tiled, unblocked: x/w update fused into the next Atprod (2 kernels / iteration)
    bool overlap_update = true;   // fused engine, 3 kernels / iteration: the x/w update of iteration k runs on a side
                                  // stream next to the Aprod of iteration k+1 (they touch disjoint vectors)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false;    // an update is in flight on the side stream that the main stream has not joined
    lsqr_b200_options opt;
    double *u = nullptr, *v = nullptr, *w = nullptr, *x = nullptr, *se = nullptr;
    double *g = nullptr;          // multi-GPU: [ A_p'u_p (n) | sum(u_p^2) ]
    double *tmp_m = nullptr, *tmp_n = nullptr;   // staging for lsqr_b200_ez_aprod with host vectors
    ncclComm_t comm = nullptr;
    int batch = 8;                // iterations per enqueue (per CUDA-graph launch)
    cudaGraphExec_t graph_exec = nullptr;
    int graph_wantse = -1;
    lsqr_b200_kernel_times times;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_t2 = nullptr;
    std::vector<cudaEvent_t> prof_ev;   // profile mode: start/stop pairs
    std::vector<int> prof_cls;
};

static void ez_free(lsqr_b200_ez *me)
{
    if (!me) return;
    cudaSetDevice(me->wk.device);
    if (me->wk.stream) cudaStreamSynchronize(me->wk.stream);
    if (me->graph_exec) cudaGraphExecDestroy(me->graph_exec);
    if (me->comm) { NcclApi *a = nccl_api(); if (a) a->CommDestroy(me->comm); }
    csr_free(&me->A);
    csr_free(&me->AT);
    for (auto &mp : me->mapA) tile_map_free(mp);
    for (auto &mp : me->mapAT) tile_map_free(mp);
    for (auto &mp : me->mapATc) tile_map_free(mp);
    for (auto e : me->ev_chunk) if (e) cudaEventDestroy(e);
    if (me->ev_comm) cudaEventDestroy(me->ev_comm);
    if (me->comm_stream) cudaStreamDestroy(me->comm_stream);
    if (me->gu) cudaFree(me->gu);
    for (double *p : {me->u, me->v, me->w, me->x, me->se, me->g, me->tmp_m, me->tmp_n}) if (p) cudaFree(p);
    for (auto e : {me->ev_t0, me->ev_t1, me->ev_t2, me->ev_fork, me->ev_join}) if (e) cudaEventDestroy(e);
    if (me->side) cudaStreamDestroy(me->side);
    for (auto e : me->prof_ev) cudaEventDestroy(e);
    me->wk.destroy();
    delete me;
}

extern "C" {

const char *lsqr_b200_error_message(int code)
{
    switch (code) {
    case LSQR_B200_OK:            return "";
    case LSQR_B200_ERR_SIZES:     return "invalid a,icol,irow sizes in initialize_ez";
    case LSQR_B200_ERR_IROW:      return "invalid irow or m in initialize_ez";
    case LSQR_B200_ERR_ICOL:      return "invalid icol or n in initialize_ez";
    case LSQR_B200_ERR_NOINIT:    return "lsqr_solver_ez class not properly initialized";
    case LSQR_B200_ERR_MODE:      return "invalid mode in aprod_ez";
    case LSQR_B200_ERR_INDEX_LOW: return "irow or icol contains an index smaller than 1";
    case LSQR_B200_ERR_NO_DEVICE: return "no CUDA device available (the engine has no CPU fallback)";
    case LSQR_B200_ERR_CUDA:      return "CUDA runtime error";
    case LSQR_B200_ERR_NCCL:      return "NCCL error";
    case LSQR_B200_ERR_ARG:       return "invalid argument";
    case LSQR_B200_ERR_TOO_LARGE: return "too many stored entries for one GPU (limit 2^32-2)";
    case LSQR_B200_ERR_CALLBACK:  return "user aprod callback failed";
    default:                      return "unknown error code";
    }
}

const char *lsqr_b200_last_error(void) { return g_last_error.c_str(); }
int lsqr_b200_version(void) { return LSQR_B200_VERSION; }

int lsqr_b200_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

void lsqr_b200_default_options(lsqr_b200_options *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->itnlim = 100;        // src/lsqr.f90:50
    o->device = -1;
    o->use_graph = 1;
    o->world_size = 1;
}

int lsqr_b200_nccl_unique_id(void *out128)
{
    if (!out128) return LSQR_B200_ERR_ARG;
    NcclApi *a = nccl_api();
    if (!a) { set_last_error("libnccl.so.2 could not be loaded"); return LSQR_B200_ERR_NCCL; }
    ncclUniqueId id;
    LSQRB_NCCL(a->GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return LSQR_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// initialize_ez  (src/lsqr.f90:91-127)
// ---------------------------------------------------------------------------------------------
static int ez_initialize_impl(lsqr_b200_ez *me, int64_t nnz, const double *a, const int32_t *irow, const int32_t *icol)
{
    Work &wk = me->wk;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    // deep copy of the triplets onto the GPU (src/lsqr.f90:113-118); host or device sources
    int32_t *d_irow = nullptr, *d_icol = nullptr;
    double *d_a = nullptr;
    const bool dev_src = is_device_ptr(a) && is_device_ptr(irow) && is_device_ptr(icol);
    if (dev_src) {
        d_irow = const_cast<int32_t *>(irow);
        d_icol = const_cast<int32_t *>(icol);
        d_a = const_cast<double *>(a);
    } else {
        LSQRB_CUDA(cudaMalloc(&d_irow, sizeof(int32_t) * nz));
        LSQRB_CUDA(cudaMalloc(&d_icol, sizeof(int32_t) * nz));
        LSQRB_CUDA(cudaMalloc(&d_a, sizeof(double) * nz));
        if (nnz > 0) {
            LSQRB_CUDA(cudaMemcpyAsync(d_irow, irow, sizeof(int32_t) * nz, cudaMemcpyDefault, wk.stream));
            LSQRB_CUDA(cudaMemcpyAsync(d_icol, icol, sizeof(int32_t) * nz, cudaMemcpyDefault, wk.stream));
            LSQRB_CUDA(cudaMemcpyAsync(d_a, a, sizeof(double) * nz, cudaMemcpyDefault, wk.stream));
        }
    }
    auto release = [&]() {
        if (!dev_src) { cudaFree(d_irow); cudaFree(d_icol); cudaFree(d_a); }
    };
    int rc = coo_validate(wk.stream, me->m, me->n, nnz, d_irow, d_icol);
    if (rc == LSQR_B200_OK) {
        // Column-blocked A: the mirror image of the row-blocked transpose below, for a v (8 n bytes) beyond the L2 budget
        int64_t block_cols = env_int("LSQR_B200_VBLOCK_COLS", 0);
        if (block_cols <= 0) {
            const int64_t budget = (int64_t)env_int("LSQR_B200_VBLOCK_MB", 48) * (1 << 20) / 8;
            if (me->n > budget) {
                const int64_t nb = (me->n + budget - 1) / budget;
                block_cols = (me->n + nb - 1) / nb;
            }
        }
        rc = coo_to_csr_device(wk.stream, me->m, nnz, d_irow, d_icol, d_a, &me->A, block_cols, me->n);
    }
    if (rc == LSQR_B200_OK) {
        // Row-blocked transpose: when u (8 m bytes) cannot stay in L2 while A' streams past it, A' is stored as one
        // CSR per block of rows of A, so that every Atprod launch gathers from a slice of u that does fit.
        int64_t block_rows = env_int("LSQR_B200_UBLOCK_ROWS", 0);
        if (block_rows <= 0) {
            const int64_t budget_rows = (int64_t)env_int("LSQR_B200_UBLOCK_MB", 48) * (1 << 20) / 8;
            if (me->m > budget_rows) {
                const int64_t nb = (me->m + budget_rows - 1) / budget_rows;
                block_rows = (me->m + nb - 1) / nb;
            }
        }
        rc = coo_to_csr_device(wk.stream, me->n, nnz, d_icol, d_irow, d_a, &me->AT, block_rows, me->m);
    }
    if (rc == LSQR_B200_OK && cudaStreamSynchronize(wk.stream) != cudaSuccess) rc = LSQR_B200_ERR_CUDA;
    release();
    LSQRB_TRY(rc);

    me->lanes_a = pick_lanes(me->A, "LSQR_B200_LANES_A");
    me->lanes_at = pick_lanes(me->AT, "LSQR_B200_LANES_AT");
    const int variant = me->opt.spmv_variant ? me->opt.spmv_variant : env_int("LSQR_B200_SPMV_VARIANT", 3);
    me->stream = variant != 1;
    me->blocked = me->AT.nblocks > 1;
    me->a_blocked = me->A.nblocks > 1;
    me->fuse_last = me->a_blocked && me->stream && env_int("LSQR_B200_FUSE_LAST_BLOCK", 1) != 0;
    me->deferred = me->stream && !me->blocked && !me->a_blocked && me->opt.world_size == 1 && env_int("LSQR_B200_DEFERRED_UPDATE", 0) != 0;
    me->overlap_update = !me->deferred && !me->blocked && !me->a_blocked && me->opt.world_size == 1 && env_int("LSQR_B200_OVERLAP_UPDATE", 1) != 0;
    if (me->overlap_update) {
        LSQRB_CUDA(cudaStreamCreateWithFlags(&me->side, cudaStreamNonBlocking));
        LSQRB_CUDA(cudaEventCreateWithFlags(&me->ev_fork, cudaEventDisableTiming));
        LSQRB_CUDA(cudaEventCreateWithFlags(&me->ev_join, cudaEventDisableTiming));
    }
    // first stored entry of every block of A and of A'
    auto block_offsets = [&](const Csr &M, std::vector<int64_t> *out) -> int {
        const int64_t nb = M.nblocks;
        out->assign((size_t)nb + 1, 0);
        std::vector<uint32_t> off((size_t)nb + 1);
        for (int64_t b = 0; b <= nb; ++b)
            LSQRB_CUDA(cudaMemcpyAsync(&off[(size_t)b], M.ptr + b * M.nkeys, sizeof(uint32_t), cudaMemcpyDeviceToHost, wk.stream));
        LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
        for (int64_t b = 0; b <= nb; ++b) (*out)[(size_t)b] = off[(size_t)b];
        return LSQR_B200_OK;
    };
    LSQRB_TRY(block_offsets(me->A, &me->a_off));
    LSQRB_TRY(block_offsets(me->AT, &me->at_off));
    if (me->stream) {
        me->mapA.resize((size_t)me->A.nblocks);
        for (int64_t b = 0; b < me->A.nblocks; ++b)
            LSQRB_TRY(build_tile_map(wk, view_of_block(me->A, b), me->a_off[(size_t)b + 1] - me->a_off[(size_t)b],
                                     variant == 2 ? 2 : 3, &me->mapA[(size_t)b]));
        me->mapAT.resize((size_t)me->AT.nblocks);
        for (int64_t b = 0; b < me->AT.nblocks; ++b)
            LSQRB_TRY(build_tile_map(wk, view_of_block(me->AT, b), me->at_off[(size_t)b + 1] - me->at_off[(size_t)b],
                                     variant == 2 ? 2 : 3, &me->mapAT[(size_t)b]));
        LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    }
    if (me->a_blocked) LSQRB_CUDA(cudaMalloc(&me->gu, sizeof(double) * (size_t)std::max<int32_t>(me->m, 1)));
    if (me->opt.world_size > 1 && me->stream) {
        // Column chunks for the pipelined all-reduce.  Off by default: measured on 8 x B200 (C5, n = 1e7, 80 MB per
        // all-reduce) 1 / 4 / 8 chunks give 3.05 / 3.14 / 3.26 ms per iteration -- the persistent SpMV grids hold every
        // SM's registers, so NCCL's CTAs only start when a launch drains and the smaller launches pay more in tails
        // than the overlap returns (profiles/r01/run2/n8_allreduce_pipelining_ab.txt).
        int k = env_int("LSQR_B200_COMM_CHUNKS", 1);
        k = std::max(1, std::min(k, 8));
        me->comm_chunks = k;
        if (k > 1) {
            me->cc.assign((size_t)k + 1, 0);
            for (int c = 0; c <= k; ++c) me->cc[(size_t)c] = c == k ? (int64_t)me->n : (((int64_t)me->n * c / k) & ~(int64_t)1);
            const int64_t bl = me->AT.nblocks - 1;          // only the last block is pipelined (no extra passes over u)
            std::vector<uint32_t> off((size_t)k + 1);
            for (int c = 0; c <= k; ++c)
                LSQRB_CUDA(cudaMemcpyAsync(&off[(size_t)c], me->AT.ptr + bl * me->AT.nkeys + me->cc[(size_t)c],
                                           sizeof(uint32_t), cudaMemcpyDeviceToHost, wk.stream));
            LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
            me->mapATc.resize((size_t)k);
            for (int c = 0; c < k; ++c) {
                CsrView V = view_of_block(me->AT, bl);
                V.ptr += me->cc[(size_t)c];
                V.nrows = me->cc[(size_t)c + 1] - me->cc[(size_t)c];
                // every range but the first runs next to the all-reduce of the previous one: leave NCCL some SMs
                LSQRB_TRY(build_tile_map(wk, V, (int64_t)off[(size_t)c + 1] - (int64_t)off[(size_t)c], variant == 2 ? 2 : 3,
                                         &me->mapATc[(size_t)c], c > 0 ? env_int("LSQR_B200_COMM_RESERVE_SMS", 0) : 0));
            }
            LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
            // highest priority: the SpMV grids are persistent and fill every SM, so the collective's CTAs can only
            // start on resources freed by a finishing SpMV launch -- they must win those against the next launch
            int prio_lo = 0, prio_hi = 0;
            LSQRB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            LSQRB_CUDA(cudaStreamCreateWithPriority(&me->comm_stream, cudaStreamNonBlocking,
                                                    env_int("LSQR_B200_COMM_PRIORITY", 1) ? prio_hi : prio_lo));
            for (int c = 0; c < k; ++c) LSQRB_CUDA(cudaEventCreateWithFlags(&me->ev_chunk[c], cudaEventDisableTiming));
            LSQRB_CUDA(cudaEventCreateWithFlags(&me->ev_comm, cudaEventDisableTiming));
        }
    }
    if (env_int("LSQR_B200_VERBOSE", 0))
        fprintf(stderr, "[lsqr_b200] m=%d n=%d nnz=%lld variant=%d A blocks=%lld (block_cols=%lld) A' blocks=%lld (block_rows=%lld) "
                        "warp tile A=%u%s (imbalance %.3f) A'=%u%s (imbalance %.3f)\n", me->m, me->n, (long long)me->nnz, variant,
                (long long)me->A.nblocks, (long long)me->A.block_rows, (long long)me->AT.nblocks, (long long)me->AT.block_rows,
                me->mapA.empty() ? 0u : me->mapA[0].tile, !me->mapA.empty() && me->mapA[0].order ? " LPT" : "",
                me->mapA.empty() ? 1.0 : me->mapA[0].imbalance,
                me->mapAT.empty() ? 0u : me->mapAT[0].tile, !me->mapAT.empty() && me->mapAT[0].order ? " LPT" : "",
                me->mapAT.empty() ? 1.0 : me->mapAT[0].imbalance);

    const size_t mm = (size_t)std::max<int32_t>(me->m, 1), nn = (size_t)std::max<int32_t>(me->n, 1);
    LSQRB_CUDA(cudaMalloc(&me->u, sizeof(double) * mm));
    LSQRB_CUDA(cudaMalloc(&me->v, sizeof(double) * nn));
    LSQRB_CUDA(cudaMalloc(&me->w, sizeof(double) * nn));
    LSQRB_CUDA(cudaMalloc(&me->x, sizeof(double) * nn));
    if (me->opt.world_size > 1 || me->blocked) LSQRB_CUDA(cudaMalloc(&me->g, sizeof(double) * (nn + 1)));
    LSQRB_CUDA(cudaEventCreate(&me->ev_t0));
    LSQRB_CUDA(cudaEventCreate(&me->ev_t1));
    LSQRB_CUDA(cudaEventCreate(&me->ev_t2));
    return LSQR_B200_OK;
}

int lsqr_b200_ez_initialize(lsqr_b200_ez **out, int32_t m, int32_t n,
                            int64_t size_a, const double *a,
                            int64_t size_irow, const int32_t *irow,
                            int64_t size_icol, const int32_t *icol,
                            const lsqr_b200_options *opts)
{
    if (!out) return LSQR_B200_ERR_ARG;
    *out = nullptr;
    // src/lsqr.f90:109 -- the three sizes must agree
    if (size_a != size_irow || size_a != size_icol) return LSQR_B200_ERR_SIZES;
    if (m < 0 || n < 0 || size_a < 0) { set_last_error("negative dimension"); return LSQR_B200_ERR_ARG; }
    if (size_a > 0 && (!a || !irow || !icol)) { set_last_error("NULL triplet array"); return LSQR_B200_ERR_ARG; }
    if (size_a > (int64_t)0xFFFFFFFEll) return LSQR_B200_ERR_TOO_LARGE;

    lsqr_b200_ez *me = new lsqr_b200_ez();
    if (opts) me->opt = *opts; else lsqr_b200_default_options(&me->opt);
    if (me->opt.world_size < 1) me->opt.world_size = 1;
    me->m = m; me->n = n; me->nnz = size_a;
    memset(&me->times, 0, sizeof me->times);
    {   // iterations per enqueue: enough to cover ~300 us of device time, so that the host's per-batch work (graph
        // launch, event wait, record drain) stays hidden, but not more: iterations enqueued past the stop are waste
        const double est_iter_us = 24.0 * (double)size_a / 2.5e6 + 20.0;   // ~2.5 TB/s effective + launch floor
        int dflt = (int)std::ceil(300.0 / est_iter_us);
        dflt = std::max(1, std::min(dflt, 8));
        me->batch = std::max(1, std::min(env_int("LSQR_B200_BATCH", dflt), kRingSize / 4));
    }

    int rc = me->wk.init(me->opt.device, me->opt.stream);
    if (rc == LSQR_B200_OK && me->opt.world_size > 1) {
        NcclApi *api = nccl_api();
        if (!api) { set_last_error("libnccl.so.2 could not be loaded"); rc = LSQR_B200_ERR_NCCL; }
        else if (!me->opt.nccl_unique_id) { set_last_error("world_size > 1 needs nccl_unique_id"); rc = LSQR_B200_ERR_ARG; }
        else {
            ncclUniqueId id;
            memcpy(&id, me->opt.nccl_unique_id, 128);
            ncclResult_t r = api->CommInitRank(&me->comm, me->opt.world_size, id, me->opt.rank);
            if (r != ncclSuccess) { set_last_error("ncclCommInitRank failed"); rc = LSQR_B200_ERR_NCCL; }
        }
        me->opt.nccl_unique_id = nullptr;   // the caller's buffer need not outlive this call
    }
    if (rc == LSQR_B200_OK) rc = ez_initialize_impl(me, size_a, a, irow, icol);
    if (rc != LSQR_B200_OK) { ez_free(me); return rc; }
    *out = me;
    return LSQR_B200_OK;
}

void lsqr_b200_ez_destroy(lsqr_b200_ez *me) { ez_free(me); }

int lsqr_b200_ez_set_options(lsqr_b200_ez *me, const lsqr_b200_options *o)
{
    if (!me || !o) return LSQR_B200_ERR_ARG;
    me->opt.atol = o->atol; me->opt.btol = o->btol; me->opt.conlim = o->conlim; me->opt.itnlim = o->itnlim;
    me->opt.log = o->log; me->opt.log_user = o->log_user; me->opt.iter = o->iter; me->opt.iter_user = o->iter_user;
    me->opt.engine = o->engine; me->opt.use_graph = o->use_graph; me->opt.profile = o->profile;
    return LSQR_B200_OK;
}

int lsqr_b200_ez_get_csr_device(lsqr_b200_ez *me, int32_t which, const uint32_t **ptr_dev, const int32_t **idx_dev,
                                const double **val_dev, const uint32_t **perm_dev)
{
    if (!me || (which != 0 && which != 1)) return LSQR_B200_ERR_ARG;
    const Csr &M = which == 0 ? me->A : me->AT;
    if (ptr_dev) *ptr_dev = M.ptr;
    if (idx_dev) *idx_dev = M.idx;
    if (val_dev) *val_dev = M.val;
    if (perm_dev) *perm_dev = M.perm;
    return LSQR_B200_OK;
}

int64_t lsqr_b200_ez_nnz(const lsqr_b200_ez *me) { return me ? me->nnz : -1; }

int lsqr_b200_ez_blocks(const lsqr_b200_ez *me, int32_t which, int64_t *nblocks, int64_t *block_size)
{
    if (!me || (which != 0 && which != 1)) return LSQR_B200_ERR_ARG;
    const Csr &M = which == 0 ? me->A : me->AT;
    if (nblocks) *nblocks = M.nblocks;
    if (block_size) *block_size = M.block_rows;
    return LSQR_B200_OK;
}

int lsqr_b200_ez_schedule(const lsqr_b200_ez *me, int32_t which, int64_t block, int64_t *ntiles,
                          int64_t *tile_entries, int32_t *balanced, double *imbalance)
{
    if (!me) return LSQR_B200_ERR_ARG;
    const std::vector<TileMapOwner> &maps = which ? me->mapAT : me->mapA;
    if (block < 0 || (size_t)block >= maps.size()) { set_last_error("no such block / no tiled schedule"); return LSQR_B200_ERR_ARG; }
    const TileMapOwner &mp = maps[(size_t)block];
    if (ntiles) *ntiles = mp.ntiles;
    if (tile_entries) *tile_entries = mp.tile;
    if (balanced) *balanced = mp.order != nullptr;
    if (imbalance) *imbalance = mp.imbalance;
    return LSQR_B200_OK;
}

int lsqr_b200_ez_get_csr(lsqr_b200_ez *me, int32_t which, int64_t *ptr, int32_t *idx, double *val, int64_t *perm)
{
    if (!me || (which != 0 && which != 1)) return LSQR_B200_ERR_ARG;
    LSQRB_CUDA(cudaSetDevice(me->wk.device));
    const Csr &M = which == 0 ? me->A : me->AT;
    if (ptr) {
        std::vector<uint32_t> p((size_t)M.nrows + 1);
        LSQRB_CUDA(cudaMemcpy(p.data(), M.ptr, sizeof(uint32_t) * p.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < p.size(); ++i) ptr[i] = (int64_t)p[i];
    }
    if (M.nnz > 0) {
        if (idx) LSQRB_CUDA(cudaMemcpy(idx, M.idx, sizeof(int32_t) * (size_t)M.nnz, cudaMemcpyDeviceToHost));
        if (val) LSQRB_CUDA(cudaMemcpy(val, M.val, sizeof(double) * (size_t)M.nnz, cudaMemcpyDeviceToHost));
        if (perm) {
            std::vector<uint32_t> p((size_t)M.nnz);
            LSQRB_CUDA(cudaMemcpy(p.data(), M.perm, sizeof(uint32_t) * p.size(), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < p.size(); ++i) perm[i] = (int64_t)p[i];
        }
    }
    return LSQR_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// aprod_ez  (src/lsqr.f90:134-200)
// ---------------------------------------------------------------------------------------------
int lsqr_b200_ez_aprod_device(void *handle, int32_t mode, int32_t m, int32_t n,
                              double *x_dev, double *y_dev, void *stream)
{
    lsqr_b200_ez *me = (lsqr_b200_ez *)handle;
    if (!me) return LSQR_B200_ERR_ARG;
    if (m != me->m || n != me->n) return LSQR_B200_ERR_NOINIT;   // :152
    Work &wk = me->wk;
    cudaStream_t saved = wk.stream;
    if (stream) wk.stream = (cudaStream_t)stream;
    int rc;
    if (mode == 1) {                                                                                                       // y += A x
        rc = LSQR_B200_OK;
        for (int64_t b = 0; b < me->A.nblocks && rc == LSQR_B200_OK; ++b)
            rc = me->stream ? launch_stream<SEPI_ACC>(wk, view_of_block(me->A, b), me->mapA[(size_t)b], x_dev, y_dev, nullptr)
                            : launch_spmv<EPI_ACC>(wk, view_of_block(me->A, b), me->lanes_a, x_dev, y_dev, nullptr);
    }
    else if (mode == 2) {                                                                                                  // x += A'y
        rc = LSQR_B200_OK;
        for (int64_t b = 0; b < me->AT.nblocks && rc == LSQR_B200_OK; ++b)
            rc = me->stream ? launch_stream<SEPI_ACC>(wk, view_of_block(me->AT, b), me->mapAT[(size_t)b], y_dev, x_dev, nullptr)
                            : launch_spmv<EPI_ACC>(wk, view_of_block(me->AT, b), me->lanes_at, y_dev, x_dev, nullptr);
    }
    else                rc = LSQR_B200_ERR_MODE;                                                     // :197
    wk.stream = saved;
    return rc;
}

int lsqr_b200_ez_aprod(lsqr_b200_ez *me, int32_t mode, int32_t m, int32_t n, double *x, double *y)
{
    if (!me) return LSQR_B200_ERR_ARG;
    if (m != me->m || n != me->n) return LSQR_B200_ERR_NOINIT;
    if (mode != 1 && mode != 2) return LSQR_B200_ERR_MODE;
    if (!x || !y) return LSQR_B200_ERR_ARG;
    Work &wk = me->wk;
    LSQRB_CUDA(cudaSetDevice(wk.device));
    double *dx = x, *dy = y;
    const bool xdev = is_device_ptr(x), ydev = is_device_ptr(y);
    if (!xdev) {
        if (!me->tmp_n) LSQRB_CUDA(cudaMalloc(&me->tmp_n, sizeof(double) * (size_t)std::max<int32_t>(n, 1)));
        dx = me->tmp_n;
        LSQRB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, wk.stream));
    }
    if (!ydev) {
        if (!me->tmp_m) LSQRB_CUDA(cudaMalloc(&me->tmp_m, sizeof(double) * (size_t)std::max<int32_t>(m, 1)));
        dy = me->tmp_m;
        LSQRB_CUDA(cudaMemcpyAsync(dy, y, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, wk.stream));
    }
    LSQRB_TRY(lsqr_b200_ez_aprod_device(me, mode, m, n, dx, dy, nullptr));
    if (mode == 1 && !ydev) LSQRB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, wk.stream));
    if (mode == 2 && !xdev) LSQRB_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// solve_ez + LSQR  (src/lsqr.f90:207-259, 432-882) -- fused engine
// ---------------------------------------------------------------------------------------------
enum { CLS_APROD = 0, CLS_ATPROD = 1, CLS_UPDATE = 2, CLS_OTHER = 3 };

struct ProfScope {   // profile mode: one event pair around a launch
    lsqr_b200_ez *me; bool on; size_t slot;
    ProfScope(lsqr_b200_ez *m, int cls) : me(m), on(m->opt.profile != 0), slot(0)
    {
        if (!on) return;
        if (me->prof_cls.size() >= 4096) { on = false; return; }
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        slot = me->prof_ev.size();
        me->prof_ev.push_back(a); me->prof_ev.push_back(b);
        me->prof_cls.push_back(cls);
        cudaEventRecord(a, me->wk.stream);
    }
    ~ProfScope() { if (on) cudaEventRecord(me->prof_ev[slot + 1], me->wk.stream); }
};

static int allreduce_g(lsqr_b200_ez *me)
{
    NcclApi *api = nccl_api();
    LSQRB_NCCL(api->AllReduce(me->g, me->g, (size_t)me->n + 1, ncclFloat64, ncclSum, me->comm, me->wk.stream));
    return LSQR_B200_OK;
}

// the SpMV flavours of one handle
static int do_aprod_fused(lsqr_b200_ez *me, double *aux)
{
    if (!me->a_blocked)
        return me->stream ? launch_stream<SEPI_APROD>(me->wk, view_of(me->A), me->mapA[0], me->v, me->u, aux)
                          : launch_spmv<EPI_FUSED_APROD>(me->wk, view_of(me->A), me->lanes_a, me->v, me->u, aux);
    // column-blocked A: gu = A v block by block (block 0 stores, the others accumulate), then the fused finish
    Work &wk = me->wk;
    StreamExtra ex;
    ex.check_done = 1;
    // tiled kernels: the LAST block's launch finishes the step itself (u' = ca_mat*(gu + s) + ca_vec*u, sum u'^2), which
    // saves one write and one read of gu and a pass over u
    const bool fuse_last = me->fuse_last;
    for (int64_t b = 0; b < me->A.nblocks; ++b) {
        const CsrView V = view_of_block(me->A, b);
        if (b == 0) LSQRB_TRY(me->stream ? launch_stream<SEPI_STORE>(wk, V, me->mapA[0], me->v, me->gu, nullptr, ex)
                                         : launch_spmv<EPI_STORE>(wk, V, me->lanes_a, me->v, me->gu, nullptr, 1));
        else if (fuse_last && b == me->A.nblocks - 1) {
            StreamExtra fx;
            fx.uw = me->gu;
            return launch_stream<SEPI_APROD_ACC>(wk, V, me->mapA[(size_t)b], me->v, me->u, aux, fx);
        }
        else        LSQRB_TRY(me->stream ? launch_stream<SEPI_ACC>(wk, V, me->mapA[(size_t)b], me->v, me->gu, nullptr, ex)
                                         : launch_spmv<EPI_ACC>(wk, V, me->lanes_a, me->v, me->gu, nullptr, 1));
    }
    ufinish_kernel<<<wk.grid_for(me->m, kThreads), kThreads, 0, wk.stream>>>(me->m, me->gu, me->u, wk.st, aux);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}
// g = A'u (unfused: multi-GPU partial and/or row-blocked transpose): block 0 stores, the others accumulate
static int do_atprod_store(lsqr_b200_ez *me)
{
    StreamExtra ex;
    ex.check_done = 1;   // over-enqueued iterations after the stop are no-ops
    for (int64_t b = 0; b < me->AT.nblocks; ++b) {
        const CsrView V = view_of_block(me->AT, b);
        if (b == 0) LSQRB_TRY(me->stream ? launch_stream<SEPI_STORE>(me->wk, V, me->mapAT[0], me->u, me->g, nullptr, ex)
                                         : launch_spmv<EPI_STORE>(me->wk, V, me->lanes_at, me->u, me->g, nullptr, 1));
        else        LSQRB_TRY(me->stream ? launch_stream<SEPI_ACC>(me->wk, V, me->mapAT[(size_t)b], me->u, me->g, nullptr, ex)
                                         : launch_spmv<EPI_ACC>(me->wk, V, me->lanes_at, me->u, me->g, nullptr, 1));
    }
    return LSQR_B200_OK;
}
// multi-GPU: g = sum over ranks of [A_p'u_p | sum u_p^2].  With comm_chunks > 1 the all-reduce of column range c
// (on comm_stream) overlaps the SpMV launches of range c+1; the Atprod is then ordered chunk-major, block-minor.
static int do_atprod_allreduce(lsqr_b200_ez *me)
{
    Work &wk = me->wk;
    NcclApi *api = nccl_api();
    const int k = me->comm_chunks;
    if (k <= 1) {
        LSQRB_TRY(do_atprod_store(me));
        return allreduce_g(me);
    }
    StreamExtra ex;
    ex.check_done = 1;
    const int64_t nb = me->AT.nblocks;
    for (int64_t b = 0; b + 1 < nb; ++b) {               // all but the last block: whole-width launches
        const CsrView V = view_of_block(me->AT, b);
        if (b == 0) LSQRB_TRY(launch_stream<SEPI_STORE>(wk, V, me->mapAT[0], me->u, me->g, nullptr, ex));
        else        LSQRB_TRY(launch_stream<SEPI_ACC>(wk, V, me->mapAT[(size_t)b], me->u, me->g, nullptr, ex));
    }
    for (int c = 0; c < k; ++c) {                        // last block, range by range, each followed by its all-reduce
        const int64_t c0 = me->cc[(size_t)c], c1 = me->cc[(size_t)c + 1];
        CsrView V = view_of_block(me->AT, nb - 1);
        V.ptr += c0;
        V.nrows = c1 - c0;
        if (nb == 1) LSQRB_TRY(launch_stream<SEPI_STORE>(wk, V, me->mapATc[(size_t)c], me->u, me->g + c0, nullptr, ex));
        else         LSQRB_TRY(launch_stream<SEPI_ACC>(wk, V, me->mapATc[(size_t)c], me->u, me->g + c0, nullptr, ex));
        LSQRB_CUDA(cudaEventRecord(me->ev_chunk[c], wk.stream));
        LSQRB_CUDA(cudaStreamWaitEvent(me->comm_stream, me->ev_chunk[c], 0));
        const size_t count = (size_t)(c1 - c0) + (c == k - 1 ? 1 : 0);     // the last range carries sum(u_p^2) in g[n]
        LSQRB_NCCL(api->AllReduce(me->g + c0, me->g + c0, count, ncclFloat64, ncclSum, me->comm, me->comm_stream));
    }
    LSQRB_CUDA(cudaEventRecord(me->ev_comm, me->comm_stream));
    LSQRB_CUDA(cudaStreamWaitEvent(wk.stream, me->ev_comm, 0));
    return LSQR_B200_OK;
}

static int do_atprod_fused(lsqr_b200_ez *me)
{
    return me->stream ? launch_stream<SEPI_ATPROD>(me->wk, view_of(me->AT), me->mapAT[0], me->u, me->v, nullptr)
                      : launch_spmv<EPI_FUSED_ATPROD>(me->wk, view_of(me->AT), me->lanes_at, me->u, me->v, nullptr);
}
static int do_atprod_init(lsqr_b200_ez *me)
{
    return me->stream ? launch_stream<SEPI_INIT_ATPROD>(me->wk, view_of(me->AT), me->mapAT[0], me->u, me->v, nullptr)
                      : launch_spmv<EPI_INIT_ATPROD>(me->wk, view_of(me->AT), me->lanes_at, me->u, me->v, nullptr);
}
static int do_atprod_upd(lsqr_b200_ez *me)   // Atprod of this iteration + deferred x/w update of the previous one
{
    StreamExtra ex;
    ex.ux = me->x; ex.uw = me->w; ex.use = me->se;
    return launch_stream<SEPI_ATPROD_UPD>(me->wk, view_of(me->AT), me->mapAT[0], me->u, me->v, nullptr, ex);
}

// one LSQR iteration, enqueued (no host synchronisation)
static int enqueue_iteration(lsqr_b200_ez *me, bool wantse)
{
    Work &wk = me->wk;
    if (me->opt.world_size > 1 || me->blocked) {
        // unfused pipeline: u' and its partial norm, g = [A'u' | sum u'^2] (all-reduced over the ranks), then
        // v' = g/beta - (beta/alpha) v with both scalar steps, then the x/w update
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod_fused(me, me->g + me->n)); }
        if (me->opt.world_size > 1) { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_allreduce(me)); }   // (timed together)
        else                        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_store(me)); }
        { ProfScope p(me, CLS_OTHER);
          vfinish_kernel<false><<<wk.grid_for(me->n, kThreads), kThreads, 0, wk.stream>>>(me->n, me->g, me->v, wk.st);
          wk.launches++; LSQRB_CUDA(cudaGetLastError()); }
    } else if (me->deferred) {
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod_fused(me, nullptr)); }
        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_upd(me)); }
        return LSQR_B200_OK;
    } else if (me->overlap_update) {
        // K3(k+1) reads v, writes u; K5(k) reads v, writes x, w: disjoint, so the update leaves the critical path.
        // K4 overwrites v and needs ||w||^2, so it joins the side stream first.
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod_fused(me, nullptr)); }
        if (me->side_pending) { LSQRB_CUDA(cudaStreamWaitEvent(wk.stream, me->ev_join, 0)); me->side_pending = false; }
        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_fused(me)); }
        LSQRB_CUDA(cudaEventRecord(me->ev_fork, wk.stream));
        LSQRB_CUDA(cudaStreamWaitEvent(me->side, me->ev_fork, 0));
        {
            cudaStream_t main_stream = wk.stream;
            wk.stream = me->side;
            int rc = launch_update<true>(wk, me->n, me->x, me->w, me->v, me->se, wantse);
            wk.stream = main_stream;
            LSQRB_TRY(rc);
        }
        LSQRB_CUDA(cudaEventRecord(me->ev_join, me->side));
        me->side_pending = true;
        return LSQR_B200_OK;
    } else {
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod_fused(me, nullptr)); }
        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_fused(me)); }
    }
    { ProfScope p(me, CLS_UPDATE); LSQRB_TRY(launch_update<true>(wk, me->n, me->x, me->w, me->v, me->se, wantse)); }
    return LSQR_B200_OK;
}

// the side stream must be joined before anything else (graph capture end, result copies, the next batch's bookkeeping)
static int join_side(lsqr_b200_ez *me)
{
    if (me->side_pending) {
        LSQRB_CUDA(cudaStreamWaitEvent(me->wk.stream, me->ev_join, 0));
        me->side_pending = false;
    }
    return LSQR_B200_OK;
}

static int build_graph(lsqr_b200_ez *me, bool wantse)
{
    if (me->graph_exec && me->graph_wantse == (int)wantse) return LSQR_B200_OK;
    if (me->graph_exec) { cudaGraphExecDestroy(me->graph_exec); me->graph_exec = nullptr; }
    Work &wk = me->wk;
    cudaGraph_t graph = nullptr;
    const int64_t saved = wk.launches;
    LSQRB_CUDA(cudaStreamBeginCapture(wk.stream, cudaStreamCaptureModeThreadLocal));
    int rc = LSQR_B200_OK;
    for (int i = 0; i < me->batch && rc == LSQR_B200_OK; ++i) rc = enqueue_iteration(me, wantse);
    if (rc == LSQR_B200_OK) rc = join_side(me);   // every fork rejoins the origin stream before the capture ends
    me->side_pending = false;
    cudaError_t e = cudaStreamEndCapture(wk.stream, &graph);
    wk.launches = saved;
    if (rc != LSQR_B200_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    LSQRB_CUDA(e);
    LSQRB_CUDA(cudaGraphInstantiate(&me->graph_exec, graph, 0));
    cudaGraphDestroy(graph);
    me->graph_wantse = (int)wantse;
    return LSQR_B200_OK;
}

static int ez_solve_fused(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                          double *se, int32_t *itn, double *anorm, double *acond,
                          double *rnorm, double *arnorm, double *xnorm)
{
    Work &wk = me->wk;
    const int64_t m = me->m, n = me->n;
    const bool wantse = se != nullptr;
    const bool dist = me->opt.world_size > 1;
    const bool unfused = dist || me->blocked;
    LSQRB_CUDA(cudaSetDevice(wk.device));
    if (wantse && !me->se) LSQRB_CUDA(cudaMalloc(&me->se, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
    wk.launches = 0;
    for (auto e : me->prof_ev) cudaEventDestroy(e);
    me->prof_ev.clear(); me->prof_cls.clear();

    LogCtx lc;
    lc.log = me->opt.log; lc.log_user = me->opt.log_user; lc.iter = me->opt.iter; lc.iter_user = me->opt.iter_user;
    lc.m = dist && me->opt.m_global > 0 ? me->opt.m_global : m; lc.n = n; lc.damp = damp;
    lc.atol = me->opt.atol; lc.btol = me->opt.btol; lc.conlim = me->opt.conlim;
    lc.ctol = me->opt.conlim > 0.0 ? 1.0 / me->opt.conlim : 0.0;
    lc.itnlim = me->opt.itnlim; lc.wantse = wantse;
    lc.header();

    LSQRB_CUDA(cudaEventRecord(me->ev_t0, wk.stream));
    LSQRB_TRY(wk.reset_state(damp, me->opt.atol, me->opt.btol, me->opt.conlim, me->opt.itnlim, wantse, dist));

    // u = b (:242); v = 0, x = 0, se = 0 (:621-630)
    if (m > 0) LSQRB_CUDA(cudaMemcpyAsync(me->u, b, sizeof(double) * (size_t)m, cudaMemcpyDefault, wk.stream));
    if (n > 0) {
        LSQRB_CUDA(cudaMemsetAsync(me->v, 0, sizeof(double) * (size_t)n, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(me->x, 0, sizeof(double) * (size_t)n, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(me->w, 0, sizeof(double) * (size_t)n, wk.stream));
        if (wantse) LSQRB_CUDA(cudaMemsetAsync(me->se, 0, sizeof(double) * (size_t)n, wk.stream));
    }
    // beta = ||u||; v = A'(u/beta); alpha = ||v||; w = v/alpha  (:632-644), lazily normalised
    if (unfused) {
        sumsq_kernel<POST_NONE><<<wk.grid_for(m, kThreads), kThreads, 0, wk.stream>>>(m, me->u, wk.st, me->g + n);
        wk.launches++;
        if (dist) LSQRB_TRY(do_atprod_allreduce(me));
        else      LSQRB_TRY(do_atprod_store(me));
        vfinish_kernel<true><<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->g, me->v, wk.st);
        wk.launches++;
    } else {
        sumsq_kernel<POST_INIT_BETA><<<wk.grid_for(m, kThreads), kThreads, 0, wk.stream>>>(m, me->u, wk.st, nullptr);
        wk.launches++;
        LSQRB_TRY(do_atprod_init(me));
    }
    init_w_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->w, me->v, wk.st);
    record0_kernel<<<1, 1, 0, wk.stream>>>(wk.st, wk.ring_d);
    wk.launches += 2;
    LSQRB_CUDA(cudaGetLastError());
    LSQRB_CUDA(cudaEventRecord(me->ev_t1, wk.stream));

    // ---- iteration loop: enqueue batch j+1, then wait for batch j and look at its records -----
    const bool use_graph = me->opt.use_graph && !me->opt.profile && !dist;
    if (use_graph) LSQRB_TRY(build_graph(me, wantse));
    const int itnlim = std::max(me->opt.itnlim, 1);   // the reference always runs one iteration (:673-676,798)
    const int B = me->batch;
    int seen = -1, enq = 0, nb = 0, done_batches = 0;   // records consumed, iterations / batches enqueued, batches finished
    bool stop = false;
    auto enqueue_batch = [&]() -> int {
        if (use_graph) {
            LSQRB_CUDA(cudaGraphLaunch(me->graph_exec, wk.stream));
            wk.launches += (int64_t)B * (me->deferred ? 2 : (me->blocked ? 2 + me->AT.nblocks : 2) + (me->a_blocked ? (me->fuse_last ? 0 : 1) + me->A.nblocks : 1));
        } else {
            for (int i = 0; i < B; ++i) LSQRB_TRY(enqueue_iteration(me, wantse));
            LSQRB_TRY(join_side(me));
        }
        LSQRB_CUDA(cudaEventRecord(wk.ev[nb & 3], wk.stream));
        enq += B;
        nb += 1;
        return LSQR_B200_OK;
    };
    auto check = [&](int upto) {
        stop = drain_ring(wk, lc, seen, upto);
        if (seen >= 0 && wk.ring_h[0].arnorm == 0.0) stop = true;   // alpha*beta = 0: no iterations (:646-648)
    };
    // Keep two batches in flight.  The device stops by itself (done flag; istop = 5 at itnlim), so an
    // over-enqueued batch is a run of no-op kernels.  Launch decisions depend only on the records of
    // fully finished batches, which makes them identical on every rank of a multi-GPU run (all ranks
    // must enqueue the same sequence of all-reduces).
    while (!stop) {
        while (nb < done_batches + 2 && enq < itnlim) LSQRB_TRY(enqueue_batch());
        if (done_batches == nb) {
            LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
            check(itnlim);
            break;
        }
        LSQRB_CUDA(cudaEventSynchronize(wk.ev[done_batches & 3]));
        done_batches += 1;
        check(std::min(done_batches * B, itnlim));
    }
    // deferred-update engine: the x/w update of the stopping iteration may still be outstanding
    if (me->deferred) LSQRB_TRY(do_atprod_upd(me));
    LSQRB_CUDA(cudaEventRecord(me->ev_t2, wk.stream));

    // se(i) = rnorm/sqrt(t) sqrt(se(i))  (:857-865)
    if (wantse && n > 0) {
        const int64_t mg = dist && me->opt.m_global > 0 ? me->opt.m_global : m;
        double t = 1.0;
        if (mg > n) t = (double)(mg - n);
        if (damp > 0.0) t = (double)mg;
        se_finish_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->se, wk.st, t);
        wk.launches++;
        LSQRB_CUDA(cudaMemcpyAsync(se, me->se, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    }
    if (n > 0) LSQRB_CUDA(cudaMemcpyAsync(x, me->x, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    LSQRB_TRY(wk.fetch_state());
    // any records the loop did not print yet (e.g. the stopping iteration)
    drain_ring(wk, lc, seen, wk.h.itn);

    int is = wk.h.istop;
    if (damp > 0.0 && is == 2) is = 3;   // :871
    lc.footer(is, wk.h);
    if (istop) *istop = is;
    if (itn) *itn = wk.h.itn;
    if (anorm) *anorm = wk.h.anorm;
    if (acond) *acond = wk.h.acond;
    if (rnorm) *rnorm = wk.h.rnorm;
    if (arnorm) *arnorm = wk.h.arnorm;
    if (xnorm) *xnorm = wk.h.xnorm;

    // timings
    float ms = 0.f;
    me->times.total_launches = wk.launches;
    if (cudaEventElapsedTime(&ms, me->ev_t0, me->ev_t1) == cudaSuccess) me->times.init_ms = ms;
    if (cudaEventElapsedTime(&ms, me->ev_t1, me->ev_t2) == cudaSuccess) me->times.loop_ms = ms;
    if (me->opt.profile) {
        double acc[4] = {0, 0, 0, 0};
        int64_t cnt[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < me->prof_cls.size(); ++i) {
            if (cudaEventElapsedTime(&ms, me->prof_ev[2 * i], me->prof_ev[2 * i + 1]) == cudaSuccess) {
                acc[me->prof_cls[i]] += ms;
                cnt[me->prof_cls[i]] += 1;
            }
        }
        me->times.aprod_ms = cnt[0] ? acc[0] / cnt[0] : 0;   me->times.aprod_launches = cnt[0];
        me->times.atprod_ms = cnt[1] ? acc[1] / cnt[1] : 0;  me->times.atprod_launches = cnt[1];
        me->times.update_ms = cnt[2] ? acc[2] / cnt[2] : 0;  me->times.update_launches = cnt[2];
        me->times.other_ms = cnt[3] ? acc[3] / cnt[3] : 0;   me->times.other_launches = cnt[3];
    }
    return LSQR_B200_OK;
}

static int ez_solve_reference_structure(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                                        double *se, int32_t *itn, double *anorm, double *acond,
                                        double *rnorm, double *arnorm, double *xnorm);

int lsqr_b200_ez_solve(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                       double *se, int32_t *itn, double *anorm, double *acond,
                       double *rnorm, double *arnorm, double *xnorm)
{
    if (!me) return LSQR_B200_ERR_ARG;
    if ((me->m > 0 && !b) || (me->n > 0 && !x)) { set_last_error("NULL b or x"); return LSQR_B200_ERR_ARG; }
    if (me->opt.engine == 1 && me->opt.world_size == 1)
        return ez_solve_reference_structure(me, b, damp, x, istop, se, itn, anorm, acond, rnorm, arnorm, xnorm);
    return ez_solve_fused(me, b, damp, x, istop, se, itn, anorm, acond, rnorm, arnorm, xnorm);
}

int lsqr_b200_ez_get_kernel_times(const lsqr_b200_ez *me, lsqr_b200_kernel_times *out)
{
    if (!me || !out) return LSQR_B200_ERR_ARG;
    *out = me->times;
    return LSQR_B200_OK;
}

}  // extern "C"

// =============================================================================================
// Operator-hook path: LSQR with the reference's pass structure (dscal / aprod / dnrm2 as
// separate steps, u and v kept normalised), src/lsqr.f90:432-882
// =============================================================================================
namespace lsqrb {

static int scal_dev(Work &wk, int64_t n, double *x, const double *coef)
{
    scal_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, x, coef, 1.0);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

template <int POST>
static int sumsq_dev(Work &wk, int64_t n, const double *x, double *result)
{
    sumsq_kernel<POST><<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, x, wk.st, result);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

static int lsqr_with_operator(Work &wk, lsqr_b200_aprod_fn aprod, void *aprod_user,
                              int32_t m, int32_t n, double damp, int wantse,
                              double *u, double *v, double *w, double *x, double *se,
                              double atol, double btol, double conlim, int32_t itnlim,
                              const lsqr_b200_options *opts,
                              int32_t *istop, int32_t *itn, double *anorm, double *acond,
                              double *rnorm, double *arnorm, double *xnorm)
{
    LSQRB_CUDA(cudaSetDevice(wk.device));
    DevState *st = wk.st;
    LogCtx lc;
    if (opts) { lc.log = opts->log; lc.log_user = opts->log_user; lc.iter = opts->iter; lc.iter_user = opts->iter_user; }
    lc.m = m; lc.n = n; lc.damp = damp; lc.atol = atol; lc.btol = btol; lc.conlim = conlim;
    lc.ctol = conlim > 0.0 ? 1.0 / conlim : 0.0; lc.itnlim = itnlim; lc.wantse = wantse;
    lc.header();
    LSQRB_TRY(wk.reset_state(damp, atol, btol, conlim, itnlim, wantse, 0));

    auto call_aprod = [&](int mode) -> int {
        int rc = aprod(aprod_user, mode, m, n, v, u, (void *)wk.stream);
        if (rc != 0) { set_last_error("aprod callback returned " + std::to_string(rc)); return LSQR_B200_ERR_CALLBACK; }
        return LSQR_B200_OK;
    };

    // :621-644
    if (n > 0) {
        LSQRB_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * (size_t)n, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, wk.stream));
        if (wantse) LSQRB_CUDA(cudaMemsetAsync(se, 0, sizeof(double) * (size_t)n, wk.stream));
    }
    LSQRB_TRY(sumsq_dev<POST_INIT_BETA>(wk, m, u, nullptr));
    LSQRB_TRY(scal_dev(wk, m, u, &st->g_c1));            // u /= beta
    LSQRB_TRY(call_aprod(2));                            // v += A'u
    LSQRB_TRY(sumsq_dev<POST_INIT_ALPHA>(wk, n, v, nullptr));
    LSQRB_TRY(scal_dev(wk, n, v, &st->g_c3));            // v /= alpha
    LSQRB_CUDA(cudaMemcpyAsync(w, v, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, wk.stream));   // dcopy
    record0_kernel<<<1, 1, 0, wk.stream>>>(st, wk.ring_d);
    wk.launches++;
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));

    int seen = -1;
    bool stop = drain_ring(wk, lc, seen, 0);
    if (wk.ring_h[0].istop != 0.0 || wk.ring_h[0].arnorm == 0.0) stop = true;   // alpha*beta = 0 (:646)
    int k = 0;
    while (!stop && k < itnlim) {
        ++k;
        LSQRB_TRY(scal_dev(wk, m, u, &st->g_c0));        // u *= -alpha            (:681)
        LSQRB_TRY(call_aprod(1));                        // u += A v               (:682)
        LSQRB_TRY(sumsq_dev<POST_G_BETA>(wk, m, u, nullptr));   // beta, anorm     (:683-689)
        LSQRB_TRY(scal_dev(wk, m, u, &st->g_c1));        // u /= beta              (:692)
        LSQRB_TRY(scal_dev(wk, n, v, &st->g_c2));        // v *= -beta             (:693)
        LSQRB_TRY(call_aprod(2));                        // v += A'u               (:694)
        LSQRB_TRY(sumsq_dev<POST_G_ALPHA>(wk, n, v, nullptr));  // alpha, rotations, tests (:695-810)
        LSQRB_TRY(scal_dev(wk, n, v, &st->g_c3));        // v /= alpha             (:697)
        LSQRB_TRY(launch_update<false>(wk, n, x, w, v, se, wantse != 0));   // (:729-745)
        LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
        stop = drain_ring(wk, lc, seen, k);
    }
    if (wantse && n > 0) {
        double t = 1.0;
        if (m > n) t = (double)(m - n);
        if (damp > 0.0) t = (double)m;
        se_finish_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, se, st, t);
        wk.launches++;
    }
    LSQRB_TRY(wk.fetch_state());
    int is = wk.h.istop;
    if (damp > 0.0 && is == 2) is = 3;
    lc.footer(is, wk.h);
    if (istop) *istop = is;
    if (itn) *itn = wk.h.itn;
    if (anorm) *anorm = wk.h.anorm;
    if (acond) *acond = wk.h.acond;
    if (rnorm) *rnorm = wk.h.rnorm;
    if (arnorm) *arnorm = wk.h.arnorm;
    if (xnorm) *xnorm = wk.h.xnorm;
    return LSQR_B200_OK;
}

// generators of acheck's "unlikely" vectors (src/lsqr.f90:946-961)
__global__ void acheck_fill_kernel(int64_t n, double *x, int reciprocal)
{
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double t = sqrt((double)(i + 2));
        x[i] = reciprocal ? 1.0 / t : t;
    }
}

__global__ void xcheck_w_kernel(int64_t n, double *w, const double *v, const double *x, double dampsq)
{
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        w[i] = v[i] - dampsq * x[i];
}

static int reduce_to_host(Work &wk, int64_t n, const double *x, const double *y, double *out)
{
    double *d_res = &wk.st->partial[kMaxPartials - 1];   // last slot is never used as a block partial (grid < kMaxPartials)
    if (y) dot_kernel<<<std::min(wk.grid_for(n, kThreads), kMaxPartials - 1), kThreads, 0, wk.stream>>>(n, x, y, wk.st, d_res);
    else   sumsq_kernel<POST_NONE><<<std::min(wk.grid_for(n, kThreads), kMaxPartials - 1), kThreads, 0, wk.stream>>>(n, x, wk.st, d_res);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    LSQRB_CUDA(cudaMemcpyAsync(out, d_res, sizeof(double), cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

static int scal_imm(Work &wk, int64_t n, double *x, double a)
{
    scal_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, x, nullptr, a);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

struct ScopedWork {   // a Work for handle-less entry points
    Work wk;
    bool ok = false;
    int rc;
    ScopedWork(const lsqr_b200_options *o, void *stream_override = nullptr)
    {
        rc = wk.init(o ? o->device : -1, stream_override ? stream_override : (o ? o->stream : nullptr));
        ok = rc == LSQR_B200_OK;
    }
    ~ScopedWork() { wk.destroy(); }
};

}  // namespace lsqrb

static int ez_solve_reference_structure(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                                        double *se, int32_t *itn, double *anorm, double *acond,
                                        double *rnorm, double *arnorm, double *xnorm)
{
    Work &wk = me->wk;
    const int64_t m = me->m, n = me->n;
    const bool wantse = se != nullptr;
    LSQRB_CUDA(cudaSetDevice(wk.device));
    if (wantse && !me->se) LSQRB_CUDA(cudaMalloc(&me->se, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
    wk.launches = 0;
    if (m > 0) LSQRB_CUDA(cudaMemcpyAsync(me->u, b, sizeof(double) * (size_t)m, cudaMemcpyDefault, wk.stream));
    lsqr_b200_options o = me->opt;
    LSQRB_TRY(lsqr_with_operator(wk, lsqr_b200_ez_aprod_device, me, me->m, me->n, damp, wantse,
                                 me->u, me->v, me->w, me->x, me->se,
                                 me->opt.atol, me->opt.btol, me->opt.conlim, me->opt.itnlim, &o,
                                 istop, itn, anorm, acond, rnorm, arnorm, xnorm));
    if (n > 0) LSQRB_CUDA(cudaMemcpyAsync(x, me->x, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    if (wantse && n > 0) LSQRB_CUDA(cudaMemcpyAsync(se, me->se, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    me->times.total_launches = wk.launches;
    return LSQR_B200_OK;
}

extern "C" {

int lsqr_b200_lsqr(lsqr_b200_aprod_fn aprod, void *aprod_user,
                   int32_t m, int32_t n, double damp, int32_t wantse,
                   double *u, double *v, double *w, double *x, double *se,
                   double atol, double btol, double conlim, int32_t itnlim,
                   const lsqr_b200_options *opts,
                   int32_t *istop, int32_t *itn, double *anorm, double *acond,
                   double *rnorm, double *arnorm, double *xnorm)
{
    if (!aprod || m < 0 || n < 0) return LSQR_B200_ERR_ARG;
    if ((m > 0 && !u) || (n > 0 && (!v || !w || !x)) || (wantse && n > 0 && !se)) {
        set_last_error("NULL work vector");
        return LSQR_B200_ERR_ARG;
    }
    ScopedWork sw(opts);
    if (!sw.ok) return sw.rc;
    return lsqr_with_operator(sw.wk, aprod, aprod_user, m, n, damp, wantse, u, v, w, x, se,
                              atol, btol, conlim, itnlim, opts, istop, itn, anorm, acond, rnorm, arnorm, xnorm);
}

// acheck, src/lsqr.f90:908-994
int lsqr_b200_acheck(lsqr_b200_aprod_fn aprod, void *aprod_user, int32_t m, int32_t n,
                     double eps, double *v, double *w, double *x, double *y,
                     const lsqr_b200_options *opts, int32_t *inform, double *relerr)
{
    if (!aprod || m < 1 || n < 1 || !v || !w || !x || !y) return LSQR_B200_ERR_ARG;
    ScopedWork sw(opts);
    if (!sw.ok) return sw.rc;
    Work &wk = sw.wk;
    const double tol = pow(eps, 0.5);   // power = 0.5 (:927)
    if (opts && opts->log) { opts->log(opts->log_user, ""); opts->log(opts->log_user, ""); opts->log(opts->log_user, "Enter acheck. Test of aprod for LSQR and CRAIG"); }
    acheck_fill_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, x, 0);
    acheck_fill_kernel<<<wk.grid_for(m, kThreads), kThreads, 0, wk.stream>>>(m, y, 1);
    double alfa, beta;
    LSQRB_TRY(reduce_to_host(wk, n, x, nullptr, &alfa));
    LSQRB_TRY(reduce_to_host(wk, m, y, nullptr, &beta));
    alfa = sqrt(alfa); beta = sqrt(beta);
    LSQRB_TRY(scal_imm(wk, n, x, 1.0 / alfa));
    LSQRB_TRY(scal_imm(wk, m, y, 1.0 / beta));
    LSQRB_CUDA(cudaMemcpyAsync(w, y, sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, wk.stream));
    LSQRB_CUDA(cudaMemcpyAsync(v, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, wk.stream));
    if (aprod(aprod_user, 1, m, n, x, w, (void *)wk.stream) != 0) return LSQR_B200_ERR_CALLBACK;   // w = y + A x
    if (aprod(aprod_user, 2, m, n, v, y, (void *)wk.stream) != 0) return LSQR_B200_ERR_CALLBACK;   // v = x + A'y
    LSQRB_TRY(reduce_to_host(wk, m, y, w, &alfa));   // y'w
    LSQRB_TRY(reduce_to_host(wk, n, x, v, &beta));   // x'v
    const double test1 = fabs(alfa - beta);
    const double test2 = 1.0 + fabs(alfa) + fabs(beta);
    const double test3 = test1 / test2;
    if (inform) *inform = test3 <= tol ? 0 : 1;
    if (relerr) *relerr = test3;
    if (opts && opts->log) {
        std::string s = (test3 <= tol ? "aprod seems OK. Relative error = " : "aprod seems incorrect. Relative error = ") + fe(10, 1, test3);
        opts->log(opts->log_user, s.c_str());
    }
    return LSQR_B200_OK;
}

// xcheck, src/lsqr.f90:1015-1154
int lsqr_b200_xcheck(lsqr_b200_aprod_fn aprod, void *aprod_user, int32_t m, int32_t n,
                     double anorm, double damp, double eps,
                     const double *b, double *u, double *v, double *w, const double *x,
                     const lsqr_b200_options *opts,
                     int32_t *inform, double *test1, double *test2, double *test3, double *norms)
{
    if (!aprod || m < 1 || n < 1 || !b || !u || !v || !w || !x) return LSQR_B200_ERR_ARG;
    ScopedWork sw(opts);
    if (!sw.ok) return sw.rc;
    Work &wk = sw.wk;
    const double dampsq = damp * damp;
    const double tol = pow(eps, 0.5);
    double *xtmp = nullptr;   // the reference copies x because aprod's x is intent(inout) (:1064)
    LSQRB_CUDA(cudaMalloc(&xtmp, sizeof(double) * (size_t)n));
    LSQRB_CUDA(cudaMemcpyAsync(xtmp, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, wk.stream));
    // u = b - A x via u = -b + A x, u = -u (:1069-1076)
    LSQRB_CUDA(cudaMemcpyAsync(u, b, sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, wk.stream));
    LSQRB_TRY(scal_imm(wk, m, u, -1.0));
    int rc = aprod(aprod_user, 1, m, n, xtmp, u, (void *)wk.stream);
    if (rc == 0) rc = scal_imm(wk, m, u, -1.0);
    // v = A'u (:1080-1083)
    if (rc == 0) rc = cudaMemsetAsync(v, 0, sizeof(double) * (size_t)n, wk.stream) == cudaSuccess ? 0 : LSQR_B200_ERR_CUDA;
    if (rc == 0) rc = aprod(aprod_user, 2, m, n, v, u, (void *)wk.stream);
    cudaStreamSynchronize(wk.stream);
    cudaFree(xtmp);
    if (rc != 0) return rc == LSQR_B200_ERR_CUDA ? rc : LSQR_B200_ERR_CALLBACK;
    // w = A'u - damp^2 x (:1089-1094)
    xcheck_w_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, w, v, x, damp != 0.0 ? dampsq : 0.0);
    double bnorm, xnorm, rho1, sigma1, rho2, sigma2;
    LSQRB_TRY(reduce_to_host(wk, m, b, nullptr, &bnorm));
    LSQRB_TRY(reduce_to_host(wk, n, x, nullptr, &xnorm));
    LSQRB_TRY(reduce_to_host(wk, m, u, nullptr, &rho1));
    LSQRB_TRY(reduce_to_host(wk, n, v, nullptr, &sigma1));
    bnorm = sqrt(bnorm); xnorm = sqrt(xnorm); rho1 = sqrt(rho1); sigma1 = sqrt(sigma1);
    if (damp == 0.0) {
        rho2 = rho1;
        sigma2 = sigma1;
    } else {
        rho2 = sqrt(rho1 * rho1 + dampsq * (xnorm * xnorm));
        LSQRB_TRY(reduce_to_host(wk, n, w, nullptr, &sigma2));
        sigma2 = sqrt(sigma2);
    }
    int inf;
    double t1 = 0, t2 = 0, t3 = 0;
    if (bnorm == 0.0 && xnorm == 0.0) {
        inf = 0;
    } else {
        inf = 4;
        t1 = rho1 / (bnorm + anorm * xnorm);
        t2 = 0.0;
        if (rho1 > 0.0) t2 = sigma1 / (anorm * rho1);
        t3 = t2;
        if (rho2 > 0.0) t3 = sigma2 / (anorm * rho2);
        if (t3 <= tol) inf = 3;
        if (t2 <= tol) inf = 2;
        if (t1 <= tol) inf = 1;
    }
    if (inform) *inform = inf;
    if (test1) *test1 = t1;
    if (test2) *test2 = t2;
    if (test3) *test3 = t3;
    if (norms) { norms[0] = bnorm; norms[1] = xnorm; norms[2] = rho1; norms[3] = sigma1; norms[4] = rho2; norms[5] = sigma2; }
    if (opts && opts->log) {
        auto L = [&](const std::string &s) { opts->log(opts->log_user, s.c_str()); };
        L(""); L("");
        L("Enter xcheck. Does x solve Ax = b, etc?");
        L(" damp            =" + fe(10, 3, damp));
        L(" norm(x)         =" + fe(10, 3, xnorm));
        L(" norm(r)         =" + fe(15, 8, rho1) + " = rho1");
        L(" norm(A'r)       =" + fe(10, 3, sigma1) + "      = sigma1");
        if (damp != 0.0) {
            L("");
            L(" norm(s)         =" + fe(10, 3, rho1 / damp));
            L(" norm(x,s)       =" + fe(10, 3, rho2 / damp));
            L(" norm(rbar)      =" + fe(15, 8, rho2) + " = rho2");
            L(" norm(Abar'rbar) =" + fe(10, 3, sigma2) + "      = sigma2");
        }
        L("");
        char buf[64];
        snprintf(buf, sizeof buf, " inform          =%2d", inf);
        L(buf);
        L(" tol             =" + fe(10, 3, tol));
        L(" test1           =" + fe(10, 3, t1) + " (Ax = b)");
        L(" test2           =" + fe(10, 3, t2) + " (least-squares)");
        L(" test3           =" + fe(10, 3, t3) + " (damped least-squares)");
    }
    return LSQR_B200_OK;
}

// ---- device BLAS-1 (src/lsqrblas.f90), stride 1 ------------------------------------------------
int lsqr_b200_dnrm2(int64_t n, const double *x, double *result, void *stream)
{
    if (!result || n < 0 || (n > 0 && !x)) return LSQR_B200_ERR_ARG;
    if (n < 1) { *result = 0.0; return LSQR_B200_OK; }   // :131
    ScopedWork sw(nullptr, stream);
    if (!sw.ok) return sw.rc;
    double s;
    LSQRB_TRY(reduce_to_host(sw.wk, n, x, nullptr, &s));
    *result = sqrt(s);
    return LSQR_B200_OK;
}

int lsqr_b200_ddot(int64_t n, const double *x, const double *y, double *result, void *stream)
{
    if (!result || n < 0 || (n > 0 && (!x || !y))) return LSQR_B200_ERR_ARG;
    if (n < 1) { *result = 0.0; return LSQR_B200_OK; }
    ScopedWork sw(nullptr, stream);
    if (!sw.ok) return sw.rc;
    return reduce_to_host(sw.wk, n, x, y, result);
}

int lsqr_b200_dscal(int64_t n, double da, double *x, void *stream)
{
    if (n < 0 || (n > 0 && !x)) return LSQR_B200_ERR_ARG;
    if (n == 0) return LSQR_B200_OK;
    ScopedWork sw(nullptr, stream);
    if (!sw.ok) return sw.rc;
    LSQRB_TRY(scal_imm(sw.wk, n, x, da));
    LSQRB_CUDA(cudaStreamSynchronize(sw.wk.stream));
    return LSQR_B200_OK;
}

int lsqr_b200_dcopy(int64_t n, const double *x, double *y, void *stream)
{
    if (n < 0 || (n > 0 && (!x || !y))) return LSQR_B200_ERR_ARG;
    if (n == 0) return LSQR_B200_OK;
    int count = lsqr_b200_device_count();
    if (count == 0) return LSQR_B200_ERR_NO_DEVICE;
    LSQRB_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    LSQRB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return LSQR_B200_OK;
}

}  // extern "C"
