// plan.cuh -- host side of the SpMV kernel (spmv.cuh): the work plan of one stored matrix and its launches.
//
// A plan is built once at initialize from the device CSR: the row cuts (the same in every block of a blocked
// matrix), the gather span of every piece, the shared-memory window the launch will stage, the kernel flavour
// (4 or 2 resident CTAs per SM) and, for very uneven rows, a balanced tile schedule.  Everything is a pure
// function of the matrix structure, so results stay reproducible run to run.
#pragma once

#include <algorithm>
#include <cmath>
#include <queue>
#include <vector>

#include "build_csr.h"
#include "host.cuh"
#include "spmv.cuh"

namespace lsqrb {

struct TilePlan {
    TileDesc *tiles = nullptr;   // device: [nblocks][ntiles + 1]
    int ntiles = 0, nblocks = 1;
    uint64_t tile = 0;           // nominal work units per tile (all blocks together)
    uint32_t row_w = 4;          // weight of a row (per block) in the tile cut
    uint32_t *order = nullptr;   // device: balanced schedule, or nullptr = round robin
    int nslots = 0;
    int ctas = 0;                // persistent grid the plan was cut for
    int epl = 4;                 // stored entries per lane and chunk of the kernel flavour in use
    int win_cap = 0;             // doubles of gather window per warp (0 = no staging)
    size_t smem = 0;             // dynamic shared memory per CTA
    double imbalance = 1.0;      // most loaded warp / mean load under the schedule in use (diagnostic)
    double windowed = 0.0;       // fraction of the stored entries whose gathers are served from shared memory
    int resident_checked = 0;    // 1: the probe found the first grid co-resident, -1: the grid had to shrink
    double lines_per_gather = 32.0;   // 128-byte lines that 32 consecutive stored entries span (32 = no locality)
    int gather_bound = 0;        // kernel flavour: 1 = random columns (FLAV_GATHER), 0 = local gathers (FLAV_LOCAL)
    unsigned long long *stat = nullptr;   // device scratch of the locality measurement
    uint32_t span_p50 = 0, span_max = 0;   // gather span (entries of the dense vector) of the pieces: median, maximum
};

static inline void plan_free(TilePlan &p)
{
    if (p.tiles) cudaFree(p.tiles);
    if (p.order) cudaFree(p.order);
    if (p.stat) cudaFree(p.stat);
    p.tiles = nullptr; p.order = nullptr; p.stat = nullptr;
}

constexpr int kEpl = 4;           // stored entries per lane and chunk of the instantiated flavour (spmv.cuh: ChunkRegs)
constexpr int kWinCap = 304;      // widest gather window: 4 CTAs x 8 warps x (128 + 304) doubles + static = 130 KB <= the 132 KB step

static inline size_t plan_smem(int win_cap) { return (size_t)kWWarps * (size_t)(32 * kEpl + win_cap) * sizeof(double); }

// Shared-memory carve-out of the SpMV kernels.  What is left of the 228 KB is L1, and the divergent gathers of the
// non-windowed path live on L1 (every pending miss holds a line: with the carve-out at 100 % the same kernel ran at
// HALF speed on every workload, profiles/r02/run1), so the carve-out is the smallest hardware step that holds the 4
// resident CTAs: 64 KB without gather windows, 132 KB with them.  It is a property of the kernel FUNCTION, i.e.
// process-wide state, and a guarded multi-block launch deadlocks on a grid that is only partly resident -- which is
// what a carve-out that changes between the kernels of one solve produces (an SM cannot be re-split while CTAs of
// the previous kernel sit on it; runs 2 and 5 hung / timed out exactly there).  Hence ONE value for every SpMV
// kernel, raised once -- and for good -- when the first windowed plan of the process is built.
static int spmv_configure(bool windows, int *ctas_per_sm)
{
    static std::mutex mu;
    static int current_pct = -1;
    std::lock_guard<std::mutex> lock(mu);
    const int want = windows ? 58 : 29;            // 132 KB / 64 KB of 228 KB
    int pct = std::max(current_pct, want);
    pct = std::max(0, std::min(env_int("LSQR_B200_SMEM_CARVEOUT_PCT", pct), 100));
    int occ = 4;
    auto one = [&](auto kernel) -> int {
        if (pct != current_pct) {
            LSQRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem(kWinCap)));
            LSQRB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
        int n = 0;
        LSQRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kWThreads, plan_smem(windows ? kWinCap : 0)));
        occ = std::min(occ, n);
        return LSQR_B200_OK;
    };
    LSQRB_TRY(one(spmv_kernel<FIN_NONE, kEpl, FLAV_LOCAL>));
    LSQRB_TRY(one(spmv_kernel<FIN_APROD, kEpl, FLAV_LOCAL>));
    LSQRB_TRY(one(spmv_kernel<FIN_ATPROD, kEpl, FLAV_LOCAL>));
    LSQRB_TRY(one(spmv_kernel<FIN_INIT_ATPROD, kEpl, FLAV_LOCAL>));
    LSQRB_TRY(one(spmv_kernel<FIN_PUSH, kEpl, FLAV_LOCAL>));
    LSQRB_TRY(one(spmv_kernel<FIN_NONE, kEpl, FLAV_WINDOW>));
    LSQRB_TRY(one(spmv_kernel<FIN_APROD, kEpl, FLAV_WINDOW>));
    LSQRB_TRY(one(spmv_kernel<FIN_ATPROD, kEpl, FLAV_WINDOW>));
    LSQRB_TRY(one(spmv_kernel<FIN_INIT_ATPROD, kEpl, FLAV_WINDOW>));
    LSQRB_TRY(one(spmv_kernel<FIN_PUSH, kEpl, FLAV_WINDOW>));
    LSQRB_TRY(one(spmv_kernel<FIN_NONE, kEpl, FLAV_GATHER>));
    LSQRB_TRY(one(spmv_kernel<FIN_APROD, kEpl, FLAV_GATHER>));
    LSQRB_TRY(one(spmv_kernel<FIN_ATPROD, kEpl, FLAV_GATHER>));
    LSQRB_TRY(one(spmv_kernel<FIN_INIT_ATPROD, kEpl, FLAV_GATHER>));
    LSQRB_TRY(one(spmv_kernel<FIN_PUSH, kEpl, FLAV_GATHER>));
    current_pct = pct;
    if (occ < 1) { set_last_error("spmv kernel does not fit an SM"); return LSQR_B200_ERR_CUDA; }
    *ctas_per_sm = occ;
    return LSQR_B200_OK;
}

// Cost of a piece in units of one 128-entry chunk: its chunks plus its row-window reloads.
static inline double piece_cost(const TileDesc &a, const TileDesc &b)
{
    if (a.row == b.row) return 0.0;   // no row starts here: skipped by the kernel
    return std::ceil((double)(b.entry - a.entry + 3u) / 128.0) + 0.25 * std::ceil((double)(b.row - a.row) / 32.0) + 1.0;
}

// Cuts the rows of M (all blocks) into tiles for a persistent grid of `ctas` CTAs.
static int plan_cut(Work &wk, const Csr &M, TilePlan *p, uint64_t forced_tile)
{
    if (p->tiles) { cudaFree(p->tiles); p->tiles = nullptr; }
    const int64_t nb = M.nblocks, nrows = M.nkeys;
    const int64_t work = M.nnz + (int64_t)p->row_w * nrows * nb;
    const int64_t nwarps = (int64_t)p->ctas * kWWarps;
    uint64_t tile = forced_tile;
    if (tile == 0) {
        // every warp gets the same number k of tiles; a piece (one tile in one block) holds at most ~8K work units
        const int forced = env_int("LSQR_B200_WARP_TILE", 0);
        const int64_t piece_cap = forced >= 128 ? forced : 8192;
        const int64_t k = std::max<int64_t>(1, (work / nb + nwarps * piece_cap - 1) / (nwarps * piece_cap));
        const int64_t t = (work + nwarps * k - 1) / (nwarps * k);
        tile = (uint64_t)std::max<int64_t>(512 * nb, (t + 3) & ~(int64_t)3);
    }
    p->tile = tile;
    p->nblocks = (int)nb;
    const int64_t nt = std::max<int64_t>(1, (work + (int64_t)tile - 1) / (int64_t)tile);
    p->ntiles = (int)nt;
    LSQRB_CUDA(cudaMalloc(&p->tiles, sizeof(TileDesc) * (size_t)nb * (size_t)(nt + 1)));
    build_tiles_kernel<<<(int)((nt + 1 + 255) / 256), 256, 0, wk.stream>>>(M.ptr, nrows, (int)nb, (int)nt, tile, p->row_w, p->tiles);
    LSQRB_CUDA(cudaGetLastError());
    const int64_t np = nb * (nt + 1);
    if (!p->stat) LSQRB_CUDA(cudaMalloc(&p->stat, 2 * sizeof(unsigned long long)));
    LSQRB_CUDA(cudaMemsetAsync(p->stat, 0, 2 * sizeof(unsigned long long), wk.stream));
    tile_span_kernel<<<(int)std::min<int64_t>((np * 32 + 255) / 256, (int64_t)wk.sms * 16), 256, 0, wk.stream>>>(M.idx, p->tiles, np, (int)nt, p->stat);
    LSQRB_CUDA(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    LSQRB_CUDA(cudaMemcpyAsync(h, p->stat, sizeof h, cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    p->lines_per_gather = h[1] ? (double)h[0] / (double)h[1] : 32.0;
    return LSQR_B200_OK;
}

static int plan_fetch(Work &wk, const TilePlan &p, std::vector<TileDesc> *t)
{
    t->resize((size_t)p.nblocks * ((size_t)p.ntiles + 1));
    LSQRB_CUDA(cudaMemcpyAsync(t->data(), p.tiles, sizeof(TileDesc) * t->size(), cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

// Tiles are row-aligned, so a row of 10 000 entries makes a tile of more than 10 000: with round-robin assignment the
// most loaded warp of a power-law matrix (C4) carries ~1.5x the mean and the whole grid waits for it.  When that
// happens the matrix is re-cut into smaller tiles and the tiles are dealt to the warps by LPT (largest first, to the
// least loaded warp).  The schedule is a pure function of ptr[], so results stay reproducible run to run.
static int plan_balance(Work &wk, const Csr &M, TilePlan *p, std::vector<TileDesc> *t)
{
    const int nw = p->ctas * kWWarps;
    const size_t stride = (size_t)p->ntiles + 1;
    std::vector<double> cost;
    double total = 0.0;
    auto costs = [&]() {
        const size_t st = (size_t)p->ntiles + 1;
        cost.assign((size_t)p->ntiles, 0.0);
        total = 0.0;
        for (int b = 0; b < p->nblocks; ++b)
            for (int i = 0; i < p->ntiles; ++i) {
                const double c = piece_cost((*t)[b * st + i], (*t)[b * st + i + 1]);
                cost[(size_t)i] += c;
                total += c;
            }
    };
    (void)stride;
    costs();
    {   // round robin: tile i belongs to warp i mod (warps of the grid)
        const int gw = std::max(1, std::min((p->ntiles + kWWarps - 1) / kWWarps, p->ctas)) * kWWarps;
        std::vector<double> load((size_t)gw, 0.0);
        for (int i = 0; i < p->ntiles; ++i) load[(size_t)(i % gw)] += cost[(size_t)i];
        p->imbalance = total > 0 ? *std::max_element(load.begin(), load.end()) / (total / nw) : 1.0;
    }
    const int64_t work = M.nnz + (int64_t)p->row_w * M.nkeys * M.nblocks;
    if (env_int("LSQR_B200_BALANCE", 1) == 0 || work <= (int64_t)nw * 512 * M.nblocks) return LSQR_B200_OK;   // nothing to deal out
    if (p->imbalance <= 1.0 + 1e-3 * env_int("LSQR_B200_BALANCE_PERMILLE", 60)) return LSQR_B200_OK;

    // finer tiles pack better (the long rows stay as long as they are)
    const uint64_t piece = p->tile / (uint64_t)M.nblocks;
    const uint64_t fine_piece = (uint64_t)std::max<int64_t>(512, std::min<int64_t>((int64_t)(piece / 4), env_int("LSQR_B200_BALANCE_TILE", 2048))) & ~3ull;
    const uint64_t fine = fine_piece * (uint64_t)M.nblocks;
    if (fine < p->tile) {
        LSQRB_TRY(plan_cut(wk, M, p, fine));
        LSQRB_TRY(plan_fetch(wk, *p, t));
        costs();
    }
    const int nt = p->ntiles;
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
    p->nslots = (int)order.size();
    p->imbalance = total > 0 ? worst / (total / nw) : 1.0;
    if (p->order) { cudaFree(p->order); p->order = nullptr; }
    LSQRB_CUDA(cudaMalloc(&p->order, sizeof(uint32_t) * std::max<size_t>(order.size(), 1)));
    LSQRB_CUDA(cudaMemcpyAsync(p->order, order.data(), sizeof(uint32_t) * order.size(), cudaMemcpyHostToDevice, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

// Decides the gather window from the spans of the pieces: fraction of the stored entries that a window of `cap`
// doubles would serve, and the largest span among those pieces.
static void window_stats(const TilePlan &p, const std::vector<TileDesc> &t, uint32_t cap, double *frac, uint32_t *need)
{
    const size_t st = (size_t)p.ntiles + 1;
    double in = 0.0, all = 0.0;
    uint32_t mx = 0;
    for (int b = 0; b < p.nblocks; ++b)
        for (int i = 0; i < p.ntiles; ++i) {
            const TileDesc &d = t[b * st + i], &e = t[b * st + i + 1];
            const double n = (double)(e.entry - d.entry);
            all += n;
            if (n > 0 && d.win_len <= cap) { in += n; mx = std::max(mx, d.win_len); }
        }
    *frac = all > 0 ? in / all : 0.0;
    *need = mx;
}

static int plan_probe(Work &wk, const TilePlan &P, bool *resident);

static int build_plan(Work &wk, const Csr &M, TilePlan *p, int reserve_sms = 0)
{
    p->row_w = (uint32_t)std::max(0, env_int("LSQR_B200_TILE_ROW_WEIGHT", 4));
    p->epl = kEpl;
    p->win_cap = 0;
    std::vector<TileDesc> t;
    // Gather windows are OPT-IN (LSQR_B200_WINDOW=1).  Measured on the banded family (C3, profiles/r02): Aprod 1.27 ms
    // with the window against 1.30 ms without, A' (span 1300 entries, 16 warps per SM) 1.74 ms against 1.38 ms -- the
    // staged reads are bank-conflict-bound (17 wavefronts per 32 gathers, ncu) where the global gathers of such a
    // matrix need ~20 lines, and a second carve-out in the process is a hazard for the guarded launches (above).
    const int want_window = env_int("LSQR_B200_WINDOW", 0);
    const double min_frac = 1e-2 * env_int("LSQR_B200_WINDOW_MIN_PERCENT", 50);
    int occ = 0;
    LSQRB_TRY(spmv_configure(false, &occ));
    p->smem = plan_smem(0);
    p->ctas = std::max(1, wk.sms - reserve_sms) * occ;
    LSQRB_TRY(plan_cut(wk, M, p, 0));
    LSQRB_TRY(plan_fetch(wk, *p, &t));
    // Kernel flavour: a warp-wide gather of a random-column matrix touches ~25-32 lines (C2, C4, C5), of a banded one
    // ~10 (C3); the two want opposite instruction orders in the chunk loop (spmv.cuh, warp_chunk_core).
    {
        const int forced = env_int("LSQR_B200_FLAVOUR", -1);
        p->gather_bound = forced >= 0 ? (forced == 2) : (p->lines_per_gather >= (double)env_int("LSQR_B200_GATHER_LINES", 20));
    }
    const int forced_cap = env_int("LSQR_B200_WINDOW_CAP", 0);
    if (M.nnz > 0 && want_window) {
        double f = 0;
        uint32_t need = 0;
        window_stats(*p, t, (uint32_t)kWinCap, &f, &need);
        if (forced_cap > 0) p->win_cap = std::min((forced_cap + 1) & ~1, kWinCap);
        else if (f >= min_frac) p->win_cap = (int)((need + 1u) & ~1u);
    }
    if (p->win_cap > 0) {
        LSQRB_TRY(spmv_configure(true, &occ));
        p->smem = plan_smem(p->win_cap);
        const int ctas = std::max(1, wk.sms - reserve_sms) * occ;
        if (ctas != p->ctas) {          // a different persistent grid: cut again for it
            p->ctas = ctas;
            LSQRB_TRY(plan_cut(wk, M, p, 0));
            LSQRB_TRY(plan_fetch(wk, *p, &t));
        }
    }
    if (M.nblocks > 1) {
        // the guarded multi-block launch needs the whole persistent grid resident at once: prove it, or shrink the grid
        for (int tries = 0; tries < 3; ++tries) {
            bool resident = false;
            LSQRB_TRY(plan_probe(wk, *p, &resident));
            if (resident) break;
            const int per_sm = std::max(1, p->ctas / std::max(1, wk.sms - reserve_sms) - 1);
            if (env_int("LSQR_B200_VERBOSE", 0)) fprintf(stderr, "[lsqr_b200] persistent grid of %d CTAs is not co-resident: %d CTAs per SM\n", p->ctas, per_sm);
            p->ctas = std::max(1, wk.sms - reserve_sms) * per_sm;
            LSQRB_TRY(plan_cut(wk, M, p, 0));
            LSQRB_TRY(plan_fetch(wk, *p, &t));
            p->resident_checked = -1;
        }
        if (p->resident_checked == 0) p->resident_checked = 1;
    }
    {   // span statistics (diagnostic) and the final window assignment
        std::vector<uint32_t> spans;
        const size_t st = (size_t)p->ntiles + 1;
        for (int b = 0; b < p->nblocks; ++b)
            for (int i = 0; i < p->ntiles; ++i)
                if (t[b * st + i + 1].entry > t[b * st + i].entry) spans.push_back(t[b * st + i].win_len);
        if (!spans.empty()) {
            std::nth_element(spans.begin(), spans.begin() + spans.size() / 2, spans.end());
            p->span_p50 = spans[spans.size() / 2];
            p->span_max = *std::max_element(spans.begin(), spans.end());
        }
        uint32_t need = 0;
        window_stats(*p, t, (uint32_t)p->win_cap, &p->windowed, &need);
        if (p->win_cap == 0) p->windowed = 0.0;
        const int64_t np = (int64_t)p->nblocks * ((int64_t)p->ntiles + 1);
        tile_window_cap_kernel<<<(int)((np + 255) / 256), 256, 0, wk.stream>>>(p->tiles, np, (uint32_t)p->win_cap);
        LSQRB_CUDA(cudaGetLastError());
        for (auto &d : t) if (d.win_len > (uint32_t)p->win_cap) { d.win_len = 0; d.win_lo = 0; }
    }
    LSQRB_TRY(plan_balance(wk, M, p, &t));
    {   // the balanced schedule may have re-cut the tiles: their windows are capped (again) here, whatever path was taken
        const int64_t np = (int64_t)p->nblocks * ((int64_t)p->ntiles + 1);
        tile_window_cap_kernel<<<(int)((np + 255) / 256), 256, 0, wk.stream>>>(p->tiles, np, (uint32_t)p->win_cap);
        LSQRB_CUDA(cudaGetLastError());
    }
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
struct ProductIo {
    const double *x = nullptr;   // gathered vector
    double *out = nullptr;       // FINAL destination (u / v), or the destination of a plain product
    double *part = nullptr;      // partial sums across blocks (gu / g)
    int first_mode = BM_STORE;
    Ssq *aux = nullptr;
    int check_done = 0;
    double *const *push = nullptr;
    int64_t push_cols = 0;
    const PeerView *peer = nullptr;
    int pdl = 0;                 // programmatic dependent launch behind the previous kernel of the stream
};

template <int FIN>
static int launch_piece(Work &wk, const TilePlan &P, const SpmvArgs &a)
{
    const int grid = P.order ? P.ctas : std::max(1, std::min((P.ntiles + kWWarps - 1) / kWWarps, P.ctas));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kWThreads);
    cfg.dynamicSmemBytes = P.smem;
    cfg.stream = wk.stream;
    // (a guarded multi-block launch is a grid barrier per block: its grid was PROVEN co-resident by the probe at
    // initialize; a cooperative launch would be the textbook guarantee, but the runtime refused these grids -- "too
    // many blocks in cooperative launch" -- although the occupancy API reports the same 4 CTAs per SM)
    cudaLaunchAttribute attr[1];
    if (a.pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    if (P.win_cap > 0)      LSQRB_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<FIN, kEpl, FLAV_WINDOW>, a));
    else if (P.gather_bound) LSQRB_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<FIN, kEpl, FLAV_GATHER>, a));
    else                    LSQRB_CUDA(cudaLaunchKernelEx(&cfg, spmv_kernel<FIN, kEpl, FLAV_LOCAL>, a));
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;
}

// Launches the flavour's kernel in probe mode on the plan's grid: true if every CTA saw every other one arrive.
static int plan_probe(Work &wk, const TilePlan &P, bool *resident)
{
    *resident = false;
    LSQRB_CUDA(cudaMemsetAsync(&wk.st->probe_count, 0, sizeof(unsigned int) + sizeof(int), wk.stream));
    SpmvArgs a;
    memset(&a, 0, sizeof a);
    a.probe = 1;
    a.st = wk.st;
    TilePlan Q = P;
    Q.order = reinterpret_cast<uint32_t *>(1);      // (forces the full persistent grid; never dereferenced in probe mode)
    LSQRB_TRY(launch_piece<FIN_ATPROD>(wk, Q, a));
    wk.launches--;
    int fail = 1;
    LSQRB_CUDA(cudaMemcpyAsync(&fail, &wk.st->probe_fail, sizeof(int), cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    *resident = fail == 0;
    return LSQR_B200_OK;
}

// One product over all blocks of M.  FIN != FIN_NONE: the last block applies the fused epilogue.  single = one
// persistent launch walks every block; otherwise one launch per block (A/B switch, and the fallback for matrices
// with more blocks than the drift counters hold).
template <int FIN>
static int launch_product(Work &wk, const Csr &M, const TilePlan &P, const ProductIo &io, bool single, int guard)
{
    SpmvArgs a;
    memset(&a, 0, sizeof a);        // (probe = 0 in particular)
    a.idx = M.idx; a.val = M.val;
    a.ptr_stride = M.nkeys;
    a.ntiles = P.ntiles;
    a.order = P.order; a.nslots = P.nslots;
    a.nrows = M.nkeys;
    a.x = io.x; a.out = io.out; a.part = io.part;
    a.win_cap = P.win_cap;
    a.check_done = io.check_done;
    a.st = wk.st; a.aux = io.aux;
    a.push = io.push; a.push_cols = io.push_cols; a.peer = io.peer;
    a.final_adds_part = M.nblocks > 1;
    a.pdl = io.pdl;
    const int nb = (int)M.nblocks;
    if (single && nb <= kMaxSpmvBlocks) {
        a.ptr = M.ptr; a.tiles = P.tiles; a.nblocks = nb;
        a.first_mode = io.first_mode;
        a.last_is_final = FIN != FIN_NONE;
        a.guard = nb > 1 ? guard : 0;
        return launch_piece<FIN>(wk, P, a);
    }
    a.nblocks = 1;
    a.guard = 0;
    for (int b = 0; b < nb; ++b) {
        a.ptr = M.ptr + (int64_t)b * M.nkeys;
        a.tiles = P.tiles + (size_t)b * ((size_t)P.ntiles + 1);
        a.first_mode = b == 0 ? io.first_mode : BM_ACC;
        if (FIN != FIN_NONE && b == nb - 1) {
            a.last_is_final = 1;
            LSQRB_TRY(launch_piece<FIN>(wk, P, a));
        } else {
            a.last_is_final = 0;
            a.check_done = FIN != FIN_NONE ? 1 : io.check_done;   // part of a fused product: a no-op once the solver has stopped
            LSQRB_TRY(launch_piece<FIN_NONE>(wk, P, a));
        }
    }
    return LSQR_B200_OK;
}

}  // namespace lsqrb
