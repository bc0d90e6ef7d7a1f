// engine.cu -- host driver + C ABI of the B200 LSQR engine (include/lsqr_b200.h).
//
// One handle owns: CSR(A) and CSR(A') in HBM with their work plans, the work vectors u(m), v,w,x,(se)(n), a
// device-resident scalar state (DevState), and a ring of per-iteration records in pinned mapped
// host memory.  The host never computes a scalar of the recurrence: it enqueues batches of
// iterations (CUDA graph), and reads the records to learn when the device decided to stop.
#include "plan.cuh"

using namespace lsqrb;

// =============================================================================================
// the ez handle
// =============================================================================================
struct lsqr_b200_ez {
    Work wk;
    int32_t m = 0, n = 0;
    int64_t nnz = 0;
    Csr A, AT;
    TilePlan planA, planAT;       // one plan per stored matrix (all its blocks)
    bool single_launch = true;    // one persistent launch per product walks every block (else one launch per block)
    int guard = 1;                // multi-block launch: block b starts when every warp has finished block b - guard (0 = no guard)
    bool a_blocked = false;       // A is column-blocked (v does not fit in L2)
    bool at_blocked = false;      // A' is row-blocked (u does not fit in L2)
    double *gu = nullptr;         // [m] partial A v across the column blocks
    bool overlap_update = true;   // the x/w update of iteration k runs on a side stream next to the Aprod of iteration
                                  // k+1 (they touch disjoint vectors)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false;    // an update is in flight on the side stream that the main stream has not joined
    lsqr_b200_options opt;
    double *u = nullptr, *v = nullptr, *w = nullptr, *x = nullptr, *se = nullptr;
    double *g = nullptr;          // [n + 3] partial A'u across the row blocks; multi-GPU (NCCL path): [ A_p'u_p | Ssq(u_p) ]
    double *tmp_m = nullptr, *tmp_n = nullptr;   // staging for lsqr_b200_ez_aprod with host vectors
    // ---- multi-GPU
    ncclComm_t comm = nullptr;
    bool peer = false;            // exchange over NVLink peer memory (peer.cuh) instead of one NCCL all-reduce per iteration
    void *sym_local = nullptr;    // this rank's symmetric exchange block (cudaMalloc'ed, IPC-exported)
    std::vector<void *> sym_peer; // every rank's block mapped into this process ([rank] = sym_local)
    PeerView pv;                  // host copy of the device view
    PeerView *pv_dev = nullptr;
    double **push_dev = nullptr;  // [world] where this rank's contributions to rank q's columns go
    int64_t slice0 = 0, slice_len = 0;   // owned columns [slice0, slice0 + slice_len)
    double *xs = nullptr, *ws = nullptr, *ses = nullptr;   // owned slices of x, w, se (peer path); xs is [world * cols] for the final gather
    // ---- batches
    int batch = 8;                // iterations per enqueue (per CUDA-graph launch)
    cudaGraphExec_t graph_exec = nullptr;
    int graph_wantse = -1;
    int64_t graph_launches = 0;   // kernels of this library inside one graph launch
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
    for (size_t q = 0; q < me->sym_peer.size(); ++q)
        if (me->sym_peer[q] && me->sym_peer[q] != me->sym_local) cudaIpcCloseMemHandle(me->sym_peer[q]);
    if (me->comm) { NcclApi *a = nccl_api(); if (a) a->CommDestroy(me->comm); }
    if (me->sym_local) cudaFree(me->sym_local);
    else if (me->v) cudaFree(me->v);              // (peer path: v lives inside the symmetric block)
    if (me->pv_dev) cudaFree(me->pv_dev);
    if (me->push_dev) cudaFree(me->push_dev);
    csr_free(&me->A);
    csr_free(&me->AT);
    plan_free(me->planA);
    plan_free(me->planAT);
    for (double *p : {me->u, me->w, me->x, me->se, me->g, me->gu, me->tmp_m, me->tmp_n, me->xs, me->ws, me->ses}) if (p) cudaFree(p);
    for (auto e : {me->ev_t0, me->ev_t1, me->ev_t2, me->ev_fork, me->ev_join}) if (e) cudaEventDestroy(e);
    if (me->side) cudaStreamDestroy(me->side);
    for (auto e : me->prof_ev) cudaEventDestroy(e);
    me->wk.destroy();
    delete me;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU peer path: the symmetric exchange block and its IPC mapping
// ---------------------------------------------------------------------------------------------
// Layout of every rank's block (byte offsets are the same on all ranks):
//   [ recv: world * cols doubles ][ v: n doubles (+ pad) ][ sc: world * kScDoubles doubles ][ flag1: world ][ flag2: world ]
struct SymLayout {
    size_t off_recv, off_v, off_sc, off_f1, off_f2, bytes;
};
static SymLayout sym_layout(int world, int64_t cols, int64_t n)
{
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    SymLayout L;
    L.off_recv = 0;
    L.off_v = up(sizeof(double) * (size_t)world * (size_t)cols);
    L.off_sc = L.off_v + up(sizeof(double) * (size_t)(n + 2));
    L.off_f1 = L.off_sc + up(sizeof(double) * (size_t)world * kScDoubles);
    L.off_f2 = L.off_f1 + up(sizeof(unsigned int) * (size_t)world);
    L.bytes = L.off_f2 + up(sizeof(unsigned int) * (size_t)world);
    return L;
}

// Every rank allocates its block, the IPC handles travel through one NCCL all-gather, and every rank maps the
// blocks of its peers.  Returns false (on every rank alike) if any rank could not set the mapping up: the caller
// then keeps the NCCL all-reduce path.
static int peer_setup(lsqr_b200_ez *me, bool *ok_out)
{
    *ok_out = false;
    NcclApi *api = nccl_api();
    Work &wk = me->wk;
    const int world = me->opt.world_size, rank = me->opt.rank;
    if (!api || !api->AllGather || world > kMaxRanks) return LSQR_B200_OK;
    const int64_t n = me->n;
    const int64_t cols = (((n + world - 1) / world) + 1) & ~(int64_t)1;     // even: 16-byte aligned slices
    const SymLayout L = sym_layout(world, cols, n);
    int good = 1;
    void *block = nullptr;
    if (cudaMalloc(&block, L.bytes) != cudaSuccess) { cudaGetLastError(); good = 0; }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (good && cudaIpcGetMemHandle(&mine, block) != cudaSuccess) { cudaGetLastError(); good = 0; }
    if (good) cudaMemsetAsync(block, 0, L.bytes, wk.stream);
    // exchange: [ handle (64 bytes) | good (8 bytes) ] per rank
    constexpr size_t kRec = sizeof(cudaIpcMemHandle_t) + 8;
    unsigned char *d_all = nullptr;
    LSQRB_CUDA(cudaMalloc(&d_all, kRec * (size_t)world));
    std::vector<unsigned char> h_all(kRec * (size_t)world, 0);
    memcpy(h_all.data() + kRec * (size_t)rank, &mine, sizeof mine);
    h_all[kRec * (size_t)rank + sizeof mine] = (unsigned char)good;
    LSQRB_CUDA(cudaMemcpyAsync(d_all + kRec * (size_t)rank, h_all.data() + kRec * (size_t)rank, kRec, cudaMemcpyHostToDevice, wk.stream));
    LSQRB_NCCL(api->AllGather(d_all + kRec * (size_t)rank, d_all, kRec, ncclChar, me->comm, wk.stream));
    LSQRB_CUDA(cudaMemcpyAsync(h_all.data(), d_all, h_all.size(), cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    for (int q = 0; q < world; ++q) good = good && h_all[kRec * (size_t)q + sizeof mine];
    me->sym_peer.assign((size_t)world, nullptr);
    if (good) {
        for (int q = 0; q < world && good; ++q) {
            if (q == rank) { me->sym_peer[(size_t)q] = block; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, h_all.data() + kRec * (size_t)q, sizeof h);
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = 0; }
            me->sym_peer[(size_t)q] = p;
        }
    }
    // second round: did every rank map every peer?
    h_all[0] = (unsigned char)good;
    LSQRB_CUDA(cudaMemcpyAsync(d_all + rank, h_all.data(), 1, cudaMemcpyHostToDevice, wk.stream));
    LSQRB_NCCL(api->AllGather(d_all + rank, d_all, 1, ncclChar, me->comm, wk.stream));
    LSQRB_CUDA(cudaMemcpyAsync(h_all.data(), d_all, (size_t)world, cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    cudaFree(d_all);
    for (int q = 0; q < world; ++q) good = good && h_all[(size_t)q];
    if (!good) {
        for (int q = 0; q < world; ++q)
            if (me->sym_peer[(size_t)q] && q != rank) cudaIpcCloseMemHandle(me->sym_peer[(size_t)q]);
        me->sym_peer.clear();
        if (block) cudaFree(block);
        return LSQR_B200_OK;
    }
    me->sym_local = block;
    PeerView &pv = me->pv;
    memset(&pv, 0, sizeof pv);
    pv.world = world; pv.rank = rank; pv.cols = cols; pv.n = n;
    std::vector<double *> push((size_t)world);
    for (int q = 0; q < world; ++q) {
        char *base = (char *)me->sym_peer[(size_t)q];
        pv.recv[q] = (double *)(base + L.off_recv);
        pv.v[q] = (double *)(base + L.off_v);
        pv.sc[q] = (double *)(base + L.off_sc);
        pv.flag1[q] = (unsigned int *)(base + L.off_f1);
        pv.flag2[q] = (unsigned int *)(base + L.off_f2);
        push[(size_t)q] = pv.recv[q] + (size_t)rank * (size_t)cols;
    }
    LSQRB_CUDA(cudaMalloc(&me->pv_dev, sizeof(PeerView)));
    LSQRB_CUDA(cudaMemcpyAsync(me->pv_dev, &pv, sizeof pv, cudaMemcpyHostToDevice, wk.stream));
    LSQRB_CUDA(cudaMalloc(&me->push_dev, sizeof(double *) * (size_t)world));
    LSQRB_CUDA(cudaMemcpyAsync(me->push_dev, push.data(), sizeof(double *) * (size_t)world, cudaMemcpyHostToDevice, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    me->v = pv.v[rank];
    me->slice0 = std::min<int64_t>(n, (int64_t)rank * cols);
    me->slice_len = std::max<int64_t>(0, std::min<int64_t>(cols, n - me->slice0));
    *ok_out = true;
    return LSQR_B200_OK;
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
        // Column-blocked A: when v (8 n bytes) cannot stay in L2 while A streams past it, A is stored as one CSR per
        // block of COLUMNS, so that every block gathers from a slice of v that does fit.
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
        // Row-blocked transpose: the mirror image for u (8 m bytes) and A'.
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

    if (me->opt.spmv_variant != 0 && me->opt.spmv_variant != 3) {
        set_last_error("spmv_variant: only the warp-autonomous segmented kernel (0 / 3) exists");
        return LSQR_B200_ERR_ARG;
    }
    me->a_blocked = me->A.nblocks > 1;
    me->at_blocked = me->AT.nblocks > 1;
    me->single_launch = env_int("LSQR_B200_SINGLE_LAUNCH", 1) != 0;
    me->guard = std::max(0, std::min(env_int("LSQR_B200_DRIFT_GUARD", 1), 8));
    me->overlap_update = me->opt.world_size == 1 && env_int("LSQR_B200_OVERLAP_UPDATE", 1) != 0;
    me->wk.pdl = env_int("LSQR_B200_PDL", 0) != 0;
    if (me->overlap_update) {
        LSQRB_CUDA(cudaStreamCreateWithFlags(&me->side, cudaStreamNonBlocking));
        LSQRB_CUDA(cudaEventCreateWithFlags(&me->ev_fork, cudaEventDisableTiming));
        LSQRB_CUDA(cudaEventCreateWithFlags(&me->ev_join, cudaEventDisableTiming));
    }
    // work plans: row cuts shared by all blocks, gather windows, kernel flavour, balanced schedule
    LSQRB_TRY(build_plan(wk, me->A, &me->planA));
    LSQRB_TRY(build_plan(wk, me->AT, &me->planAT));
    if (env_int("LSQR_B200_VERBOSE", 0)) {
        auto say = [&](const char *name, const Csr &M, const TilePlan &P) {
            fprintf(stderr, "[lsqr_b200] %s: blocks=%lld (block size %lld) tiles=%d (%llu work units) grid=%d CTAs x %d/SM  window=%d doubles "
                            "(%.1f%% of the entries staged; piece span median %u, max %u) lines/gather %.1f%s%s imbalance %.3f\n",
                    name, (long long)M.nblocks, (long long)M.block_rows, P.ntiles, (unsigned long long)P.tile, P.ctas, P.ctas / std::max(1, me->wk.sms), P.win_cap,
                    100.0 * P.windowed, P.span_p50, P.span_max, P.lines_per_gather, P.win_cap > 0 ? " [window flavour]" : (P.gather_bound ? " [gather-bound flavour]" : " [local flavour]"), P.order ? " LPT" : "", P.imbalance);
        };
        fprintf(stderr, "[lsqr_b200] m=%d n=%d nnz=%lld single_launch=%d guard=%d\n", me->m, me->n, (long long)me->nnz, (int)me->single_launch, (int)me->guard);
        say("A ", me->A, me->planA);
        say("A'", me->AT, me->planAT);
    }

    const size_t mm = (size_t)std::max<int32_t>(me->m, 1), nn = (size_t)std::max<int32_t>(me->n, 1);
    LSQRB_CUDA(cudaMalloc(&me->u, sizeof(double) * (mm + 2)));
    if (me->a_blocked) LSQRB_CUDA(cudaMalloc(&me->gu, sizeof(double) * (mm + 2)));
    if (me->opt.world_size > 1 || me->at_blocked) LSQRB_CUDA(cudaMalloc(&me->g, sizeof(double) * (nn + 4)));
    if (me->opt.world_size > 1) {
        const int want = env_int("LSQR_B200_PEER_EXCHANGE", 1);
        bool ok = false;
        if (want) LSQRB_TRY(peer_setup(me, &ok));
        me->peer = ok;
        if (env_int("LSQR_B200_VERBOSE", 0))
            fprintf(stderr, "[lsqr_b200] rank %d/%d: exchange over %s\n", me->opt.rank, me->opt.world_size,
                    me->peer ? "NVLink peer memory (IPC-mapped symmetric blocks)" : "one NCCL all-reduce per iteration");
    }
    if (!me->peer) {
        LSQRB_CUDA(cudaMalloc(&me->v, sizeof(double) * (nn + 2)));
        LSQRB_CUDA(cudaMalloc(&me->w, sizeof(double) * (nn + 2)));
        LSQRB_CUDA(cudaMalloc(&me->x, sizeof(double) * (nn + 2)));
    } else {
        const size_t sl = (size_t)std::max<int64_t>(me->pv.cols, 1);
        LSQRB_CUDA(cudaMalloc(&me->ws, sizeof(double) * sl));
        LSQRB_CUDA(cudaMalloc(&me->xs, sizeof(double) * sl * (size_t)me->opt.world_size));   // gathered in place at the end
    }
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
    {   // iterations per enqueue (= per CUDA-graph launch): enough to cover ~1 ms of device time.  Consecutive graph
        // launches leave ~15 us between them (2 us between the kernels inside a graph), which a small problem feels
        // (C2, 125 us per iteration: 131 / 123 / 121 us with 1 / 3 / 8 iterations per graph, profiles/r02/run19);
        // iterations enqueued past the stop are empty kernels, so not more than 8
        const double est_iter_us = 24.0 * (double)size_a / 2.5e6 + 20.0;   // ~2.5 TB/s effective + launch floor
        int dflt = (int)std::ceil(1000.0 / est_iter_us);
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
    const Csr &M = which ? me->AT : me->A;
    const TilePlan &P = which ? me->planAT : me->planA;
    if (block < 0 || block >= M.nblocks) { set_last_error("no such block"); return LSQR_B200_ERR_ARG; }
    // one plan covers every block of the matrix: the same row cuts, the same warp for a tile in every block
    if (ntiles) *ntiles = P.ntiles;
    if (tile_entries) *tile_entries = (int64_t)(P.tile / (uint64_t)std::max<int64_t>(M.nblocks, 1));
    if (balanced) *balanced = P.order != nullptr;
    if (imbalance) *imbalance = P.imbalance;
    return LSQR_B200_OK;
}

int lsqr_b200_ez_plan(const lsqr_b200_ez *me, int32_t which, lsqr_b200_plan_info *out)
{
    if (!me || !out || (which != 0 && which != 1)) return LSQR_B200_ERR_ARG;
    const Csr &M = which ? me->AT : me->A;
    const TilePlan &P = which ? me->planAT : me->planA;
    memset(out, 0, sizeof *out);
    out->nblocks = M.nblocks;
    out->ntiles = P.ntiles;
    out->grid_ctas = P.ctas;
    out->ctas_per_sm = P.ctas / std::max(1, me->wk.sms);
    out->window_doubles = P.win_cap;
    out->windowed_fraction = P.windowed;
    out->span_median = P.span_p50;
    out->span_max = P.span_max;
    out->balanced = P.order != nullptr;
    out->imbalance = P.imbalance;
    out->single_launch = me->single_launch && M.nblocks <= kMaxSpmvBlocks;
    out->peer_exchange = me->peer;
    out->entries_per_lane = P.epl;
    out->lines_per_gather = P.lines_per_gather;
    out->flavour = P.win_cap > 0 ? 1 : (P.gather_bound ? 2 : 0);
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
    if (mode != 1 && mode != 2) return LSQR_B200_ERR_MODE;       // :197
    Work &wk = me->wk;
    // `stream` is an ordinary cudaStream_t argument: NULL is the legacy default stream (as everywhere in CUDA), not
    // the handle's own stream -- the caller's work on that stream is ordered before and after these launches
    cudaStream_t saved = wk.stream;
    wk.stream = stream ? (cudaStream_t)stream : cudaStreamLegacy;
    int rc = cudaSetDevice(wk.device) == cudaSuccess ? LSQR_B200_OK : LSQR_B200_ERR_CUDA;
    ProductIo io;
    io.first_mode = BM_ACC;   // every block accumulates straight into the caller's vector
    if (rc == LSQR_B200_OK) {
        if (mode == 1) {                                                                                                   // y += A x
            io.x = x_dev; io.out = y_dev; io.part = y_dev;
            rc = launch_product<FIN_NONE>(wk, me->A, me->planA, io, me->single_launch, me->guard);
        } else {                                                                                                           // x += A'y
            io.x = y_dev; io.out = x_dev; io.part = x_dev;
            rc = launch_product<FIN_NONE>(wk, me->AT, me->planAT, io, me->single_launch, me->guard);
        }
    }
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
    LSQRB_TRY(lsqr_b200_ez_aprod_device(me, mode, m, n, dx, dy, (void *)wk.stream));
    if (mode == 1 && !ydev) LSQRB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, wk.stream));
    if (mode == 2 && !xdev) LSQRB_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    return LSQR_B200_OK;
}

}  // extern "C"

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

// u' = ca_mat (A v) + ca_vec u, ||u'||: one launch; aux != NULL: the local sum of squares goes there (multi-GPU)
static int do_aprod(lsqr_b200_ez *me, Ssq *aux)
{
    ProductIo io;
    io.x = me->v; io.out = me->u; io.part = me->gu; io.first_mode = BM_STORE; io.aux = aux;
    // the A v kernel directly follows the A'u kernel of the previous iteration in the stream: with LSQR_B200_PDL its
    // launch and prologue overlap that kernel's tail, grid reduction and scalar step
    io.pdl = me->wk.pdl && me->opt.world_size == 1;
    return launch_product<FIN_APROD>(me->wk, me->A, me->planA, io, me->single_launch, me->guard);
}
template <int FIN>   // FIN_ATPROD / FIN_INIT_ATPROD: v' = ct_mat (A'u) + ct_vec v, ||v'||: one launch
static int do_atprod(lsqr_b200_ez *me)
{
    ProductIo io;
    io.x = me->u; io.out = me->v; io.part = me->g; io.first_mode = BM_STORE;
    return launch_product<FIN>(me->wk, me->AT, me->planAT, io, me->single_launch, me->guard);
}
// multi-GPU, NCCL path: g = [ A_p'u_p | Ssq(u_p) ] summed over the ranks by one all-reduce
static int do_atprod_allreduce(lsqr_b200_ez *me)
{
    ProductIo io;
    io.x = me->u; io.out = me->g; io.part = me->g; io.first_mode = BM_STORE; io.check_done = 1;
    LSQRB_TRY(launch_product<FIN_NONE>(me->wk, me->AT, me->planAT, io, me->single_launch, me->guard));
    NcclApi *api = nccl_api();
    LSQRB_NCCL(api->AllReduce(me->g, me->g, (size_t)me->n + 3, ncclFloat64, ncclSum, me->comm, me->wk.stream));
    return LSQR_B200_OK;
}
// multi-GPU, peer path: the partial A_p'u_p goes straight to the owners of the columns (peer.cuh)
static int do_atprod_push(lsqr_b200_ez *me)
{
    ProductIo io;
    io.x = me->u; io.out = nullptr; io.part = me->g; io.first_mode = BM_STORE; io.check_done = 1;
    io.push = me->push_dev; io.push_cols = me->pv.cols; io.peer = me->pv_dev;
    return launch_product<FIN_PUSH>(me->wk, me->AT, me->planAT, io, me->single_launch, me->guard);
}

static int peer_grid(const lsqr_b200_ez *me) { return me->wk.grid_for(std::max<int64_t>(me->slice_len, 1), kThreads); }

// one LSQR iteration, enqueued (no host synchronisation)
static int enqueue_iteration(lsqr_b200_ez *me, bool wantse)
{
    Work &wk = me->wk;
    if (me->peer) {
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod(me, &wk.st->usq_local)); }
        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_push(me)); }
        { ProfScope p(me, CLS_OTHER);
          peer_vfinish_kernel<false><<<peer_grid(me), kThreads, 0, wk.stream>>>(me->pv, wk.st);
          peer_step_kernel<false><<<1, 32, 0, wk.stream>>>(me->pv, wk.st);
          wk.launches += 2; LSQRB_CUDA(cudaGetLastError()); }
        { ProfScope p(me, CLS_UPDATE);
          const size_t off = (size_t)me->opt.rank * (size_t)me->pv.cols;      // this rank's slot of the gather buffers
          LSQRB_TRY(launch_update<true>(wk, me->slice_len, me->xs + off, me->ws, me->v + me->slice0,
                                        wantse ? me->ses + off : nullptr, wantse, 1)); }
        return LSQR_B200_OK;
    }
    if (me->opt.world_size > 1) {
        // NCCL path: u' and its partial norm, g = [A'u' | Ssq(u')] all-reduced over the ranks, then every rank forms
        // v' = g/beta - (beta/alpha) v, both scalar steps and the x/w update redundantly
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod(me, (Ssq *)(me->g + me->n))); }
        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod_allreduce(me)); }   // (product and all-reduce timed together)
        { ProfScope p(me, CLS_OTHER);
          vfinish_kernel<false><<<wk.grid_for(me->n, kThreads), kThreads, 0, wk.stream>>>(me->n, me->g, me->v, wk.st);
          wk.launches++; LSQRB_CUDA(cudaGetLastError()); }
        { ProfScope p(me, CLS_UPDATE); LSQRB_TRY(launch_update<true>(wk, me->n, me->x, me->w, me->v, me->se, wantse)); }
        return LSQR_B200_OK;
    }
    if (me->overlap_update) {
        // K3(k+1) reads v, writes u; K5(k) reads v, writes x, w: disjoint, so the update leaves the critical path.
        // K4 overwrites v and needs ||w||, so it joins the side stream first.
        { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod(me, nullptr)); }
        if (me->side_pending) { LSQRB_CUDA(cudaStreamWaitEvent(wk.stream, me->ev_join, 0)); me->side_pending = false; }
        { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod<FIN_ATPROD>(me)); }
        LSQRB_CUDA(cudaEventRecord(me->ev_fork, wk.stream));
        LSQRB_CUDA(cudaStreamWaitEvent(me->side, me->ev_fork, 0));
        {
            cudaStream_t main_stream = wk.stream;
            wk.stream = me->side;
            int rc;
            { ProfScope p(me, CLS_UPDATE); rc = launch_update<true>(wk, me->n, me->x, me->w, me->v, me->se, wantse); }
            wk.stream = main_stream;
            LSQRB_TRY(rc);
        }
        LSQRB_CUDA(cudaEventRecord(me->ev_join, me->side));
        me->side_pending = true;
        return LSQR_B200_OK;
    }
    { ProfScope p(me, CLS_APROD);  LSQRB_TRY(do_aprod(me, nullptr)); }
    { ProfScope p(me, CLS_ATPROD); LSQRB_TRY(do_atprod<FIN_ATPROD>(me)); }
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
    me->graph_launches = wk.launches - saved;
    wk.launches = saved;
    if (rc == LSQR_B200_OK && e == cudaSuccess) e = cudaGraphInstantiate(&me->graph_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if ((rc != LSQR_B200_OK || e != cudaSuccess) && wk.pdl) {
        // a runtime that cannot capture programmatic launch edges: capture again with ordinary edges
        cudaGetLastError();
        if (me->graph_exec) { cudaGraphExecDestroy(me->graph_exec); me->graph_exec = nullptr; }
        wk.pdl = 0;
        return build_graph(me, wantse);
    }
    if (rc != LSQR_B200_OK) { cudaGetLastError(); return rc; }   // (clears the capture's sticky error)
    LSQRB_CUDA(e);
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
    const bool peer = me->peer;
    LSQRB_CUDA(cudaSetDevice(wk.device));
    if (wantse && !peer && !me->se) LSQRB_CUDA(cudaMalloc(&me->se, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
    if (wantse && peer && !me->ses) LSQRB_CUDA(cudaMalloc(&me->ses, sizeof(double) * (size_t)std::max<int64_t>(me->pv.cols, 1) * (size_t)me->opt.world_size));
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
    double *xs = peer ? me->xs + (size_t)me->opt.rank * (size_t)me->pv.cols : nullptr;      // owned slices (peer path)
    double *ses = peer && wantse ? me->ses + (size_t)me->opt.rank * (size_t)me->pv.cols : nullptr;
    if (m > 0) LSQRB_CUDA(cudaMemcpyAsync(me->u, b, sizeof(double) * (size_t)m, cudaMemcpyDefault, wk.stream));
    if (n > 0 && !peer) {
        LSQRB_CUDA(cudaMemsetAsync(me->v, 0, sizeof(double) * (size_t)n, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(me->x, 0, sizeof(double) * (size_t)n, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(me->w, 0, sizeof(double) * (size_t)n, wk.stream));
        if (wantse) LSQRB_CUDA(cudaMemsetAsync(me->se, 0, sizeof(double) * (size_t)n, wk.stream));
    } else if (peer) {
        // (v is rewritten slice by slice by its owners in the first exchange; x, w, se are owned slices)
        LSQRB_CUDA(cudaMemsetAsync(me->xs, 0, sizeof(double) * (size_t)me->pv.cols * (size_t)me->opt.world_size, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(me->ws, 0, sizeof(double) * (size_t)me->pv.cols, wk.stream));
        if (wantse) LSQRB_CUDA(cudaMemsetAsync(me->ses, 0, sizeof(double) * (size_t)me->pv.cols * (size_t)me->opt.world_size, wk.stream));
    }
    // beta = ||u||; v = A'(u/beta); alpha = ||v||; w = v/alpha  (:632-644), lazily normalised
    if (peer) {
        nrm2_kernel<POST_SSQ><<<wk.grid_for(m, kThreads), kThreads, 0, wk.stream>>>(m, me->u, wk.st, (double *)&wk.st->usq_local);
        wk.launches++;
        LSQRB_TRY(do_atprod_push(me));
        peer_vfinish_kernel<true><<<peer_grid(me), kThreads, 0, wk.stream>>>(me->pv, wk.st);
        peer_step_kernel<true><<<1, 32, 0, wk.stream>>>(me->pv, wk.st);
        init_w_kernel<<<wk.grid_for(std::max<int64_t>(me->slice_len, 1), kThreads), kThreads, 0, wk.stream>>>(me->slice_len, me->ws, me->v + me->slice0, wk.st);
        wk.launches += 3;
    } else if (dist) {
        nrm2_kernel<POST_SSQ><<<wk.grid_for(m, kThreads), kThreads, 0, wk.stream>>>(m, me->u, wk.st, me->g + n);
        wk.launches++;
        LSQRB_TRY(do_atprod_allreduce(me));
        vfinish_kernel<true><<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->g, me->v, wk.st);
        init_w_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->w, me->v, wk.st);
        wk.launches += 2;
    } else {
        nrm2_kernel<POST_INIT_BETA><<<wk.grid_for(m, kThreads), kThreads, 0, wk.stream>>>(m, me->u, wk.st, nullptr);
        wk.launches++;
        LSQRB_TRY(do_atprod<FIN_INIT_ATPROD>(me));
        init_w_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->w, me->v, wk.st);
        wk.launches++;
    }
    record0_kernel<<<1, 1, 0, wk.stream>>>(wk.st, wk.ring_d);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    LSQRB_CUDA(cudaEventRecord(me->ev_t1, wk.stream));

    // ---- iteration loop: enqueue batch j+1, then wait for batch j and look at its records -----
    const bool use_graph = me->opt.use_graph && !me->opt.profile && (!dist || env_int("LSQR_B200_DIST_GRAPH", 1) != 0);
    if (use_graph) { LSQRB_TRY(build_graph(me, wantse)); me->times.iteration_launches = me->graph_launches / std::max(me->batch, 1); }
    const int itnlim = std::max(me->opt.itnlim, 1);   // the reference always runs one iteration (:673-676,798)
    const int B = me->batch;
    int seen = -1, enq = 0, nb = 0, done_batches = 0;   // records consumed, iterations / batches enqueued, batches finished
    bool stop = false;
    auto enqueue_batch = [&]() -> int {
        if (use_graph) {
            LSQRB_CUDA(cudaGraphLaunch(me->graph_exec, wk.stream));
            wk.launches += me->graph_launches;
        } else {
            const int64_t before = wk.launches;
            for (int i = 0; i < B; ++i) LSQRB_TRY(enqueue_iteration(me, wantse));
            LSQRB_TRY(join_side(me));
            me->times.iteration_launches = (wk.launches - before) / B;
        }
        // peer path: a rank that never arrives makes the waits time out; the latch travels with the batch
        if (peer) LSQRB_CUDA(cudaMemcpyAsync(&wk.err_h[nb & 3], &wk.st->comm_error, sizeof(int), cudaMemcpyDeviceToHost, wk.stream));
        LSQRB_CUDA(cudaEventRecord(wk.ev[nb & 3], wk.stream));
        enq += B;
        nb += 1;
        return LSQR_B200_OK;
    };
    auto check = [&](int upto) {
        stop = drain_ring(wk, lc, seen, upto);
        if (seen >= 0 && wk.ring_h[0].arnorm == 0.0) stop = true;   // alpha*beta = 0: no iterations (:646-648)
        if (peer && done_batches > 0 && wk.err_h[(done_batches - 1) & 3]) stop = true;
    };
    // Keep two batches in flight.  The device stops by itself (done flag; istop = 5 at itnlim), so an
    // over-enqueued batch is a run of no-op kernels.  Launch decisions depend only on the records of
    // fully finished batches, which makes them identical on every rank of a multi-GPU run (all ranks
    // must enqueue the same sequence of collectives).
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
    LSQRB_CUDA(cudaEventRecord(me->ev_t2, wk.stream));

    // se(i) = rnorm/sqrt(t) sqrt(se(i))  (:857-865)
    const double *x_src = me->x, *se_src = me->se;
    if (wantse && n > 0) {
        const int64_t mg = dist && me->opt.m_global > 0 ? me->opt.m_global : m;
        double t = 1.0;
        if (mg > n) t = (double)(mg - n);
        if (damp > 0.0) t = (double)mg;
        if (peer) se_finish_kernel<<<wk.grid_for(std::max<int64_t>(me->slice_len, 1), kThreads), kThreads, 0, wk.stream>>>(me->slice_len, ses, wk.st, t);
        else      se_finish_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, me->se, wk.st, t);
        wk.launches++;
    }
    if (peer) {
        // x (and se) exist as owned slices: gather them once, in place
        NcclApi *api = nccl_api();
        LSQRB_NCCL(api->AllGather(xs, me->xs, (size_t)me->pv.cols, ncclFloat64, me->comm, wk.stream));
        if (wantse) LSQRB_NCCL(api->AllGather(ses, me->ses, (size_t)me->pv.cols, ncclFloat64, me->comm, wk.stream));
        x_src = me->xs; se_src = me->ses;
    }
    if (wantse && n > 0) LSQRB_CUDA(cudaMemcpyAsync(se, se_src, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    if (n > 0) LSQRB_CUDA(cudaMemcpyAsync(x, x_src, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    LSQRB_TRY(wk.fetch_state());
    // any records the loop did not print yet (e.g. the stopping iteration)
    drain_ring(wk, lc, seen, wk.h.itn);
    if (wk.h.comm_error) { set_last_error("multi-GPU peer exchange timed out (a rank did not arrive)"); return LSQR_B200_ERR_NCCL; }
    if (wk.h.guard_error) { set_last_error("multi-block SpMV launch: the drift guard timed out (grid not co-resident?)"); return LSQR_B200_ERR_CUDA; }

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

static int ez_solve_through_hook(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                                 double *se, int32_t *itn, double *anorm, double *acond,
                                 double *rnorm, double *arnorm, double *xnorm);

extern "C" {

int lsqr_b200_ez_solve(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                       double *se, int32_t *itn, double *anorm, double *acond,
                       double *rnorm, double *arnorm, double *xnorm)
{
    if (!me) return LSQR_B200_ERR_ARG;
    if ((me->m > 0 && !b) || (me->n > 0 && !x)) { set_last_error("NULL b or x"); return LSQR_B200_ERR_ARG; }
    if (me->opt.engine == 1 && me->opt.world_size == 1)
        return ez_solve_through_hook(me, b, damp, x, istop, se, itn, anorm, acond, rnorm, arnorm, xnorm);
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
// Operator-hook path: lsqr_solver%lsqr with a caller-supplied device operator (src/lsqr.f90:16-30,67-82,432-882).
//
// The callback can only ACCUMULATE (y += A x, x += A'y), so the dscal passes that precede it in the reference
// (:681, :693) stay separate kernels; everything else is folded:
//   K_a  u <- (-alpha/beta) u            [+ the x/w update of the previous iteration and v <- v/(alpha beta), one launch]
//   cb1  u += A v                         (v normalised)            u = A v - alpha u_true          unnormalised, ||u|| = beta'
//   K_b  ||u||  -> step_after_aprod
//   K_c  v <- (-beta'^2) v
//   cb2  v += A'u                         (u unnormalised by beta') v = beta' (A'u_true - beta' v)   ||v|| = beta' alpha'
//   K_d  ||v||  -> step_after_atprod (alpha' = ||v|| / beta')
// i.e. 4 launches of ours per iteration instead of 7, u and v each rescaled once instead of twice, and no host
// synchronisation inside the loop: iterations are enqueued in batches and the device stops itself (done flag);
// the host learns of the stop from the record ring, exactly like the ez path.
// =============================================================================================
namespace lsqrb {

// K_a.  m-part: u <- cu u.  n-part (skipped in the very first iteration: nothing to update yet):
//   v <- cn v (normalise)  ;  x += t1 w ;  w <- v + t2 w ;  se += (t3 w)^2 ; ||w||  -> step_after_update
// first != 0: the launch that follows the initial bidiagonalisation: w = v <- cn v, no update.
template <bool WANTSE>
__global__ void __launch_bounds__(kThreads)
hook_scale_update_kernel(int64_t m, double *__restrict__ u, int64_t n, double *__restrict__ v, double *__restrict__ w,
                         double *__restrict__ x, double *__restrict__ se, DevState *st,
                         volatile lsqr_b200_iter_record *ring, int first)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_exc[2 * kThreads];
    if (st->done) return;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    const bool stopping = st->istop != 0;          // the stop is decided: only the update of that iteration is left
    // ---- n-part first: it closes iteration k (record, done flag)
    s_exc[threadIdx.x] = 0.0;
    s_exc[kThreads + threadIdx.x] = 0.0;
    double sq = 0.0;
    // v holds beta * alpha * v_true; beta = 0: the A' half was skipped and v is still the normalised v (:691-699)
    const double cn = st->beta == 0.0 ? 1.0 : st->inv_alpha * st->inv_beta;
    if (first) {
        for (int64_t i = tid; i < n; i += nthr) { const double t = cn * v[i]; v[i] = t; w[i] = t; }
    } else {
        const double t1 = st->t1, t2 = st->t2, t3 = st->t3;
        for (int64_t i = tid; i < n; i += nthr) {
            const double vn = cn * v[i];
            const double wo = w[i];
            v[i] = vn;
            x[i] = t1 * wo + x[i];
            const double wn = t2 * wo + vn;
            w[i] = wn;
            ssq_add(sq, s_exc + threadIdx.x, kThreads, wn);
            if (WANTSE) se[i] += (t3 * wo) * (t3 * wo);
        }
    }
    // ---- m-part: u <- (-alpha/beta) u  (:681 with the :692 normalisation folded in)
    if (!stopping) {
        const double cu = st->ca_vec;
        for (int64_t i = tid; i < m; i += nthr) u[i] = cu * u[i];
    }
    if (first) return;
    Ssq total;
    if (finish_ssq<kThreads>(st, 1, st->partial2, sq, s_exc, s_red, &total)) {
        __threadfence();
        step_after_update(*st, ssq_norm(total), n > 0 ? __ldcg(x) : 0.0, ring);
    }
}

// K_c: v <- (-beta^2) v  (:693 with the :692 normalisation of u folded into the coefficient); nothing when beta = 0
__global__ void __launch_bounds__(kThreads)
hook_scale_v_kernel(int64_t n, double *__restrict__ v, const DevState *st)
{
    if (st->done || st->istop != 0 || st->beta == 0.0) return;
    const double c = -(st->beta * st->beta);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) v[i] = c * v[i];
}

// K_b / K_d: the norm that follows a callback.  AFTER_AT: v holds beta * (A'u_true - beta v): alpha = ||v|| / beta.
template <bool AFTER_AT, bool INIT>
__global__ void __launch_bounds__(kThreads)
hook_norm_kernel(int64_t n, const double *__restrict__ x, DevState *st)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_exc[2 * kThreads];
    if (st->done || (!INIT && st->istop != 0)) return;
    if (AFTER_AT && !INIT && st->beta == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) step_after_atprod(*st, 0.0, false);   // A' half skipped, alpha kept (:691-699)
        return;
    }
    s_exc[threadIdx.x] = 0.0;
    s_exc[kThreads + threadIdx.x] = 0.0;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        ssq_add(sq, s_exc + threadIdx.x, kThreads, x[i]);
    Ssq total;
    if (finish_ssq<kThreads>(st, 2, st->partial, sq, s_exc, s_red, &total)) {
        const double nrm = ssq_norm(total);
        if (INIT) {
            if (!AFTER_AT) step_init_beta(*st, nrm);
            else           step_init_alpha(*st, st->beta > 0.0 ? nrm * st->inv_beta : 0.0);
        } else {
            if (!AFTER_AT) step_after_aprod(*st, nrm);
            else           step_after_atprod(*st, nrm * st->inv_beta, true);
        }
    }
}

static int lsqr_with_operator(Work &wk, lsqr_b200_aprod_fn aprod, void *aprod_user,
                              int32_t m, int32_t n, double damp, int wantse,
                              double *u, double *v, double *w, double *x, double *se,
                              double atol, double btol, double conlim, int32_t itnlim,
                              const lsqr_b200_options *opts,
                              int32_t *istop, int32_t *itn, double *anorm, double *acond,
                              double *rnorm, double *arnorm, double *xnorm, int64_t est_bytes_per_iter = 0)
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
    const int gm = wk.grid_for(m, kThreads), gn = wk.grid_for(n, kThreads), gmn = std::max(gm, gn);
    auto scale_update = [&](int first) -> int {
        if (wantse) hook_scale_update_kernel<true><<<gmn, kThreads, 0, wk.stream>>>(m, u, n, v, w, x, se, st, wk.ring_d, first);
        else        hook_scale_update_kernel<false><<<gmn, kThreads, 0, wk.stream>>>(m, u, n, v, w, x, se, st, wk.ring_d, first);
        wk.launches++;
        LSQRB_CUDA(cudaGetLastError());
        return LSQR_B200_OK;
    };

    // :621-653: beta = ||u||, v = A'u (u unnormalised), alpha = ||v|| / beta, v <- v/(alpha beta), w = v
    if (n > 0) {
        LSQRB_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * (size_t)n, wk.stream));
        LSQRB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, wk.stream));
        if (wantse) LSQRB_CUDA(cudaMemsetAsync(se, 0, sizeof(double) * (size_t)n, wk.stream));
    }
    hook_norm_kernel<false, true><<<gm, kThreads, 0, wk.stream>>>(m, u, st);
    LSQRB_TRY(call_aprod(2));
    hook_norm_kernel<true, true><<<gn, kThreads, 0, wk.stream>>>(n, v, st);
    wk.launches += 2;
    LSQRB_TRY(scale_update(1));                          // w = v <- v/(alpha beta); u <- (-alpha/beta) u for iteration 1
    record0_kernel<<<1, 1, 0, wk.stream>>>(st, wk.ring_d);
    wk.launches++;
    LSQRB_CUDA(cudaGetLastError());
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));

    int seen = -1;
    bool stop = drain_ring(wk, lc, seen, 0);
    if (wk.ring_h[0].istop != 0.0 || wk.ring_h[0].arnorm == 0.0) stop = true;   // alpha*beta = 0 (:646)

    // iterations per enqueue: ~300 us of device time (the callbacks are opaque, so no CUDA graph here)
    int B = 1;
    if (est_bytes_per_iter > 0) {
        const double est_us = (double)est_bytes_per_iter / 2.5e6 + 30.0;
        B = std::max(1, std::min((int)std::ceil(300.0 / est_us), 8));
    }
    B = std::max(1, std::min(env_int("LSQR_B200_HOOK_BATCH", B), kRingSize / 4));
    const int lim = std::max(itnlim, 1);
    int enq = 0, nb = 0, done_batches = 0;
    auto enqueue_batch = [&]() -> int {
        const int64_t before = wk.launches;
        for (int i = 0; i < B; ++i) {
            LSQRB_TRY(call_aprod(1));                                                  // u += A v           (:682)
            hook_norm_kernel<false, false><<<gm, kThreads, 0, wk.stream>>>(m, u, st);  // beta, anorm       (:683-689)
            hook_scale_v_kernel<<<gn, kThreads, 0, wk.stream>>>(n, v, st);             // v *= -beta        (:693)
            LSQRB_TRY(call_aprod(2));                                                  // v += A'u           (:694)
            hook_norm_kernel<true, false><<<gn, kThreads, 0, wk.stream>>>(n, v, st);   // alpha, rotations, tests (:695-810)
            wk.launches += 3;
            LSQRB_TRY(scale_update(0));                                                // x, w, se (:729-745); v, u rescaled
        }
        wk.iter_launches = (wk.launches - before) / B;     // ours + the operator's, if it counts its launches here
        LSQRB_CUDA(cudaGetLastError());
        LSQRB_CUDA(cudaEventRecord(wk.ev[nb & 3], wk.stream));
        enq += B;
        nb += 1;
        return LSQR_B200_OK;
    };
    while (!stop) {
        while (nb < done_batches + 2 && enq < lim) LSQRB_TRY(enqueue_batch());
        if (done_batches == nb) {
            LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
            stop = drain_ring(wk, lc, seen, lim);
            break;
        }
        LSQRB_CUDA(cudaEventSynchronize(wk.ev[done_batches & 3]));
        done_batches += 1;
        stop = drain_ring(wk, lc, seen, std::min(done_batches * B, lim));
    }
    if (wantse && n > 0) {
        double t = 1.0;
        if (m > n) t = (double)(m - n);
        if (damp > 0.0) t = (double)m;
        se_finish_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, se, st, t);
        wk.launches++;
    }
    LSQRB_TRY(wk.fetch_state());
    drain_ring(wk, lc, seen, wk.h.itn);
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
    // y != NULL: x'y ; y == NULL: ||x|| (scaled, src/lsqrblas.f90:123-159)
    double *d_res = &wk.st->partial[0][kMaxPartials - 1];   // last slot is never used as a block partial (grid < kMaxPartials)
    if (y) dot_kernel<<<std::min(wk.grid_for(n, kThreads), kMaxPartials - 1), kThreads, 0, wk.stream>>>(n, x, y, wk.st, d_res);
    else   nrm2_kernel<POST_NONE><<<std::min(wk.grid_for(n, kThreads), kMaxPartials - 1), kThreads, 0, wk.stream>>>(n, x, wk.st, d_res);
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
    ScopedWork(const lsqr_b200_options *o)
    {
        rc = wk.init(o ? o->device : -1, o ? o->stream : nullptr);
        ok = rc == LSQR_B200_OK;
    }
    // BLAS-1 entry points: `stream` is a plain cudaStream_t argument, NULL = the legacy default stream
    explicit ScopedWork(void *stream_arg)
    {
        rc = wk.init(-1, stream_arg, true);
        ok = rc == LSQR_B200_OK;
    }
    ~ScopedWork() { wk.destroy(); }
};

}  // namespace lsqrb

static int ez_solve_through_hook(lsqr_b200_ez *me, const double *b, double damp, double *x, int32_t *istop,
                                 double *se, int32_t *itn, double *anorm, double *acond,
                                 double *rnorm, double *arnorm, double *xnorm)
{
    // engine = 1: the ez matrix driven through the operator-hook loop (separate dscal / aprod / dnrm2 steps), the way
    // class(lsqr_solver_ez) is a lsqr_solver in the reference; an A/B check of the fused engine
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
                                 istop, itn, anorm, acond, rnorm, arnorm, xnorm, 24 * me->nnz));
    if (n > 0) LSQRB_CUDA(cudaMemcpyAsync(x, me->x, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    if (wantse && n > 0) LSQRB_CUDA(cudaMemcpyAsync(se, me->se, sizeof(double) * (size_t)n, cudaMemcpyDefault, wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(wk.stream));
    me->times.total_launches = wk.launches;
    me->times.iteration_launches = wk.iter_launches;
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
    const cudaError_t sync_err = cudaStreamSynchronize(wk.stream);
    cudaFree(xtmp);
    if (rc != 0) return rc == LSQR_B200_ERR_CUDA ? rc : LSQR_B200_ERR_CALLBACK;
    LSQRB_CUDA(sync_err);
    // w = A'u - damp^2 x (:1089-1094)
    xcheck_w_kernel<<<wk.grid_for(n, kThreads), kThreads, 0, wk.stream>>>(n, w, v, x, damp != 0.0 ? dampsq : 0.0);
    double bnorm, xnorm, rho1, sigma1, rho2, sigma2;
    LSQRB_TRY(reduce_to_host(wk, m, b, nullptr, &bnorm));
    LSQRB_TRY(reduce_to_host(wk, n, x, nullptr, &xnorm));
    LSQRB_TRY(reduce_to_host(wk, m, u, nullptr, &rho1));
    LSQRB_TRY(reduce_to_host(wk, n, v, nullptr, &sigma1));
    if (damp == 0.0) {
        rho2 = rho1;
        sigma2 = sigma1;
    } else {
        rho2 = sqrt(rho1 * rho1 + dampsq * (xnorm * xnorm));
        LSQRB_TRY(reduce_to_host(wk, n, w, nullptr, &sigma2));
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

// =============================================================================================
// The abstract class with the REFERENCE's own signatures: host arrays and a host operator
// (aprod_func, src/lsqr.f90:67-82).  The operator is the user's CPU code, so every product makes a
// host round trip; everything else of LSQR / acheck / xcheck runs on the GPU exactly as in the
// device-operator entry points above.  This is what an unmodified  type,extends(lsqr_solver)  of the
// reference (e.g. test/lsqrtest_module.f90:35-44) binds to; a device operator is the fast path.
// =============================================================================================
}  // extern "C"

namespace lsqrb {

// a device image of a vector that may live on the host
struct DevVec {
    double *d = nullptr;
    double *host = nullptr;
    size_t n = 0;
    bool owned = false;
    int in(cudaStream_t s, double *p, int64_t count, bool copy_in)
    {
        n = (size_t)std::max<int64_t>(count, 0);
        if (!p || is_device_ptr(p)) { d = p; return LSQR_B200_OK; }
        host = p;
        owned = true;
        LSQRB_CUDA(cudaMalloc(&d, sizeof(double) * std::max<size_t>(n, 1)));
        if (copy_in && n) LSQRB_CUDA(cudaMemcpyAsync(d, p, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        return LSQR_B200_OK;
    }
    int out(cudaStream_t s)
    {
        if (owned && n) LSQRB_CUDA(cudaMemcpyAsync(host, d, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        return LSQR_B200_OK;
    }
    ~DevVec() { if (owned && d) cudaFree(d); }
};

struct HostOperator {   // adapts a host aprod to the device-operator hook
    lsqr_b200_aprod_host_fn fn;
    void *user;
    std::vector<double> hx, hy;
};

static int host_operator_trampoline(void *user, int32_t mode, int32_t m, int32_t n, double *x_dev, double *y_dev, void *stream)
{
    HostOperator *op = (HostOperator *)user;
    cudaStream_t s = (cudaStream_t)stream;
    op->hx.resize((size_t)std::max(n, 1));
    op->hy.resize((size_t)std::max(m, 1));
    if (cudaMemcpyAsync(op->hx.data(), x_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s) != cudaSuccess) return 2;
    if (cudaMemcpyAsync(op->hy.data(), y_dev, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s) != cudaSuccess) return 2;
    if (cudaStreamSynchronize(s) != cudaSuccess) return 2;
    const int rc = op->fn(op->user, mode, m, n, op->hx.data(), op->hy.data());
    if (rc != 0) return rc;
    // both arrays are intent(inout) in the reference's interface: only the accumulated one can have changed legally
    if (mode == 1) { if (cudaMemcpyAsync(y_dev, op->hy.data(), sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s) != cudaSuccess) return 2; }
    else           { if (cudaMemcpyAsync(x_dev, op->hx.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s) != cudaSuccess) return 2; }
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 2;   // (the staging vectors are reused by the next call)
}

}  // namespace lsqrb

extern "C" {

int lsqr_b200_lsqr_host(lsqr_b200_aprod_host_fn aprod, void *aprod_user,
                        int32_t m, int32_t n, double damp, int32_t wantse,
                        double *u, double *v, double *w, double *x, double *se,
                        double atol, double btol, double conlim, int32_t itnlim,
                        const lsqr_b200_options *opts,
                        int32_t *istop, int32_t *itn, double *anorm, double *acond,
                        double *rnorm, double *arnorm, double *xnorm)
{
    if (!aprod || m < 0 || n < 0) return LSQR_B200_ERR_ARG;
    if ((m > 0 && !u) || (n > 0 && (!v || !w || !x)) || (wantse && n > 0 && !se)) { set_last_error("NULL work vector"); return LSQR_B200_ERR_ARG; }
    ScopedWork sw(opts);
    if (!sw.ok) return sw.rc;
    cudaStream_t s = sw.wk.stream;
    DevVec du, dv, dw, dx, dse;
    LSQRB_TRY(du.in(s, u, m, true));            // u holds b on entry (src/lsqr.f90:461-462)
    LSQRB_TRY(dv.in(s, v, n, false));
    LSQRB_TRY(dw.in(s, w, n, false));
    LSQRB_TRY(dx.in(s, x, n, false));
    LSQRB_TRY(dse.in(s, wantse ? se : nullptr, n, false));
    HostOperator op{aprod, aprod_user, {}, {}};
    LSQRB_TRY(lsqr_with_operator(sw.wk, host_operator_trampoline, &op, m, n, damp, wantse, du.d, dv.d, dw.d, dx.d, dse.d,
                                 atol, btol, conlim, itnlim, opts, istop, itn, anorm, acond, rnorm, arnorm, xnorm));
    LSQRB_TRY(du.out(s)); LSQRB_TRY(dv.out(s)); LSQRB_TRY(dw.out(s)); LSQRB_TRY(dx.out(s)); LSQRB_TRY(dse.out(s));
    LSQRB_CUDA(cudaStreamSynchronize(s));
    return LSQR_B200_OK;
}

int lsqr_b200_acheck_host(lsqr_b200_aprod_host_fn aprod, void *aprod_user, int32_t m, int32_t n, double eps,
                          double *v, double *w, double *x, double *y,
                          const lsqr_b200_options *opts, int32_t *inform, double *relerr)
{
    if (!aprod || m < 1 || n < 1 || !v || !w || !x || !y) return LSQR_B200_ERR_ARG;
    ScopedWork sw(opts);
    if (!sw.ok) return sw.rc;
    cudaStream_t s = sw.wk.stream;
    DevVec dv, dw, dx, dy;
    LSQRB_TRY(dv.in(s, v, n, false)); LSQRB_TRY(dw.in(s, w, m, false)); LSQRB_TRY(dx.in(s, x, n, false)); LSQRB_TRY(dy.in(s, y, m, false));
    HostOperator op{aprod, aprod_user, {}, {}};
    lsqr_b200_options o;
    if (opts) o = *opts; else lsqr_b200_default_options(&o);
    o.stream = (void *)s;
    LSQRB_TRY(lsqr_b200_acheck(host_operator_trampoline, &op, m, n, eps, dv.d, dw.d, dx.d, dy.d, &o, inform, relerr));
    LSQRB_TRY(dv.out(s)); LSQRB_TRY(dw.out(s)); LSQRB_TRY(dx.out(s)); LSQRB_TRY(dy.out(s));
    LSQRB_CUDA(cudaStreamSynchronize(s));
    return LSQR_B200_OK;
}

int lsqr_b200_xcheck_host(lsqr_b200_aprod_host_fn aprod, void *aprod_user, int32_t m, int32_t n,
                          double anorm, double damp, double eps,
                          const double *b, double *u, double *v, double *w, const double *x,
                          const lsqr_b200_options *opts,
                          int32_t *inform, double *test1, double *test2, double *test3, double *norms)
{
    if (!aprod || m < 1 || n < 1 || !b || !u || !v || !w || !x) return LSQR_B200_ERR_ARG;
    ScopedWork sw(opts);
    if (!sw.ok) return sw.rc;
    cudaStream_t s = sw.wk.stream;
    DevVec db, du, dv, dw, dx;
    LSQRB_TRY(db.in(s, const_cast<double *>(b), m, true)); LSQRB_TRY(dx.in(s, const_cast<double *>(x), n, true));
    LSQRB_TRY(du.in(s, u, m, false)); LSQRB_TRY(dv.in(s, v, n, false)); LSQRB_TRY(dw.in(s, w, n, false));
    HostOperator op{aprod, aprod_user, {}, {}};
    lsqr_b200_options o;
    if (opts) o = *opts; else lsqr_b200_default_options(&o);
    o.stream = (void *)s;
    LSQRB_TRY(lsqr_b200_xcheck(host_operator_trampoline, &op, m, n, anorm, damp, eps, db.d, du.d, dv.d, dw.d, dx.d, &o,
                               inform, test1, test2, test3, norms));
    LSQRB_TRY(du.out(s)); LSQRB_TRY(dv.out(s)); LSQRB_TRY(dw.out(s));      // r, A'r, A'r - damp^2 x (b and x are inputs)
    LSQRB_CUDA(cudaStreamSynchronize(s));
    return LSQR_B200_OK;
}

// ---- BLAS-1 (src/lsqrblas.f90), stride 1.  Arrays may be DEVICE or HOST arrays (host arrays are staged through the GPU:
// the arithmetic always happens on the device); the result is a host scalar.
int lsqr_b200_dnrm2(int64_t n, const double *x, double *result, void *stream)
{
    if (!result || n < 0 || (n > 0 && !x)) return LSQR_B200_ERR_ARG;
    if (n < 1) { *result = 0.0; return LSQR_B200_OK; }   // :131
    ScopedWork sw(stream);
    if (!sw.ok) return sw.rc;
    DevVec dx;
    LSQRB_TRY(dx.in(sw.wk.stream, const_cast<double *>(x), n, true));
    if (n == 1) {   // :132 -- abs(x(1))
        double t;
        LSQRB_CUDA(cudaMemcpyAsync(&t, dx.d, sizeof(double), cudaMemcpyDeviceToHost, sw.wk.stream));
        LSQRB_CUDA(cudaStreamSynchronize(sw.wk.stream));
        *result = fabs(t);
        return LSQR_B200_OK;
    }
    return reduce_to_host(sw.wk, n, dx.d, nullptr, result);
}

int lsqr_b200_ddot(int64_t n, const double *x, const double *y, double *result, void *stream)
{
    if (!result || n < 0 || (n > 0 && (!x || !y))) return LSQR_B200_ERR_ARG;
    if (n < 1) { *result = 0.0; return LSQR_B200_OK; }
    ScopedWork sw(stream);
    if (!sw.ok) return sw.rc;
    DevVec dx, dy;
    LSQRB_TRY(dx.in(sw.wk.stream, const_cast<double *>(x), n, true));
    LSQRB_TRY(dy.in(sw.wk.stream, const_cast<double *>(y), n, true));
    return reduce_to_host(sw.wk, n, dx.d, dy.d, result);
}

int lsqr_b200_dscal(int64_t n, double da, double *x, void *stream)
{
    if (n < 0 || (n > 0 && !x)) return LSQR_B200_ERR_ARG;
    if (n == 0) return LSQR_B200_OK;
    ScopedWork sw(stream);
    if (!sw.ok) return sw.rc;
    DevVec dx;
    LSQRB_TRY(dx.in(sw.wk.stream, x, n, true));
    LSQRB_TRY(scal_imm(sw.wk, n, dx.d, da));
    LSQRB_TRY(dx.out(sw.wk.stream));
    LSQRB_CUDA(cudaStreamSynchronize(sw.wk.stream));
    return LSQR_B200_OK;
}

int lsqr_b200_dcopy(int64_t n, const double *x, double *y, void *stream)
{
    if (n < 0 || (n > 0 && (!x || !y))) return LSQR_B200_ERR_ARG;
    if (n == 0) return LSQR_B200_OK;
    int count = lsqr_b200_device_count();
    if (count == 0) return LSQR_B200_ERR_NO_DEVICE;
    LSQRB_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * (size_t)n, cudaMemcpyDefault, (cudaStream_t)stream));
    LSQRB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return LSQR_B200_OK;
}

}  // extern "C"
