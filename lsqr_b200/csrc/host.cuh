// host.cuh -- host-side plumbing shared by the ez path and the operator-hook path: error latch, NCCL loaded at run
// time, the Work context (stream + device state + record ring), the nout log rebuilt from the record ring.
#pragma once

#include <dlfcn.h>
#include <nccl.h>   // types only; the library is dlopen'ed so single-GPU use has no NCCL dependency

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "vecops.cuh"

namespace lsqrb {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;   // per host thread: different handles may be driven from different threads
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
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
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
    int *err_h = nullptr;                         // pinned: comm_error of the last 4 batches (multi-GPU peer path)
    int max_grid = kNumSMs * 8;
    int sms = kNumSMs;
    int64_t launches = 0;
    int64_t iter_launches = 0;   // operator-hook loop: launches per iteration of the last solve
    int pdl = 0;             // LSQR_B200_PDL: programmatic dependent launch of the A v kernel behind the A'u kernel
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    // user_stream == NULL: a library-owned BLOCKING stream (it synchronises implicitly with the legacy default stream,
    // so data the caller produced on stream 0 -- torch's default -- is ordered before our reads, and our results before
    // the caller's later stream-0 work); null_is_legacy: NULL means the legacy default stream itself (CUDA's convention
    // for a `cudaStream_t stream` ARGUMENT, as opposed to the options field).
    int init(int dev, void *user_stream, bool null_is_legacy = false)
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
        } else if (null_is_legacy) {
            stream = cudaStreamLegacy;
        } else {
            LSQRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamDefault));
            own_stream = true;
        }
        LSQRB_CUDA(cudaMalloc(&st, sizeof(DevState)));
        LSQRB_CUDA(cudaMemsetAsync(st, 0, sizeof(DevState), stream));
        LSQRB_CUDA(cudaHostAlloc(&ring_h, sizeof(lsqr_b200_iter_record) * kRingSize, cudaHostAllocMapped));
        LSQRB_CUDA(cudaHostGetDevicePointer(&ring_d, ring_h, 0));
        LSQRB_CUDA(cudaHostAlloc(&err_h, sizeof(int) * 4, cudaHostAllocDefault));
        memset(err_h, 0, sizeof(int) * 4);
        h.epoch = 0;
        for (auto &e : ev) LSQRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDefault));
        LSQRB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        sms = std::max(1, sms);
        max_grid = std::min(kMaxPartials, std::max(1, sms) * env_int("LSQR_B200_BLOCKS_PER_SM", 8));
        return LSQR_B200_OK;
    }

    void destroy()
    {
        if (st) cudaFree(st);
        if (ring_h) cudaFreeHost(ring_h);
        if (err_h) cudaFreeHost(err_h);
        err_h = nullptr;
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
        const unsigned int epoch = h.epoch;   // the exchange epoch of the multi-GPU peer path never goes back
        memset(&h, 0, offsetof(DevState, partial));
        h.epoch = epoch;
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

static inline int vec_ok(const void *a, const void *b, const void *c, const void *d)
{
    auto al = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
    return al(a) && al(b) && al(c) && al(d);
}

template <bool LAZY>
static int launch_update(Work &wk, int64_t n, double *x, double *w, const double *v, double *se, bool wantse, int local_only = 0)
{
    const int grid = wk.grid_for((n + 1) / 2, kThreads);
    const int vok = vec_ok(x, w, v, wantse ? se : nullptr);
    if (wantse) xw_update_kernel<true, LAZY><<<grid, kThreads, 0, wk.stream>>>(n, x, w, v, se, wk.st, wk.ring_d, vok, local_only);
    else        xw_update_kernel<false, LAZY><<<grid, kThreads, 0, wk.stream>>>(n, x, w, v, se, wk.st, wk.ring_d, vok, local_only);
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
