// peer.cuh -- K8, multi-GPU exchange over NVLink peer memory (one process per GPU, A row-partitioned).
//
// The n columns are split into one contiguous slice per rank.  Per iteration:
//   1. Atprod (spmv.cuh, FIN_PUSH): every rank's partial (A_p'u_p)[j] is stored straight into the receive buffer of
//      the rank that owns column j (a coalesced peer store from the SpMV epilogue, so the transfer overlaps the
//      product); the rank's partial sums of squares ride along, then one flag per peer.
//   2. peer_vfinish_kernel (owner of a slice): waits for the P flags, sums the P contributions of each of its
//      columns in rank order (deterministic), forms v' = g/beta - (beta/alpha) v on the slice only and stores it into
//      the v of EVERY rank (the all-gather, again by peer stores); partial sum v'^2 and a second flag follow.
//   3. peer_step_kernel (every rank, one thread): waits for the second flags, sums the partials in rank order --
//      beta, alpha, ||w|| are therefore bit-identical on every rank -- and advances the scalar recurrence.
//   4. the x/w update runs on the owned slice only (x, w are never replicated; x is gathered once at the end).
// Compared with one all-reduce of n+1 doubles and replicated n-vector passes this moves the same bytes over NVLink
// but hides the reduce half behind the SpMV, and divides the n-vector work by P.
// Flags carry an iteration epoch; the two-flag handshake makes single buffers safe: a rank can only start pushing
// iteration k+1 after every owner has consumed iteration k (it needs all of v' first).
#pragma once

#include "steps.cuh"

namespace lsqrb {

constexpr int kMaxRanks = 16;
constexpr int kScDoubles = 16;   // per-rank scalar record: usq[3] wsq[3] vsq[3]

// Device view of the symmetric exchange block of every rank (same layout everywhere).
struct PeerView {
    int world, rank;
    int64_t cols;                        // columns per slice (the last slice may be shorter)
    int64_t n;
    double *recv[kMaxRanks];             // recv[q] : rank q's receive area, [world][cols]
    double *v[kMaxRanks];                // v[q]    : rank q's full v (n entries)
    double *sc[kMaxRanks];               // sc[q]   : rank q's scalar area, [world][kScDoubles]
    unsigned int *flag1[kMaxRanks];      // flag1[q]: rank q's "contributions of rank p have landed" flags, [world]
    unsigned int *flag2[kMaxRanks];      // flag2[q]: "slice of rank p has landed" flags, [world]
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One thread waits until every flag of flags[0..world) has reached `want`.  A peer that never arrives would hang the
// GPU: give up after ~20 s, latch comm_error and let the solve run to its (meaningless) end; the host reports it.
__device__ __forceinline__ void peer_wait(DevState *st, const unsigned int *flags, int world, unsigned int want)
{
    const unsigned long long t0 = globaltimer_ns();
    for (int p = 0; p < world; ++p) {
        while ((int)(ld_acquire_sys(flags + p) - want) < 0) {
            if (*(volatile int *)&st->comm_error) return;
            if (globaltimer_ns() - t0 > 20000000000ull) { st->comm_error = 1; return; }
            __nanosleep(100);
        }
    }
}

// Tail of the FIN_PUSH product: executed by thread 0 of the block that finishes last.
__device__ __forceinline__ void peer_publish_partials(DevState *st, const PeerView &pv)
{
    __threadfence_system();                        // every block fenced its peer stores before taking its ticket
    const unsigned int want = st->epoch + 1u;
    for (int q = 0; q < pv.world; ++q) {
        double *d = pv.sc[q] + (size_t)pv.rank * kScDoubles;
        d[0] = st->usq_local.med; d[1] = st->usq_local.big; d[2] = st->usq_local.sml;
        d[3] = st->wsq_local.med; d[4] = st->wsq_local.big; d[5] = st->wsq_local.sml;
    }
    __threadfence_system();
    for (int q = 0; q < pv.world; ++q) st_release_sys(pv.flag1[q] + pv.rank, want);
}

// Owner side of the exchange.  Grid-stride over the owned slice.
template <bool INIT>
__global__ void __launch_bounds__(kThreads)
peer_vfinish_kernel(PeerView pv, DevState *st)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_exc[2 * kThreads];
    __shared__ double s_beta;
    if (st->done) return;
    const int tid = threadIdx.x;
    const unsigned int want = st->epoch + 1u;
    if (tid == 0) {
        peer_wait(st, pv.flag1[pv.rank], pv.world, want);
        Ssq usq = Ssq{0.0, 0.0, 0.0};
        for (int p = 0; p < pv.world; ++p) {
            const double *d = pv.sc[pv.rank] + (size_t)p * kScDoubles;
            usq = ssq_sum(usq, Ssq{__ldcg(d + 0), __ldcg(d + 1), __ldcg(d + 2)});
        }
        s_beta = ssq_norm(usq);
    }
    __syncthreads();
    const double beta = s_beta;
    s_exc[tid] = 0.0;
    s_exc[kThreads + tid] = 0.0;
    double sq = 0.0;
    const int64_t c0 = (int64_t)pv.rank * pv.cols;
    const int64_t len = max((int64_t)0, min(pv.cols, pv.n - c0));
    if (beta != 0.0) {
        const double cm = 1.0 / beta;
        const double cv = INIT ? 0.0 : -beta * st->inv_alpha;
        const double *rbase = pv.recv[pv.rank];
        double *vmine = pv.v[pv.rank] + c0;
        for (int64_t j = (int64_t)blockIdx.x * kThreads + tid; j < len; j += (int64_t)gridDim.x * kThreads) {
            double g = 0.0;
            for (int p = 0; p < pv.world; ++p) g += __ldcg(rbase + (size_t)p * pv.cols + j);   // rank order: deterministic
            const double r = INIT ? cm * g : cm * g + cv * vmine[j];
            for (int q = 0; q < pv.world; ++q) pv.v[q][c0 + j] = r;                               // all-gather by peer stores
            ssq_add(sq, s_exc + tid, kThreads, r);
        }
    }
    __syncthreads();
    if (tid == 0) __threadfence_system();          // this block's peer stores, before its ticket
    Ssq total;
    if (finish_ssq<kThreads>(st, 3, st->partial, sq, s_exc, s_red, &total)) {
        st->vsq_local = total;
        __threadfence_system();
        for (int q = 0; q < pv.world; ++q) {
            double *d = pv.sc[q] + (size_t)pv.rank * kScDoubles;
            d[6] = total.med; d[7] = total.big; d[8] = total.sml;
        }
        __threadfence_system();
        for (int q = 0; q < pv.world; ++q) st_release_sys(pv.flag2[q] + pv.rank, want);
    }
}

// Every rank: all slices of v' have landed; combine the partial sums in rank order and take the scalar step.
template <bool INIT>
__global__ void peer_step_kernel(PeerView pv, DevState *st)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st->done) return;
    const unsigned int want = st->epoch + 1u;
    peer_wait(st, pv.flag2[pv.rank], pv.world, want);
    Ssq usq = Ssq{0.0, 0.0, 0.0}, wsq = usq, vsq = usq;
    for (int p = 0; p < pv.world; ++p) {
        const double *d = pv.sc[pv.rank] + (size_t)p * kScDoubles;
        usq = ssq_sum(usq, Ssq{__ldcg(d + 0), __ldcg(d + 1), __ldcg(d + 2)});
        wsq = ssq_sum(wsq, Ssq{__ldcg(d + 3), __ldcg(d + 4), __ldcg(d + 5)});
        vsq = ssq_sum(vsq, Ssq{__ldcg(d + 6), __ldcg(d + 7), __ldcg(d + 8)});
    }
    const double beta = ssq_norm(usq);
    if (INIT) {
        step_init_beta(*st, beta);
        step_init_alpha(*st, beta > 0.0 ? ssq_norm(vsq) : 0.0);
    } else {
        if (st->itn > 0) st->wnorm = ssq_norm(wsq);   // ||w|| over all slices (the update left per-rank partials)
        step_after_aprod(*st, beta);
        step_after_atprod(*st, ssq_norm(vsq), beta > 0.0);
    }
    st->epoch = want;
}

}  // namespace lsqrb
