// vecops.cuh -- K5 and the BLAS-1 kernels of the hot path (hand-written, sm_100a).
//
//   K5 xw_update     x += t1 w; w = v/alpha + t2 w; sum(w^2) (+ se)   replaces the loop at src/lsqr.f90:729-745
//   BLAS-1           dnrm2 / ddot / dscal equivalents (src/lsqrblas.f90) for the reference-structure path, acheck,
//                    xcheck and the public lsqr_b200_dnrm2/ddot/dscal/dcopy
//   K8 vfinish       multi-GPU (NCCL path): v' from the all-reduced [A'u' | sum u'^2]
// All of it is HBM-bound streaming work: coalesced 16-byte accesses, grids sized from the SM count, deterministic
// (fixed-slot, fixed-tree) reductions with Blue's scaled accumulators.
#pragma once

#include "steps.cuh"

namespace lsqrb {

// =============================================================================================
// K5: x/w(/se) update (src/lsqr.f90:729-745) + ||w'||; last block closes the iteration.
//   x += t1*w ;  w' = inv_alpha*v + t2*w ;  se += (t3*w)^2
// Multi-GPU peer path: the kernel works on this rank's slice [0, n) of x, w (v is offset by the caller) and leaves
// its partial sum in st->wsq_local instead of closing the iteration (local_only).
// =============================================================================================
template <bool WANTSE, bool LAZY>
__global__ void __launch_bounds__(kThreads)
xw_update_kernel(int64_t n, double *__restrict__ x, double *__restrict__ w, const double *__restrict__ v,
                 double *__restrict__ se, DevState *st, volatile lsqr_b200_iter_record *ring, int vec_ok, int local_only)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_exc[2 * kThreads];
    if (st->done) return;
    const double t1 = st->t1, t2 = st->t2, t3 = st->t3;
    const double ia = LAZY ? st->inv_alpha : 1.0;   // LAZY: v is stored unnormalised
    s_exc[threadIdx.x] = 0.0;
    s_exc[kThreads + threadIdx.x] = 0.0;
    double *exc = s_exc + threadIdx.x;
    double sq = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    if (vec_ok) {
        const int64_t n2 = n >> 1;
        double2 *x2 = reinterpret_cast<double2 *>(x);
        double2 *w2 = reinterpret_cast<double2 *>(w);
        const double2 *v2 = reinterpret_cast<const double2 *>(v);
        double2 *se2 = reinterpret_cast<double2 *>(se);
        for (int64_t i = tid; i < n2; i += nthr) {
            const double2 wo = w2[i], vv = v2[i];
            double2 xo = x2[i];
            xo.x = t1 * wo.x + xo.x;
            xo.y = t1 * wo.y + xo.y;
            double2 wn;
            wn.x = t2 * wo.x + ia * vv.x;
            wn.y = t2 * wo.y + ia * vv.y;
            x2[i] = xo;
            w2[i] = wn;
            ssq_add(sq, exc, kThreads, wn.x);
            ssq_add(sq, exc, kThreads, wn.y);
            if (WANTSE) {
                double2 s = se2[i];
                s.x += (t3 * wo.x) * (t3 * wo.x);
                s.y += (t3 * wo.y) * (t3 * wo.y);
                se2[i] = s;
            }
        }
    }
    // scalar tail (odd n) or the whole range when the arrays are not 16-byte aligned
    for (int64_t i = (vec_ok ? (n & ~(int64_t)1) : 0) + tid; i < n; i += nthr) {
        const double wo = w[i];
        x[i] = t1 * wo + x[i];
        const double wn = t2 * wo + ia * v[i];
        w[i] = wn;
        ssq_add(sq, exc, kThreads, wn);
        if (WANTSE) se[i] += (t3 * wo) * (t3 * wo);
    }
    Ssq total;
    // own partial slots and ticket: this kernel may run next to the Aprod of the following iteration
    if (finish_ssq<kThreads>(st, 1, st->partial2, sq, s_exc, s_red, &total)) {
        // local_only (multi-GPU peer path): the sum covers this rank's slice; it travels with the next exchange and
        // ||w|| is formed by peer_step_kernel.  The record of the iteration is published either way.
        if (local_only) st->wsq_local = total;
        __threadfence();
        const double x1 = n > 0 ? __ldcg(x) : 0.0;   // x(1) after the update (rank 0 owns it)
        step_after_update(*st, local_only ? st->wnorm : ssq_norm(total), x1, ring);
    }
}

// w = v/alpha (src/lsqr.f90:641-644), lazy-normalised form
__global__ void __launch_bounds__(kThreads)
init_w_kernel(int64_t n, double *__restrict__ w, const double *__restrict__ v, const DevState *st)
{
    if (st->done) return;
    const double ia = st->inv_alpha;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        w[i] = ia * v[i];
}

// se(i) = rnorm/sqrt(t) * sqrt(se(i))  (src/lsqr.f90:857-865); only if at least one iteration ran
__global__ void __launch_bounds__(kThreads)
se_finish_kernel(int64_t n, double *__restrict__ se, const DevState *st, double tdiv)
{
    if (st->itn == 0) return;
    const double t = st->rnorm / sqrt(tdiv);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        se[i] = t * sqrt(se[i]);
}

// =============================================================================================
// BLAS-1 kernels (src/lsqrblas.f90).  The reducing kernels hand their result to the last block.
// =============================================================================================
enum NormPost {
    POST_NONE = 0,        // result only
    POST_INIT_BETA = 1,   // step_init_beta
    POST_INIT_ALPHA = 2,  // step_init_alpha
    POST_G_BETA = 3,      // step_after_aprod   (reference-structure path: u already holds A v - alpha u)
    POST_G_ALPHA = 4,     // step_after_atprod  (reference-structure path: v already holds A'u - beta v)
    POST_SSQ = 5          // the three accumulators to result[0..2] (multi-GPU: partial sum to be all-reduced)
};

// dnrm2 (src/lsqrblas.f90:123-159): scaled sum of squares; result = the norm (POST_NONE) or the accumulators
template <int POST>
__global__ void __launch_bounds__(kThreads)
nrm2_kernel(int64_t n, const double *__restrict__ x, DevState *st, double *result)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_exc[2 * kThreads];
    if ((POST == POST_G_BETA || POST == POST_G_ALPHA) && st->done) return;
    s_exc[threadIdx.x] = 0.0;
    s_exc[kThreads + threadIdx.x] = 0.0;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        ssq_add(sq, s_exc + threadIdx.x, kThreads, x[i]);
    Ssq total;
    if (finish_ssq<kThreads>(st, 2, st->partial, sq, s_exc, s_red, &total)) {
        if (POST == POST_SSQ) { result[0] = total.med; result[1] = total.big; result[2] = total.sml; return; }
        const double nrm = ssq_norm(total);
        if (POST == POST_INIT_BETA) step_init_beta(*st, nrm);
        if (POST == POST_INIT_ALPHA) step_init_alpha(*st, nrm);
        if (POST == POST_G_BETA) step_after_aprod(*st, nrm);
        if (POST == POST_G_ALPHA) step_after_atprod(*st, nrm, st->beta > 0.0);
        if (result) *result = nrm;
    }
}

// ddot (src/lsqrblas.f90:74-116).  Own partial slots: plain sums, no scaling (the reference has none either).
__global__ void __launch_bounds__(kThreads)
dot_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y, DevState *st, double *result)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ int s_last;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        acc += x[i] * y[i];
    const double bs = block_sum<kThreads>(acc, s_red);
    if (threadIdx.x == 0) {
        __stcg(&st->partial[0][blockIdx.x], bs);
        __threadfence();
        s_last = atomicAdd(&st->counter[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) a += __ldcg(&st->partial[0][i]);
    a = block_sum<kThreads>(a, s_red);
    if (threadIdx.x == 0) { st->counter[2] = 0; *result = a; }
}

// x *= *coef (coef on device) or x *= imm when coef == nullptr
__global__ void __launch_bounds__(kThreads)
scal_kernel(int64_t n, double *__restrict__ x, const double *coef, double imm)
{
    const double a = coef ? *coef : imm;
    if (a == 1.0) return;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        x[i] = a * x[i];
}

// =============================================================================================
// K8 (multi-GPU, NCCL path): after the all-reduce of g = [ A'u' (n entries) | Ssq(u') ] every rank forms
//   beta = ||u'||;  v' = g/beta - (beta/alpha) v;  ||v'||
// redundantly (v is replicated), so alpha, the rotations and the stop decision are bit-identical
// on all ranks and need no further collective.
// =============================================================================================
template <bool INIT>
__global__ void __launch_bounds__(kThreads)
vfinish_kernel(int64_t n, const double *__restrict__ g, double *__restrict__ v, DevState *st)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ double s_exc[2 * kThreads];
    if (st->done) return;
    const Ssq usq = Ssq{g[n], g[n + 1], g[n + 2]};
    const double beta = ssq_norm(usq);
    if (beta == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (INIT) { step_init_beta(*st, 0.0); step_init_alpha(*st, 0.0); }
            else      { step_after_aprod(*st, 0.0); step_after_atprod(*st, 0.0, false); }
        }
        return;
    }
    s_exc[threadIdx.x] = 0.0;
    s_exc[kThreads + threadIdx.x] = 0.0;
    const double cm = 1.0 / beta;
    const double cv = INIT ? 0.0 : -beta * st->inv_alpha;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double r = INIT ? cm * g[i] : cm * g[i] + cv * v[i];
        v[i] = r;
        ssq_add(sq, s_exc + threadIdx.x, kThreads, r);
    }
    Ssq total;
    if (finish_ssq<kThreads>(st, 3, st->partial, sq, s_exc, s_red, &total)) {
        if (INIT) { step_init_beta(*st, beta); step_init_alpha(*st, ssq_norm(total)); }
        else      { step_after_aprod(*st, beta); step_after_atprod(*st, ssq_norm(total), true); }
    }
}

}  // namespace lsqrb
