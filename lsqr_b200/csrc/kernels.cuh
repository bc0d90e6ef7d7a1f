// kernels.cuh -- hand-written sm_100a kernels of the LSQR hot path.
//
//   K3 aprod_fused   u' = ca_mat*(A v) + ca_vec*u,  sum(u'^2)   replaces dscal+aprod(1)+dnrm2 (src/lsqr.f90:681-683)
//   K4 atprod_fused  v' = ct_mat*(A'u') + ct_vec*v, sum(v'^2)   replaces dscal+aprod(2)+dnrm2 (src/lsqr.f90:692-697)
//   K5 xw_update     x += t1 w; w = v/alpha + t2 w; sum(w^2)    replaces the loop at src/lsqr.f90:729-745
//   K6 scalar steps  run by the LAST block of K3/K4/K5 (src/lsqr.f90:687-689,703-721,751-810,843-850)
//   BLAS-1           dnrm2 / ddot / dscal / dcopy equivalents (src/lsqrblas.f90) for the operator-hook path
//
// FP64 SpMV is HBM-bound (2 flop per 12 B): no tensor cores.  The rules that matter are
// coalesced wide loads of val/idx, keeping the gathered dense vector in L2, persistent grids
// sized in multiples of 148 SMs, and deterministic (fixed-slot, fixed-tree) reductions.
#pragma once

#include "common.cuh"

namespace lsqrb {

constexpr int kThreads = 256;

// =============================================================================================
// Scalar recurrence (K6).  Each function is executed by exactly one thread.
// =============================================================================================

// after ||b||:  src/lsqr.f90:632-636
__device__ __forceinline__ void step_init_beta(DevState &s, double sumsq)
{
    s.beta = sqrt(sumsq);
    s.inv_beta = s.beta > 0.0 ? 1.0 / s.beta : 1.0;
    s.alpha = 0.0;
    s.inv_alpha = 1.0;
    s.ct_mat = s.inv_beta;   // v = A'(u/beta)
    s.ct_vec = 0.0;
    s.g_c1 = s.inv_beta;
}

// after ||A'u||:  src/lsqr.f90:637-653
__device__ __forceinline__ void step_init_alpha(DevState &s, double sumsq)
{
    s.alpha = (s.beta > 0.0) ? sqrt(sumsq) : 0.0;
    s.inv_alpha = s.alpha > 0.0 ? 1.0 / s.alpha : 1.0;
    s.arnorm = s.alpha * s.beta;
    if (s.arnorm != 0.0) {
        s.rhobar = s.alpha;
        s.phibar = s.beta;
        s.bnorm = s.beta;
        s.rnorm = s.beta;
        s.ca_mat = s.inv_alpha;
        s.ca_vec = -s.alpha * s.inv_beta;
        s.wnorm2 = sumsq * s.inv_alpha * s.inv_alpha;   // w = v/alpha
        s.g_c0 = -s.alpha;
        s.g_c3 = s.inv_alpha;
    } else {
        // x = 0 is the exact solution (istop = 0, no iterations).  The reference leaves rnorm
        // unassigned on this path (src/lsqr.f90:646-653); we report ||b||, the true residual of x = 0.
        s.rnorm = s.beta;
        s.bnorm = s.beta;
        s.istop = 0;
        s.done = 1;
    }
}

// after ||u'||:  src/lsqr.f90:676,683-693
__device__ __forceinline__ void step_after_aprod(DevState &s, double sumsq)
{
    s.itn += 1;
    const double beta = sqrt(sumsq);
    s.beta = beta;
    double temp = d2norm(s.alpha, beta);
    temp = d2norm(temp, s.damp);
    s.anorm = d2norm(s.anorm, temp);
    if (beta > 0.0) {
        s.inv_beta = 1.0 / beta;
        s.ct_mat = s.inv_beta;
        s.ct_vec = -beta * s.inv_alpha;
        s.g_c1 = s.inv_beta;
        s.g_c2 = -beta;
    } else {
        s.inv_beta = 1.0;   // u is not rescaled and the A' half is skipped (src/lsqr.f90:691)
        s.g_c1 = 1.0;
        s.g_c2 = 1.0;
    }
}

// after ||v'||: rotations, estimates and stopping tests, src/lsqr.f90:695-721,724-726,751-810,843-850
__device__ __forceinline__ void step_after_atprod(DevState &s, double sumsq, bool new_alpha)
{
    if (new_alpha) {
        s.alpha = sqrt(sumsq);
        s.inv_alpha = s.alpha > 0.0 ? 1.0 / s.alpha : 1.0;   // alpha = 0: v is left unscaled (:696-698)
    }
    const double alpha = s.alpha, beta = s.beta;
    s.g_c3 = (new_alpha && alpha > 0.0) ? s.inv_alpha : 1.0;
    s.g_c0 = -alpha;

    // plane rotation that removes damp (:703-710)
    double rhbar1 = s.rhobar;
    if (s.damped) {
        rhbar1 = d2norm(s.rhobar, s.damp);
        const double cs1 = s.rhobar / rhbar1;
        const double sn1 = s.damp / rhbar1;
        s.psi = sn1 * s.phibar;
        s.phibar = cs1 * s.phibar;
    }

    // plane rotation that removes the subdiagonal beta (:714-721)
    const double rho = d2norm(rhbar1, beta);
    const double cs = rhbar1 / rho;
    const double sn = beta / rho;
    const double theta = sn * alpha;
    s.rhobar = -cs * alpha;
    const double phi = cs * s.phibar;
    s.phibar = sn * s.phibar;
    const double tau = sn * phi;

    // coefficients of the x/w update (:724-726)
    s.t1 = phi / rho;
    s.t2 = -theta / rho;
    s.t3 = 1.0 / rho;

    // dknorm = sqrt(sum (t3 w_i)^2) = |t3| ||w||  (:729-751); ||w||^2 was produced when w was written
    const double dknorm = fabs(s.t3) * sqrt(s.wnorm2);
    s.dnorm = d2norm(s.dnorm, dknorm);
    const double dxk = fabs(phi * dknorm);
    if (s.dxmax < dxk) {
        s.dxmax = dxk;
        s.maxdx = s.itn;
    }

    // right rotation, estimate of norm(x) (:762-771)
    const double delta = s.sn2 * rho;
    const double gambar = -s.cs2 * rho;
    const double rhs = phi - delta * s.z;
    const double zbar = rhs / gambar;
    s.xnorm = d2norm(s.xnorm1, zbar);
    const double gamma = d2norm(gambar, theta);
    s.cs2 = gambar / gamma;
    s.sn2 = theta / gamma;
    s.z = rhs / gamma;
    s.xnorm1 = d2norm(s.xnorm1, s.z);

    // estimates (:776-790)
    s.acond = s.anorm * s.dnorm;
    s.res2 = d2norm(s.res2, s.psi);
    s.rnorm = d2norm(s.res2, s.phibar);
    s.arnorm = alpha * fabs(tau);

    s.alfopt = sqrt(s.rnorm / (s.dnorm * s.xnorm));
    const double test1 = s.rnorm / s.bnorm;
    double test2 = 0.0;
    if (s.rnorm > 0.0) test2 = s.arnorm / (s.anorm * s.rnorm);
    const double test3 = 1.0 / s.acond;
    double t1 = test1 / (1.0 + s.anorm * s.xnorm / s.bnorm);
    const double rtol = s.btol + s.atol * s.anorm * s.xnorm / s.bnorm;

    // stopping tests, later assignments win (:798-810)
    const double t3 = 1.0 + test3;
    const double t2 = 1.0 + test2;
    t1 = 1.0 + t1;
    int istop = s.istop;
    if (s.itn >= s.itnlim) istop = 5;
    if (t3 <= 1.0) istop = 4;
    if (t2 <= 1.0) istop = 2;
    if (t1 <= 1.0) istop = 1;
    if (test3 <= s.ctol) istop = 4;
    if (test2 <= s.atol) istop = 2;
    if (test1 <= rtol) istop = 1;

    // nconv = 1 gate (:843-850)
    if (istop == 0) {
        s.nstop = 0;
    } else {
        const int nconv = 1;
        s.nstop = s.nstop + 1;
        if (s.nstop < nconv && s.itn < s.itnlim) istop = 0;
    }
    s.istop = istop;

    s.phi = phi;
    s.dknorm = dknorm;
    s.dxk = dxk;
    s.test1 = test1;
    s.test2 = test2;

    // snapshot of this iteration's scalars; x(1) is added when the x/w update has been applied
    s.rec.itn = (double)s.itn;
    s.rec.istop = (double)istop;
    s.rec.rnorm = s.rnorm;
    s.rec.test1 = test1;
    s.rec.test2 = test2;
    s.rec.anorm = s.anorm;
    s.rec.acond = s.acond;
    s.rec.phi = phi;
    s.rec.dknorm = dknorm;
    s.rec.dxk = dxk;
    s.rec.alfopt = s.alfopt;
    s.rec.alpha = alpha;
    s.rec.beta = beta;
    s.rec.xnorm = s.xnorm;
    s.rec.arnorm = s.arnorm;

    // coefficients of the next Aprod:  u'' = A (v'/alpha) - alpha (u'/beta)
    s.ca_mat = s.inv_alpha;
    s.ca_vec = -alpha * s.inv_beta;
}

// after the x/w update of iteration rec.itn: publish its record, close the iteration
__device__ __forceinline__ void step_after_update(DevState &s, double sum_w2, double x1,
                                                  volatile lsqr_b200_iter_record *ring)
{
    s.wnorm2 = sum_w2;
    s.x1 = x1;
    const int itn = (int)s.rec.itn;
    volatile lsqr_b200_iter_record *r = ring + (itn % kRingSize);
    r->istop = s.rec.istop;
    r->x1 = x1;
    r->rnorm = s.rec.rnorm;
    r->test1 = s.rec.test1;
    r->test2 = s.rec.test2;
    r->anorm = s.rec.anorm;
    r->acond = s.rec.acond;
    r->phi = s.rec.phi;
    r->dknorm = s.rec.dknorm;
    r->dxk = s.rec.dxk;
    r->alfopt = s.rec.alfopt;
    r->alpha = s.rec.alpha;
    r->beta = s.rec.beta;
    r->xnorm = s.rec.xnorm;
    r->arnorm = s.rec.arnorm;
    // No system-scope fence here: the host reads a record only after the event that follows the batch has
    // completed, when every write of the kernel is visible; a fence would put a PCIe round trip on the
    // critical path of every iteration.  itn doubles as the "record is the one I expect" tag.
    r->itn = s.rec.itn;
    if (s.rec.istop != 0.0) s.done = 1;
}

// =============================================================================================
// "Last block finishes" reduction.  Every block stores its partial in a fixed slot; the block
// that draws the last ticket sums the slots in index order with a fixed tree, so the result does
// not depend on block scheduling.  Returns true in thread 0 of the last block only.
// =============================================================================================
// `slots`: which partial array to use (two kernels that may run concurrently must not share one)
template <int THREADS>
__device__ __forceinline__ bool finish_reduction(DevState *st, int cslot, double thread_val,
                                                 double *smem, double *total, double *slots = nullptr)
{
    __shared__ int s_is_last;
    if (!slots) slots = st->partial;
    const double bs = block_sum<THREADS>(thread_val, smem);
    if (threadIdx.x == 0) {
        __stcg(&slots[blockIdx.x], bs);
        __threadfence();
        const unsigned int ticket = atomicAdd(&st->counter[cslot], 1u);
        s_is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_is_last) return false;
    __threadfence();
    double acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += THREADS) acc += __ldcg(&slots[i]);
    const double tot = block_sum<THREADS>(acc, smem);
    if (threadIdx.x == 0) {
        st->counter[cslot] = 0;
        *total = tot;
        return true;
    }
    return false;
}

// Two simultaneous reductions behind one ticket (sum v'^2 and sum w'^2 of the fused kernel).
template <int THREADS>
__device__ __forceinline__ bool finish_reduction2(DevState *st, int cslot, double v1, double v2,
                                                  double *smem, double *total1, double *total2)
{
    __shared__ int s_is_last2;
    const double b1 = block_sum<THREADS>(v1, smem);
    const double b2 = block_sum<THREADS>(v2, smem);
    if (threadIdx.x == 0) {
        __stcg(&st->partial[blockIdx.x], b1);
        __stcg(&st->partial2[blockIdx.x], b2);
        __threadfence();
        const unsigned int ticket = atomicAdd(&st->counter[cslot], 1u);
        s_is_last2 = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_is_last2) return false;
    __threadfence();
    double a1 = 0.0, a2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += THREADS) {
        a1 += __ldcg(&st->partial[i]);
        a2 += __ldcg(&st->partial2[i]);
    }
    const double t1 = block_sum<THREADS>(a1, smem);
    const double t2 = block_sum<THREADS>(a2, smem);
    if (threadIdx.x == 0) {
        st->counter[cslot] = 0;
        *total1 = t1;
        *total2 = t2;
        return true;
    }
    return false;
}

// =============================================================================================
// CSR views
// =============================================================================================
struct CsrView {
    const uint32_t *ptr;   // [nrows+1]
    const int32_t  *idx;   // [nnz] 0-based other coordinate
    const double   *val;   // [nnz]
    int64_t         nrows;
};

enum SpmvEpilogue {
    EPI_FUSED_APROD  = 0,  // out = ca_mat*sum + ca_vec*out, sum(out^2), step_after_aprod
    EPI_FUSED_ATPROD = 1,  // out = ct_mat*sum + ct_vec*out, sum(out^2), step_after_atprod
    EPI_INIT_ATPROD  = 2,  // out = ct_mat*sum,              sum(out^2), step_init_alpha
    EPI_ACC          = 3,  // out += sum            (plain aprod:  y = y + A x)
    EPI_STORE        = 4   // out  = sum            (multi-GPU partial A_p' u_p)
};

// ---------------------------------------------------------------------------------------------
// K3/K4, variant 1: LANES threads cooperate on one row (sub-warp per row, warp per row for
// LANES = 32).  Persistent grid; rows are visited in a fixed round-robin order.
// ---------------------------------------------------------------------------------------------
template <int LANES, int EPI>
__global__ void __launch_bounds__(kThreads)
spmv_rowgroup_kernel(CsrView A, const double *__restrict__ x, double *out, DevState *st, double *aux, int check_done)
{
    constexpr bool kFused = (EPI == EPI_FUSED_APROD || EPI == EPI_FUSED_ATPROD || EPI == EPI_INIT_ATPROD);
    __shared__ double s_red[kThreads / 32];

    double cm = 1.0, cv = 0.0;
    if (kFused) {
        if (st->done) return;
        if (EPI == EPI_FUSED_ATPROD && st->beta == 0.0) {
            // beta = 0: the reference skips the A' half and keeps alpha (src/lsqr.f90:691-699)
            if (blockIdx.x == 0 && threadIdx.x == 0) step_after_atprod(*st, 0.0, false);
            return;
        }
        if (EPI == EPI_FUSED_APROD) { cm = st->ca_mat; cv = st->ca_vec; }
        else                        { cm = st->ct_mat; cv = st->ct_vec; }
    } else if (check_done && st->done) {
        return;   // unfused STORE / ACC launched from the solve loop after the solver has stopped
    }

    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    constexpr int kRowsPerPass = kThreads / LANES;
    const int lane = threadIdx.x % LANES;
    const int sub = threadIdx.x / LANES;
    double sq = 0.0;

    for (int64_t base = (int64_t)blockIdx.x * kRowsPerPass; base < A.nrows;
         base += (int64_t)gridDim.x * kRowsPerPass) {
        const int64_t row = base + sub;
        double sum = 0.0;
        if (row < A.nrows) {
            const uint32_t p0 = A.ptr[row], p1 = A.ptr[row + 1];
#pragma unroll 4
            for (uint32_t k = p0 + lane; k < p1; k += LANES) {
                const double a = ldg_stream_f64(A.val + k, pol_stream);
                const int32_t c = ldg_stream_s32(A.idx + k, pol_stream);
                sum += a * ldg_keep_f64(x + c, pol_keep);
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0 && row < A.nrows) {
            if (EPI == EPI_ACC) {
                out[row] += sum;
            } else if (EPI == EPI_STORE) {
                out[row] = sum;
            } else if (EPI == EPI_INIT_ATPROD) {
                const double r = cm * sum;
                out[row] = r;
                sq += r * r;
            } else {
                const double r = cm * sum + cv * out[row];
                out[row] = r;
                sq += r * r;
            }
        }
    }

    if (kFused) {
        double total;
        if (finish_reduction<kThreads>(st, 0, sq, s_red, &total)) {
            if (EPI == EPI_FUSED_APROD) {
                // multi-GPU: this is only the local part of sum(u'^2); it rides to the all-reduce in
                // slot n of the A'u buffer and the step is taken by vfinish_kernel on every rank.
                if (aux) *aux = total; else step_after_aprod(*st, total);
            }
            else if (EPI == EPI_FUSED_ATPROD) step_after_atprod(*st, total, true);
            else step_init_alpha(*st, total);
        }
    }
}

// =============================================================================================
// K5: x/w(/se) update (src/lsqr.f90:729-745) + sum(w'^2); last block closes the iteration.
//   x += t1*w ;  w' = inv_alpha*v + t2*w ;  se += (t3*w)^2
// =============================================================================================
template <bool WANTSE, bool LAZY>
__global__ void __launch_bounds__(kThreads)
xw_update_kernel(int64_t n, double *__restrict__ x, double *__restrict__ w, const double *__restrict__ v,
                 double *__restrict__ se, DevState *st, volatile lsqr_b200_iter_record *ring, int vec_ok)
{
    __shared__ double s_red[kThreads / 32];
    if (st->done) return;
    const double t1 = st->t1, t2 = st->t2, t3 = st->t3;
    const double ia = LAZY ? st->inv_alpha : 1.0;   // LAZY: v is stored unnormalised
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
            sq += wn.x * wn.x + wn.y * wn.y;
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
        sq += wn * wn;
        if (WANTSE) se[i] += (t3 * wo) * (t3 * wo);
    }
    double total;
    // own partial slots and ticket: this kernel may run next to the Aprod of the following iteration
    if (finish_reduction<kThreads>(st, 1, sq, s_red, &total, st->partial2)) {
        __threadfence();
        const double x1 = __ldcg(x);   // x(1) after the update
        step_after_update(*st, total, x1, ring);
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

// On exit the reference's u and v are normalised; the fused engine keeps them unnormalised.
// Not needed by solve_ez (u, v are private workspaces there).

// =============================================================================================
// BLAS-1 kernels (src/lsqrblas.f90) for the operator-hook path and the public device BLAS.
// The reducing kernels write one double to `result` (device) from the last block.
// =============================================================================================
enum NormPost {
    POST_NONE = 0,        // result only
    POST_INIT_BETA = 1,   // step_init_beta
    POST_INIT_ALPHA = 2,  // step_init_alpha
    POST_G_BETA = 3,      // step_after_aprod   (operator-hook path: u already holds A v - alpha u)
    POST_G_ALPHA = 4      // step_after_atprod  (operator-hook path: v already holds A'u - beta v)
};

// sum of squares (dnrm2 without the serial rescaling loop; see DESIGN.md for the overflow note)
template <int POST>
__global__ void __launch_bounds__(kThreads)
sumsq_kernel(int64_t n, const double *__restrict__ x, DevState *st, double *result)
{
    __shared__ double s_red[kThreads / 32];
    if (POST >= POST_G_BETA && st->done) return;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double t = x[i];
        sq += t * t;
    }
    double total;
    if (finish_reduction<kThreads>(st, 2, sq, s_red, &total)) {
        if (POST == POST_INIT_BETA) step_init_beta(*st, total);
        if (POST == POST_INIT_ALPHA) step_init_alpha(*st, total);
        if (POST == POST_G_BETA) step_after_aprod(*st, total);
        if (POST == POST_G_ALPHA) step_after_atprod(*st, total, st->beta > 0.0);
        if (result) *result = total;
    }
}

__global__ void __launch_bounds__(kThreads)
dot_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y, DevState *st, double *result)
{
    __shared__ double s_red[kThreads / 32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        acc += x[i] * y[i];
    double total;
    if (finish_reduction<kThreads>(st, 2, acc, s_red, &total)) *result = total;
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

__global__ void __launch_bounds__(kThreads)
axpy_kernel(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        y[i] = y[i] + a * x[i];
}

// =============================================================================================
// Column-blocked A (v does not fit in L2): gu = A v has been formed block by block; finish the Aprod step
//   u' = ca_mat * gu + ca_vec * u ;  sum(u'^2)  ->  step_after_aprod (or the partial to aux, multi-GPU)
// =============================================================================================
__global__ void __launch_bounds__(kThreads)
ufinish_kernel(int64_t m, const double *__restrict__ gu, double *__restrict__ u, DevState *st, double *aux)
{
    __shared__ double s_red[kThreads / 32];
    if (st->done || st->istop != 0) return;
    const double cm = st->ca_mat, cv = st->ca_vec;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < m; i += (int64_t)gridDim.x * kThreads) {
        const double r = cm * gu[i] + cv * u[i];
        u[i] = r;
        sq += r * r;
    }
    double total;
    if (finish_reduction<kThreads>(st, 0, sq, s_red, &total)) {
        if (aux) *aux = total; else step_after_aprod(*st, total);
    }
}

// =============================================================================================
// K8 (multi-GPU): after the all-reduce of g = [ A'u' (n entries) | sum(u'^2) ] every rank forms
//   beta = sqrt(g[n]);  v' = g/beta - (beta/alpha) v;  sum(v'^2)
// redundantly (v is replicated), so alpha, the rotations and the stop decision are bit-identical
// on all ranks and need no further collective.
// =============================================================================================
template <bool INIT>
__global__ void __launch_bounds__(kThreads)
vfinish_kernel(int64_t n, const double *__restrict__ g, double *__restrict__ v, DevState *st)
{
    __shared__ double s_red[kThreads / 32];
    if (st->done) return;
    const double sumsq_u = g[n];
    const double beta = sqrt(sumsq_u);
    if (beta == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (INIT) { step_init_beta(*st, 0.0); step_init_alpha(*st, 0.0); }
            else      { step_after_aprod(*st, 0.0); step_after_atprod(*st, 0.0, false); }
        }
        return;
    }
    const double cm = 1.0 / beta;
    const double cv = INIT ? 0.0 : -beta * st->inv_alpha;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double r = INIT ? cm * g[i] : cm * g[i] + cv * v[i];
        v[i] = r;
        sq += r * r;
    }
    double total;
    if (finish_reduction<kThreads>(st, 3, sq, s_red, &total)) {
        if (INIT) { step_init_beta(*st, sumsq_u); step_init_alpha(*st, total); }
        else      { step_after_aprod(*st, sumsq_u); step_after_atprod(*st, total, true); }
    }
}

}  // namespace lsqrb
