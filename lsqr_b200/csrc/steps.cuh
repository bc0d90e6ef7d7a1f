// steps.cuh -- K6: the scalar recurrence of LSQR on the device, and the deterministic grid reduction that feeds it.
//
// Every function named step_* is executed by exactly ONE thread (thread 0 of the block that draws the last ticket of
// the kernel producing its input).  They restate src/lsqr.f90:632-653 (initial bidiagonalisation), :683-693 (beta,
// anorm), :695-721 (alpha, the two plane rotations), :724-726 (update coefficients), :751-810 (estimates, stopping
// rules) and :843-850 (nconv gate), in that order and with the reference's association order.
#pragma once

#include "common.cuh"

namespace lsqrb {

constexpr int kThreads = 256;
#ifdef LSQRB_EXPERIMENT_SERIAL_STEP   // A/B only: the fused A'u kernel runs its scalar step in one thread
constexpr bool kWarpStep = false;
#else
constexpr bool kWarpStep = true;
#endif

// after ||b||:  src/lsqr.f90:632-636
__device__ __forceinline__ void step_init_beta(DevState &s, double nrm)
{
    s.beta = nrm;
    s.inv_beta = s.beta > 0.0 ? 1.0 / s.beta : 1.0;
    s.alpha = 0.0;
    s.inv_alpha = 1.0;
    s.ct_mat = s.inv_beta;   // v = A'(u/beta)
    s.ct_vec = 0.0;
    s.g_c1 = s.inv_beta;
}

// after ||A'u||:  src/lsqr.f90:637-653
__device__ __forceinline__ void step_init_alpha(DevState &s, double nrm)
{
    s.alpha = (s.beta > 0.0) ? nrm : 0.0;
    s.inv_alpha = s.alpha > 0.0 ? 1.0 / s.alpha : 1.0;
    s.arnorm = s.alpha * s.beta;
    if (s.arnorm != 0.0) {
        s.rhobar = s.alpha;
        s.phibar = s.beta;
        s.bnorm = s.beta;
        s.rnorm = s.beta;
        s.ca_mat = s.inv_alpha;
        s.ca_vec = -s.alpha * s.inv_beta;
        s.wnorm = nrm * s.inv_alpha;   // w = v/alpha
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
__device__ __forceinline__ void step_after_aprod(DevState &s, double nrm)
{
    s.itn += 1;
    const double beta = nrm;
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

// ---- lane-parallel divisions and square roots --------------------------------------------------------------------
// The step after A'u is a chain of ~36 FP64 divisions and ~15 square roots (each an inline Newton sequence with a
// branch, so one thread runs them strictly one after the other: measured 8-14 us per iteration, a tenth of a C2
// iteration).  Its dependency graph is only ~9 levels deep, so the fused SpMV kernel hands it to a whole warp: every
// lane holds every scalar, the K independent quotients (roots) of one level are computed by lanes 0..K-1 and broadcast
// by shuffles.  Division and square root are correctly rounded and everything else is the same source code, so the
// one-thread form (WARP = false; the small kernels of the other paths) produces the same bits.
template <bool WARP, int K>
__device__ __forceinline__ void par_div(int lane, const double (&n)[K], const double (&d)[K], double (&r)[K])
{
    if (WARP) {
        double nn = 1.0, dd = 1.0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (lane == k) { nn = n[k]; dd = d[k]; }
        const double rr = nn / dd;
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = __shfl_sync(0xffffffffu, rr, k);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = n[k] / d[k];
    }
}

template <bool WARP, int K>
__device__ __forceinline__ void par_sqrt(int lane, const double (&a)[K], double (&r)[K])
{
    if (WARP) {
        double aa = 1.0;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (lane == k) aa = a[k];
        const double rr = sqrt(aa);
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = __shfl_sync(0xffffffffu, rr, k);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = sqrt(a[k]);
    }
}

// d2norm (src/lsqr.f90:1164-1179) in two levels: scale and the two quotients, then scale * sqrt(p^2 + q^2).
// A zero scale divides by one instead: p = q = 0 and the result is the reference's 0.
struct Hyp {
    double scale, den;
    __device__ __forceinline__ Hyp(double a, double b) : scale(fabs(a) + fabs(b)) { den = scale == 0.0 ? 1.0 : scale; }
    __device__ __forceinline__ double arg(double p, double q) const { return p * p + q * q; }
    __device__ __forceinline__ double fin(double root) const { return scale * root; }
};

// after ||v'||: rotations, estimates and stopping tests, src/lsqr.f90:695-721,724-726,751-810,843-850
// WARP: called by all 32 lanes of one warp (every lane computes every scalar, lane 0 stores).
template <bool WARP>
__device__ __forceinline__ void step_after_atprod_impl(DevState &s, double nrm, bool new_alpha, int lane)
{
    const bool writer = !WARP || lane == 0;
    const double alpha = new_alpha ? nrm : s.alpha;
    const double beta = s.beta;
    const double damp = s.damp, wnorm = s.wnorm, anorm = s.anorm, bnorm = s.bnorm;
    const double rhobar0 = s.rhobar, dnorm0 = s.dnorm, xnorm10 = s.xnorm1, z0 = s.z, sn20 = s.sn2, cs20 = s.cs2;
    const double res20 = s.res2, atol = s.atol, btol = s.btol, ctol = s.ctol;
    const int itn = s.itn, itnlim = s.itnlim, damped = s.damped;
    double phibar = s.phibar, psi = s.psi, inv_alpha = s.inv_alpha;
    int istop = s.istop, nstop = s.nstop;
    const bool want_ia = new_alpha && alpha > 0.0;

    // plane rotation that removes damp (:703-710)
    double rhbar1 = rhobar0;
    if (damped) {
        const Hyp h(rhobar0, damp);
        double r3[3], r1[1], r2[2];
        par_div<WARP, 3>(lane, {rhobar0, damp, 1.0}, {h.den, h.den, want_ia ? alpha : 1.0}, r3);
        par_sqrt<WARP, 1>(lane, {h.arg(r3[0], r3[1])}, r1);
        rhbar1 = h.fin(r1[0]);
        par_div<WARP, 2>(lane, {rhobar0, damp}, {rhbar1, rhbar1}, r2);
        const double cs1 = r2[0], sn1 = r2[1];
        psi = sn1 * phibar;
        phibar = cs1 * phibar;
        if (new_alpha) inv_alpha = want_ia ? r3[2] : 1.0;        // alpha = 0: v is left unscaled (:696-698)
    }

    // plane rotation that removes the subdiagonal beta (:714-721); res2 (:778) only needs psi
    const Hyp hr(rhbar1, beta), hs(res20, psi);
    double q5[5], w2[2];
    par_div<WARP, 5>(lane, {rhbar1, beta, res20, psi, 1.0}, {hr.den, hr.den, hs.den, hs.den, (want_ia && !damped) ? alpha : 1.0}, q5);
    if (new_alpha && !damped) inv_alpha = want_ia ? q5[4] : 1.0;
    par_sqrt<WARP, 2>(lane, {hr.arg(q5[0], q5[1]), hs.arg(q5[2], q5[3])}, w2);
    const double rho = hr.fin(w2[0]);
    const double res2 = hs.fin(w2[1]);
    double q3[3];
    par_div<WARP, 3>(lane, {rhbar1, beta, 1.0}, {rho, rho, rho}, q3);
    const double cs = q3[0], sn = q3[1], t3u = q3[2];
    const double theta = sn * alpha;
    const double rhobar = -cs * alpha;
    const double phi = cs * phibar;
    phibar = sn * phibar;
    const double tau = sn * phi;

    // dknorm = sqrt(sum (t3 w_i)^2) = |t3| ||w||  (:729-751); ||w|| was produced when w was written
    const double dknorm = fabs(t3u) * wnorm;
    const double dxk = fabs(phi * dknorm);
    // right rotation (:762-766)
    const double delta = sn20 * rho;
    const double gambar = -cs20 * rho;
    const double rhs = phi - delta * z0;
    // update coefficients (:724-726), zbar, and the quotients of dnorm (:751), gamma (:767), rnorm (:779)
    const Hyp hd(dnorm0, dknorm), hg(gambar, theta), hn(res2, phibar);
    double q9[9];
    par_div<WARP, 9>(lane, {phi, -theta, rhs, dnorm0, dknorm, gambar, theta, res2, phibar},
                     {rho, rho, gambar, hd.den, hd.den, hg.den, hg.den, hn.den, hn.den}, q9);
    const double t1u = q9[0], t2u = q9[1], zbar = q9[2];
    const Hyp hx(xnorm10, zbar);
    double qx[2], w4[4];
    par_div<WARP, 2>(lane, {xnorm10, zbar}, {hx.den, hx.den}, qx);
    par_sqrt<WARP, 4>(lane, {hd.arg(q9[3], q9[4]), hg.arg(q9[5], q9[6]), hn.arg(q9[7], q9[8]), hx.arg(qx[0], qx[1])}, w4);
    const double dnorm = hd.fin(w4[0]);
    const double gamma = hg.fin(w4[1]);
    const double rnorm = hn.fin(w4[2]);
    const double xnorm = hx.fin(w4[3]);      // estimate of norm(x) (:767)

    // estimates (:776-790) and the quotients of the stopping tests (:791-797)
    const double acond = anorm * dnorm;
    const double arnorm = alpha * fabs(tau);
    const bool has_r = rnorm > 0.0;
    double q8[8];
    par_div<WARP, 8>(lane, {gambar, theta, rhs, rnorm, has_r ? arnorm : 1.0, 1.0, rnorm, anorm * xnorm},
                     {gamma, gamma, gamma, bnorm, has_r ? anorm * rnorm : 1.0, acond, dnorm * xnorm, bnorm}, q8);
    const double cs2 = q8[0], sn2 = q8[1], z = q8[2];
    const double test1 = q8[3];
    const double test2 = has_r ? q8[4] : 0.0;
    const double test3 = q8[5];
    const Hyp h1(xnorm10, z);
    double q4[4], w2b[2];
    par_div<WARP, 4>(lane, {xnorm10, z, test1, atol * anorm * xnorm}, {h1.den, h1.den, 1.0 + q8[7], bnorm}, q4);
    par_sqrt<WARP, 2>(lane, {h1.arg(q4[0], q4[1]), q8[6]}, w2b);
    const double xnorm1 = h1.fin(w2b[0]);
    const double alfopt = w2b[1];
    double t1 = q4[2];
    const double rtol = btol + q4[3];

    // stopping tests, later assignments win (:798-810)
    const double t3 = 1.0 + test3;
    const double t2 = 1.0 + test2;
    t1 = 1.0 + t1;
    if (itn >= itnlim) istop = 5;
    if (t3 <= 1.0) istop = 4;
    if (t2 <= 1.0) istop = 2;
    if (t1 <= 1.0) istop = 1;
    if (test3 <= ctol) istop = 4;
    if (test2 <= atol) istop = 2;
    if (test1 <= rtol) istop = 1;

    // nconv = 1 gate (:843-850)
    if (istop == 0) {
        nstop = 0;
    } else {
        const int nconv = 1;
        nstop = nstop + 1;
        if (nstop < nconv && itn < itnlim) istop = 0;
    }
    if (!writer) return;

    if (new_alpha) { s.alpha = alpha; s.inv_alpha = inv_alpha; }
    s.g_c3 = want_ia ? inv_alpha : 1.0;
    s.g_c0 = -alpha;
    s.psi = psi;
    s.phibar = phibar;
    s.rhobar = rhobar;
    s.t1 = t1u;
    s.t2 = t2u;
    s.t3 = t3u;
    s.dnorm = dnorm;
    if (s.dxmax < dxk) {
        s.dxmax = dxk;
        s.maxdx = itn;
    }
    s.xnorm = xnorm;
    s.cs2 = cs2;
    s.sn2 = sn2;
    s.z = z;
    s.xnorm1 = xnorm1;
    s.acond = acond;
    s.res2 = res2;
    s.rnorm = rnorm;
    s.arnorm = arnorm;
    s.alfopt = alfopt;
    s.nstop = nstop;
    s.istop = istop;
    s.phi = phi;
    s.dknorm = dknorm;
    s.dxk = dxk;
    s.test1 = test1;
    s.test2 = test2;

    // snapshot of this iteration's scalars; x(1) is added when the x/w update has been applied
    s.rec.itn = (double)itn;
    s.rec.istop = (double)istop;
    s.rec.rnorm = rnorm;
    s.rec.test1 = test1;
    s.rec.test2 = test2;
    s.rec.anorm = anorm;
    s.rec.acond = acond;
    s.rec.phi = phi;
    s.rec.dknorm = dknorm;
    s.rec.dxk = dxk;
    s.rec.alfopt = alfopt;
    s.rec.alpha = alpha;
    s.rec.beta = beta;
    s.rec.xnorm = xnorm;
    s.rec.arnorm = arnorm;

    // coefficients of the next Aprod:  u'' = A (v'/alpha) - alpha (u'/beta)
    s.ca_mat = inv_alpha;
    s.ca_vec = -alpha * s.inv_beta;
}

// one-thread form (the small kernels of the NCCL, peer and operator-hook paths)
__device__ __forceinline__ void step_after_atprod(DevState &s, double nrm, bool new_alpha)
{
    step_after_atprod_impl<false>(s, nrm, new_alpha, 0);
}

// warp form: all 32 lanes of one warp call it with the same arguments
__device__ __forceinline__ void step_after_atprod_warp(DevState &s, double nrm, bool new_alpha, int lane)
{
    step_after_atprod_impl<true>(s, nrm, new_alpha, lane);
}

// after the x/w update of iteration rec.itn: publish its record, close the iteration
__device__ __forceinline__ void step_after_update(DevState &s, double wnorm, double x1,
                                                  volatile lsqr_b200_iter_record *ring)
{
    s.wnorm = wnorm;
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
// "Last block finishes" reduction of a scaled sum of squares.  Every thread brings its mid-range
// accumulator (a register) and its two exceptional accumulators (shared memory, exc[tid] = big,
// exc[THREADS + tid] = small; see ssq_add).  Every block stores its partial triple in a fixed slot;
// the block that draws the last ticket sums the slots in index order with a fixed tree, so the
// result does not depend on block scheduling.  Returns true in thread 0 of the last block only (WARP0: in all 32
// lanes of its first warp, each holding the total, for the lane-parallel scalar step).
// The exceptional accumulators are only reduced when some thread of the grid used them.
// =============================================================================================
template <int THREADS, bool WARP0 = false>
__device__ __forceinline__ bool finish_ssq(DevState *st, int cslot, double (*slots)[kMaxPartials],
                                           double med, const double *exc, double *smem, Ssq *total)
{
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int any_exc = __syncthreads_or(exc[tid] != 0.0 || exc[THREADS + tid] != 0.0);
    const double bmed = block_sum<THREADS>(med, smem);
    double bbig = 0.0, bsml = 0.0;
    if (any_exc) {
        bbig = block_sum<THREADS>(exc[tid], smem);
        bsml = block_sum<THREADS>(exc[THREADS + tid], smem);
    }
    if (tid == 0) {
        __stcg(&slots[0][blockIdx.x], bmed);
        if (any_exc) {
            __stcg(&slots[1][blockIdx.x], bbig);
            __stcg(&slots[2][blockIdx.x], bsml);
            atomicOr(&st->exc_flag[cslot], 1u);
        }
        __threadfence();
        const unsigned int ticket = atomicAdd(&st->counter[cslot], 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    const bool exc_any = *(volatile unsigned int *)&st->exc_flag[cslot] != 0u;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int i = tid; i < (int)gridDim.x; i += THREADS) a0 += __ldcg(&slots[0][i]);
    a0 = block_sum<THREADS>(a0, smem);
    if (exc_any) {
        // blocks that saw no exceptional value did not write their big / small slots: the slots are kept at zero
        // between launches (re-zeroed below), so summing all of them is exact
        for (int i = tid; i < (int)gridDim.x; i += THREADS) { a1 += __ldcg(&slots[1][i]); a2 += __ldcg(&slots[2][i]); }
        a1 = block_sum<THREADS>(a1, smem);
        a2 = block_sum<THREADS>(a2, smem);
        for (int i = tid; i < (int)gridDim.x; i += THREADS) { __stcg(&slots[1][i], 0.0); __stcg(&slots[2][i], 0.0); }
    }
    if (tid == 0) {
        st->counter[cslot] = 0;
        st->exc_flag[cslot] = 0;
    }
    if (tid == 0 || (WARP0 && tid < 32)) {     // block_sum leaves the sum in every lane of warp 0
        total->med = a0; total->big = a1; total->sml = a2;
        return true;
    }
    return false;
}

// The same ticket without a reduction: true in thread 0 of the block that finishes last (multi-block launches use it
// to reset their drift counters).
__device__ __forceinline__ bool last_block_ticket(DevState *st, int cslot)
{
    __shared__ int s_last_t;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(&st->counter[cslot], 1u);
        s_last_t = (ticket == gridDim.x - 1);
        if (s_last_t) st->counter[cslot] = 0;
    }
    __syncthreads();
    return s_last_t && threadIdx.x == 0;
}

}  // namespace lsqrb
