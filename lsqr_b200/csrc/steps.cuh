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

// after ||v'||: rotations, estimates and stopping tests, src/lsqr.f90:695-721,724-726,751-810,843-850
__device__ __forceinline__ void step_after_atprod(DevState &s, double nrm, bool new_alpha)
{
    if (new_alpha) {
        s.alpha = nrm;
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

    // dknorm = sqrt(sum (t3 w_i)^2) = |t3| ||w||  (:729-751); ||w|| was produced when w was written
    const double dknorm = fabs(s.t3) * s.wnorm;
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
// result does not depend on block scheduling.  Returns true in thread 0 of the last block only.
// The exceptional accumulators are only reduced when some thread of the grid used them.
// =============================================================================================
template <int THREADS>
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
