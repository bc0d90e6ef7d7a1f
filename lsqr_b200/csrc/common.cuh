// common.cuh -- shared declarations of the B200 LSQR engine (device state, helpers).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/lsqr_b200.h"

namespace lsqrb {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_last_error(const std::string &msg);

#define LSQRB_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t err__ = (call);                                                              \
        if (err__ != cudaSuccess) {                                                              \
            ::lsqrb::set_last_error(std::string(#call) + ": " + cudaGetErrorString(err__) +      \
                                    " (" __FILE__ ":" + std::to_string(__LINE__) + ")");         \
            return LSQR_B200_ERR_CUDA;                                                           \
        }                                                                                        \
    } while (0)

#define LSQRB_TRY(call)                                                                          \
    do {                                                                                         \
        int rc__ = (call);                                                                       \
        if (rc__ != LSQR_B200_OK) return rc__;                                                   \
    } while (0)

constexpr int kNumSMs = 148;             // B200: 2 dies x 74 SMs
constexpr int kMaxPartials = 148 * 16;   // upper bound on the grid of any reducing kernel
constexpr int kRingSize = 256;           // per-iteration records in pinned mapped memory
constexpr int kTraceSlots = 1024;

// ---------------------------------------------------------------------------------------------
// Device-resident solver state: every scalar of the LSQR recurrence (src/lsqr.f90:566-575)
// lives here, is advanced by the last block of the kernel that produces its inputs, and never
// visits the host inside the loop.
// ---------------------------------------------------------------------------------------------
// Scaled sum of squares (J. L. Blue's three-accumulator scheme, the algorithm of LAPACK 3.10's dnrm2): values in
// the mid range are squared directly, very large ones are scaled down and very small ones scaled up before
// squaring, so sqrt(sum x^2) neither overflows nor underflows -- the single-pass, order-independent (hence
// parallel and deterministic) counterpart of the reference's serial scaled recurrence src/lsqrblas.f90:136-154.
struct Ssq {
    double med, big, sml;
};
constexpr double kBlueSsml = 4.4989137945431964e+161;   // 2^537 : scale of the small accumulator
constexpr double kBlueSbig = 1.1113793747425387e-162;   // 2^-538: scale of the big accumulator
constexpr unsigned kBlueExpLo = 512u;                   // biased exponent of 2^-511 (tsml)
constexpr unsigned kBlueExpSpan = 996u;                 // mid range: 2^-511 <= |x| < 2^486 (tbig)

struct DevState {
    // configuration (set by the host before the solve)
    double damp, atol, btol, ctol;
    int    itnlim, wantse, damped, dist;
    int    tr_on, tr_n;      // LSQR_B200_TRACE: device-side timeline of the fused kernels (globaltimer)

    // Golub-Kahan scalars
    double alpha, beta, inv_alpha, inv_beta;
    // coefficients consumed by the next vector kernel (lazy normalisation: u and v are stored
    // unnormalised, the 1/beta and 1/alpha factors ride in these coefficients)
    double ca_mat, ca_vec;   // Aprod : u' = ca_mat * (A v)  + ca_vec * u
    double ct_mat, ct_vec;   // Atprod: v' = ct_mat * (A'u') + ct_vec * v
    double t1, t2, t3;       // update: x += t1 w ; w' = inv_alpha * v' + t2 w ; se += (t3 w)^2
    double wnorm;            // ||w|| of the current w (gives dknorm = |t3| wnorm)
    // coefficients of the reference-structure path, which keeps u and v normalised like the reference:
    // dscal(u,-alpha) | dscal(u,1/beta) | dscal(v,-beta) | dscal(v,1/alpha)  (src/lsqr.f90:681,692,693,697)
    double g_c0, g_c1, g_c2, g_c3;

    // QR / estimate recurrences (src/lsqr.f90:597-617, 650-653)
    double rhobar, phibar, bnorm, anorm, acond, dnorm, res2, psi, xnorm, xnorm1, cs2, sn2, z;
    double rnorm, arnorm, dxmax;
    // per-iteration print-only values
    double phi, dknorm, dxk, alfopt, test1, test2, x1;

    int itn, istop, nstop, maxdx;
    int done;                // set after the x/w update of the stopping iteration
    int hook_lazy;           // operator-hook path: u, v are kept unnormalised (fused hook loop)

    // multi-GPU exchange (peer-memory path): iteration epoch the flags are compared with, error latch of a timed-out wait
    unsigned int epoch;
    int comm_error;
    int guard_error;         // a guarded multi-block launch waited for seconds on the other warps of its grid
    unsigned int probe_count;   // residency probe of a persistent grid (initialize): CTAs that have arrived ...
    int probe_fail;             // ... and whether one of them gave up waiting for the rest
    Ssq usq_local, vsq_local, wsq_local;   // this rank's partial sums of squares of u', v' (slice), w' (slice)

    // scalars of the iteration whose x/w update is still outstanding; published to the host ring
    // (with x(1)) by step_after_update
    lsqr_b200_iter_record rec;

    // completion counters of the "last block finishes the reduction" pattern, and "an exceptional (scaled) value was
    // seen" flags of the same slots
    unsigned int counter[4];
    unsigned int exc_flag[4];
    // multi-block SpMV launches: warps that have finished block b (bounds the drift between the warps of one launch)
    unsigned int blk_done[64];

    // partial sums of the reducing kernels (fixed slot per block => deterministic final sum): [0] mid-range squares,
    // [1] scaled big, [2] scaled small
    double partial[3][kMaxPartials];
    double partial2[3][kMaxPartials];   // the x/w update runs next to the following Aprod: its own slots

    // trace[0]: first instruction of block 0, trace[1]: last block has the final sums, trace[2]: scalar step done
    unsigned long long trace[4][kTraceSlots];
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// sqrt(a^2+b^2) with the reference's scaling, src/lsqr.f90:1164-1179
__device__ __forceinline__ double d2norm(double a, double b)
{
    double scale = fabs(a) + fabs(b);
    if (scale == 0.0) return 0.0;
    double p = a / scale, q = b / scale;
    return scale * sqrt(p * p + q * q);
}

// ---- Blue's scaled sum of squares -----------------------------------------------------------------------------
// exc[0] / exc[stride] are the calling thread's private big / small accumulators.  They live in shared memory so that
// they cost no registers in the kernels' hot loops, and the rare path is inline code in a branch (a real call would
// put the whole kernel under the ABI's register conventions: measured 40 % slower fused SpMV kernels).
__device__ __forceinline__ void ssq_add(double &med, double *exc, int stride, double r)
{
#ifdef LSQRB_EXPERIMENT_PLAIN_SSQ   // measurement only (cost of the scaling test): NOT overflow-safe, never shipped
    med += r * r; (void)exc; (void)stride; return;
#endif
    const unsigned e = ((unsigned)__double2hiint(r) >> 20) & 0x7ffu;
    if (e - kBlueExpLo <= kBlueExpSpan) {                    // the common case: one subtract + compare
        med += r * r;
    } else if (e == 0x7ffu) {
        med += r * r;                                        // Inf / NaN propagate
    } else if (e > kBlueExpLo + kBlueExpSpan) {
        const double t = r * kBlueSbig;
        exc[0] += t * t;
    } else if (r != 0.0) {                                   // tiny or subnormal (zero adds nothing)
        const double t = r * kBlueSsml;
        exc[stride] += t * t;
    }
}
// sqrt(sum x^2) from the three accumulators (the combination step of LAPACK 3.10 dnrm2)
__device__ __host__ inline double ssq_norm(const Ssq &a)
{
    double med = a.med;
    if (a.big > 0.0) {
        double big = a.big;
        if (med > 0.0 || med != med) big += (med * kBlueSbig) * kBlueSbig;
        return sqrt(big) / kBlueSbig;
    }
    if (a.sml > 0.0) {
        if (med > 0.0 || med != med) {
            med = sqrt(med);
            const double sm = sqrt(a.sml) / kBlueSsml;
            const double ymin = med < sm ? med : sm, ymax = med < sm ? sm : med;
            return ymax * sqrt(1.0 + (ymin / ymax) * (ymin / ymax));
        }
        return sqrt(a.sml) / kBlueSsml;
    }
    return sqrt(med);
}
__device__ __host__ inline Ssq ssq_sum(const Ssq &a, const Ssq &b) { return Ssq{a.med + b.med, a.big + b.big, a.sml + b.sml}; }

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed reduction tree (deterministic).  Result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *smem /* >= THREADS/32 doubles */)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < THREADS / 32) ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Cache-policy helpers.  The matrix streams are read exactly once per kernel: keep them out of L1
// (no_allocate) and mark them evict-first in L2 so the gathered dense vector (evict-last) stays
// resident in the 126 MB L2.  sm_100a accepts the direct .L2::evict_* qualifier only on 256-bit
// loads (LDG.E.256); narrower loads carry a createpolicy descriptor instead.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ldg_stream_f64(const double *p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int32_t ldg_stream_s32(const int32_t *p, uint64_t pol)
{
    int32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
// 256-bit streaming loads (Blackwell LDG.E.256): 4 doubles / 8 int32 per thread, 32-byte aligned.
__device__ __forceinline__ void ldg_stream_f64x4(const double *p, double (&v)[4])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ldg_stream_s32x8(const int32_t *p, int32_t (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// Bulk L2 prefetch (sm_90+): pulls `bytes` (multiple of 16, 16-byte aligned) from HBM into L2 without holding registers.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
// Gathered dense vector: read-only path, keep in L2.
__device__ __forceinline__ double ldg_keep_f64(const double *p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}

}  // namespace lsqrb
