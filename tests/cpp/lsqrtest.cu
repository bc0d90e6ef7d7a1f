// lsqrtest.cu -- counterpart of the reference's test/lsqrtest.f90 + test/lsqrtest_module.f90 for the B200 engine:
// the 18 LSTP problems (A = HY*D*HZ, two Householder reflections around a diagonal) run through the LOW-LEVEL
// class lsqr_solver with a user operator, exactly like `type,extends(lsqr_solver) :: test_solver`
// (test/lsqrtest_module.f90:35-44).  Here the operator lives on the GPU: aprod1 / aprod2 / hprod
// (:319-403) are small CUDA kernels that receive the engine's DEVICE vectors and stream.
//
// The problem data (d, hy, hz, b, xtrue) comes from the CPU oracle's generator (oracle_lstp_generate, the
// restatement of lstp :422-505), and every problem is also solved by the oracle so that the device path is
// compared on identical inputs: acheck inform, istop, iteration count, x, and the xcheck verdict.
// The log (header, acheck / LSQR / xcheck reports, solution, verdict) is written like LSQR.LIS.
//
//   usage: lsqrtest [log file (default LSQR_B200.LIS)] [nbar (default 1000)]
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lsqr_b200.hpp"
#include "../../oracle/lsqr_oracle.h"

using namespace lsqr_module;

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e__)); std::exit(3); } } while (0)

// y = x - 2 (hz'x) hz for one vector of length n <= a few thousand: one block, fixed-tree reduction (hprod, :385-403)
__global__ void hprod_kernel(int n, const double *__restrict__ hz, const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ double red[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += hz[i] * x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double t = red[0] + red[0];
    for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = x[i] - t * hz[i];
}
// w(1:minmn) = d * w ; w(minmn+1:len) = 0
__global__ void scale_pad_kernel(int minmn, int len, const double *__restrict__ d, double *__restrict__ w)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) w[i] = i < minmn ? d[i] * w[i] : 0.0;
}
__global__ void add_kernel(int n, const double *__restrict__ w, double *__restrict__ y)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = y[i] + w[i];
}

// type,extends(lsqr_solver) :: test_solver  (test/lsqrtest_module.f90:35-44)
struct test_solver : lsqr_solver {
    int m = 0, n = 0, minmn = 0, maxmn = 0;
    double *d = nullptr, *hy = nullptr, *hz = nullptr, *w = nullptr;   // rw(locd..), device copies

    void setup(const oracle_lstp *p)
    {
        release();
        m = p->m; n = p->n; minmn = p->minmn; maxmn = p->maxmn;
        CK(cudaMalloc(&d, sizeof(double) * minmn));
        CK(cudaMalloc(&hy, sizeof(double) * m));
        CK(cudaMalloc(&hz, sizeof(double) * n));
        CK(cudaMalloc(&w, sizeof(double) * maxmn));
        CK(cudaMemcpy(d, p->d, sizeof(double) * minmn, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(hy, p->hy, sizeof(double) * m, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(hz, p->hz, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    void release()
    {
        cudaFree(d); cudaFree(hy); cudaFree(hz); cudaFree(w);
        d = hy = hz = w = nullptr;
    }
    ~test_solver() override { release(); }

    // aprod_test_solver (:283-309) -> aprod1 (:319-343) / aprod2 (:353-377)
    void aprod(int mode, int m_, int n_, wp *x, wp *y, void *stream) override
    {
        cudaStream_t s = (cudaStream_t)stream;
        const int grid = 8;
        if (mode == 1) {            // y = y + HY*D*HZ*x
            hprod_kernel<<<1, 1024, 0, s>>>(n_, hz, x, w);
            scale_pad_kernel<<<grid, 256, 0, s>>>(minmn, m_, d, w);
            hprod_kernel<<<1, 1024, 0, s>>>(m_, hy, w, w);
            add_kernel<<<grid, 256, 0, s>>>(m_, w, y);
        } else {                    // x = x + HZ*D*HY*y
            hprod_kernel<<<1, 1024, 0, s>>>(m_, hy, y, w);
            scale_pad_kernel<<<grid, 256, 0, s>>>(minmn, n_, d, w);
            hprod_kernel<<<1, 1024, 0, s>>>(n_, hz, w, w);
            add_kernel<<<grid, 256, 0, s>>>(n_, w, x);
        }
    }
};

static int g_fail = 0;

// subroutine test (:119-272)
static void test(test_solver &me, std::FILE *nout, int m, int n, int nduplc, int npower, double damp)
{
    const double eps = DBL_EPSILON;
    oracle_lstp *p = oracle_lstp_new(m, n);
    std::vector<double> xtrue((size_t)n), b((size_t)m);
    for (int j = 0; j < n; ++j) xtrue[(size_t)j] = (j + 1) * 0.1;
    double acond = 0, rnorm_gen = 0;
    oracle_lstp_generate(p, nduplc, npower, damp, ORACLE_FOURPI_F64, xtrue.data(), b.data(), &acond, &rnorm_gen);
    me.setup(p);

    std::fprintf(nout, "\n\n ------------------------------------------------------------------------\n"
                       " Least-Squares Test Problem      P(%5d%5d%5d%5d%12.2E )\n\n"
                       " Condition no. =%12.4E     Residual function =%17.9E\n"
                       " ------------------------------------------------------------------------\n",
                 m, n, nduplc, npower, damp, acond, rnorm_gen);

    double *u, *v, *w, *x, *y, *se, *bd;
    const int maxmn = m > n ? m : n;
    CK(cudaMalloc(&u, sizeof(double) * m)); CK(cudaMalloc(&v, sizeof(double) * maxmn)); CK(cudaMalloc(&w, sizeof(double) * maxmn));
    CK(cudaMalloc(&x, sizeof(double) * n)); CK(cudaMalloc(&y, sizeof(double) * m)); CK(cudaMalloc(&se, sizeof(double) * n));
    CK(cudaMalloc(&bd, sizeof(double) * m));
    CK(cudaMemcpy(bd, b.data(), sizeof(double) * m, cudaMemcpyHostToDevice));

    // Check that aprod generates y + Ax and x + A'y consistently (:183-188)
    int inform = -1;
    me.acheck(m, n, nout, eps, v, w, x, y, inform);
    if (inform > 0) { std::fprintf(nout, " Check eps and power in subroutine acheck\n"); ++g_fail; }

    // Solve the problem defined by aprod, damp and b (:195-206)
    CK(cudaMemcpy(u, bd, sizeof(double) * m, cudaMemcpyDeviceToDevice));   // dcopy(m, b, 1, u, 1)
    const bool wantse = false;
    const double atol = std::pow(eps, 0.99), btol = atol, conlim = 1000.0 * acond;
    const int itnlim = 4 * (m + n + 50);
    int istop = -1, itn = -1;
    double anorm, acond_est, rnorm, arnorm, xnorm;
    me.lsqr(m, n, damp, wantse, u, v, w, x, se, atol, btol, conlim, itnlim, nout, istop, itn, anorm, acond_est, rnorm, arnorm, xnorm);

    // Examine the results (:216-218)
    int xinform = -1;
    double test1, test2, test3;
    me.xcheck(m, n, nout, anorm, damp, eps, bd, u, v, w, x, xinform, test1, test2, test3);

    std::vector<double> xh((size_t)n);
    CK(cudaMemcpy(xh.data(), x, sizeof(double) * n, cudaMemcpyDeviceToHost));
    std::fprintf(nout, "\n\n Solution  x:\n");
    const int nprint = std::min(std::min(m, n), 8);
    for (int j = 0; j < nprint; ++j) std::fprintf(nout, "%6d%14.6G%s", j + 1, xh[(size_t)j], (j % 4 == 3 || j == nprint - 1) ? "\n" : "");

    double wn = 0, xn = 0;
    for (int j = 0; j < n; ++j) { wn += (xh[(size_t)j] - xtrue[(size_t)j]) * (xh[(size_t)j] - xtrue[(size_t)j]); xn += xtrue[(size_t)j] * xtrue[(size_t)j]; }
    const double enorm = std::sqrt(wn) / (1.0 + std::sqrt(xn));
    std::fprintf(nout, enorm <= 0.001 ? "\n LSQR  appears to be successful.     Relative error in  x  =%10.2E\n"
                                      : "\n LSQR  appears to have failed.       Relative error in  x  =%10.2E\n", enorm);

    // ---- parity with the oracle on the identical problem
    oracle_lstp_result ref;
    std::vector<double> xo((size_t)n);
    oracle_lstp_test(m, n, nduplc, npower, damp, ORACLE_FOURPI_F64, nullptr, nullptr, nullptr, nullptr, &ref, xo.data());
    double dn = 0, on = 0;
    for (int j = 0; j < n; ++j) { dn += (xh[(size_t)j] - xo[(size_t)j]) * (xh[(size_t)j] - xo[(size_t)j]); on += xo[(size_t)j] * xo[(size_t)j]; }
    const double relx = std::sqrt(dn / on);
    // These runs stop at atol = eps^0.99, i.e. inside rounding noise: even a serial restatement in a different
    // language lands 0..25 iterations away from the committed LSQR.LIS, so the exit iteration is pinned
    // only loosely (25 %) and x is compared at the accuracy class the problem's conditioning allows.  What must
    // agree exactly: istop, the acheck / xcheck verdicts and the "successful / failed" classification.
    const double xtol = 1e-6 * std::fmax(1.0, ref.acond * 1e-3);
    const bool ok = inform == ref.acheck_inform && istop == ref.istop && xinform == ref.xcheck_inform &&
                    std::abs(itn - ref.itn) <= std::max(30, ref.itn / 4) && relx <= xtol &&
                    ((enorm <= 0.001) == (ref.enorm <= 0.001));
    std::printf("P(%4d,%4d,%2d,%d) istop %d/%d itn %4d/%4d acheck %d/%d xcheck %d/%d enorm %.2e/%.2e rel x %.1e %s\n",
                m, n, nduplc, npower, istop, ref.istop, itn, ref.itn, inform, ref.acheck_inform, xinform, ref.xcheck_inform,
                enorm, ref.enorm, relx, ok ? "ok" : "MISMATCH");
    if (!ok) ++g_fail;

    cudaFree(u); cudaFree(v); cudaFree(w); cudaFree(x); cudaFree(y); cudaFree(se); cudaFree(bd);
    oracle_lstp_free(p);
}

// subroutine lsqr_test (:55-94)
int main(int argc, char **argv)
{
    const char *path = argc > 1 ? argv[1] : "LSQR_B200.LIS";
    const int nbar = argc > 2 ? std::atoi(argv[2]) : 1000;
    std::FILE *nout = std::fopen(path, "w");
    if (!nout) { std::perror(path); return 2; }
    try {
        test_solver solver;
        const int nduplc = 40;
        const int shapes[3][2] = {{2 * nbar, nbar}, {nbar, nbar}, {nbar, 2 * nbar}};
        for (auto &mn : shapes)
            for (int ndamp = 2; ndamp <= 7; ++ndamp) {
                const int npower = ndamp;
                const double damp = std::pow(10.0, -ndamp - 6);
                test(solver, nout, mn[0], mn[1], nduplc, npower, damp);
            }
    } catch (const std::exception &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        std::fclose(nout);
        return 2;
    }
    std::fclose(nout);
    std::printf(g_fail ? "FAILED (%d problems)\n" : "ALL 18 LSTP PROBLEMS MATCH THE ORACLE\n", g_fail);
    return g_fail ? 1 : 0;
}
