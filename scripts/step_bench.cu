// Microbenchmark of the single-thread scalar steps (K6) on the device: how long does the last block's thread 0 take?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include scripts/step_bench.cu -o build/step_bench
#include <cstdio>
#include <vector>
#include "../lsqr_b200/csrc/kernels.cuh"
using namespace lsqrb;
namespace lsqrb { void set_last_error(const std::string &) {} }

__global__ void bench(DevState *st, volatile lsqr_b200_iter_record *ring, unsigned long long *t, int reps)
{
    for (int i = 0; i < reps; ++i) {
        const unsigned long long t0 = globaltimer_ns();
        step_after_aprod(*st, 1.0 + 1e-3 * i);
        __threadfence();
        const unsigned long long t1 = globaltimer_ns();
        step_after_atprod(*st, 2.0 + 1e-3 * i, true);
        __threadfence();
        const unsigned long long t2 = globaltimer_ns();
        step_after_update(*st, 3.0, 0.5, ring);
        __threadfence();
        const unsigned long long t3 = globaltimer_ns();
        st->done = 0; st->istop = 0;
        t[3 * i] = t1 - t0; t[3 * i + 1] = t2 - t1; t[3 * i + 2] = t3 - t2;
    }
}

int main()
{
    const int reps = 200;
    DevState *st; cudaMalloc(&st, sizeof(DevState)); cudaMemset(st, 0, sizeof(DevState));
    DevState h; memset(&h, 0, offsetof(DevState, partial));
    h.damp = 1e-3; h.damped = 1; h.atol = h.btol = 1e-10; h.ctol = 1e-8; h.itnlim = 1 << 30; h.cs2 = -1; h.inv_alpha = h.inv_beta = 1;
    h.alpha = 1.3; h.beta = 0.7; h.rhobar = 1.3; h.phibar = 0.7; h.bnorm = 0.7; h.wnorm2 = 1.0;
    cudaMemcpy(st, &h, offsetof(DevState, partial), cudaMemcpyHostToDevice);
    lsqr_b200_iter_record *ring_h, *ring_d;
    cudaHostAlloc(&ring_h, sizeof(lsqr_b200_iter_record) * kRingSize, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&ring_d, ring_h, 0);
    lsqr_b200_iter_record *ring_dev; cudaMalloc(&ring_dev, sizeof(lsqr_b200_iter_record) * kRingSize);
    unsigned long long *t; cudaMalloc(&t, sizeof(unsigned long long) * 3 * reps);
    std::vector<unsigned long long> th(3 * reps);
    for (int pass = 0; pass < 2; ++pass) {
        bench<<<1, 1>>>(st, pass == 0 ? ring_d : ring_dev, t, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(th.data(), t, sizeof(unsigned long long) * 3 * reps, cudaMemcpyDeviceToHost);
        double a[3] = {0, 0, 0};
        for (int i = reps / 2; i < reps; ++i) for (int k = 0; k < 3; ++k) a[k] += (double)th[3 * i + k];
        printf("ring in %s: step_after_aprod %.0f ns, step_after_atprod %.0f ns, step_after_update %.0f ns\n",
               pass == 0 ? "pinned mapped host memory" : "device memory", a[0] / (reps / 2), a[1] / (reps / 2), a[2] / (reps / 2));
    }
    return 0;
}
