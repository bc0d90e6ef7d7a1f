// Microbenchmark: how many random 8-byte gathers per second does a B200 sustain through (a) the LSU path
// (ld.global.nc), (b) the texture path (tex1Dfetch<int2> on a linear texture object), (c) LSU with L1 bypass?
// The SpMV kernels are bound by this rate on matrices with random columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/gather_bench.cu -o build/gather_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const int4 *__restrict__ idx, int64_t nquads, const double *__restrict__ x,
                                                      cudaTextureObject_t tex, double *out)
{
    double acc = 0.0;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < nquads; q += (int64_t)gridDim.x * 256) {
        int4 c;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(idx + q));
        double v0, v1, v2, v3;
        if (MODE == 0) {
            v0 = __ldg(x + c.x); v1 = __ldg(x + c.y); v2 = __ldg(x + c.z); v3 = __ldg(x + c.w);
        } else if (MODE == 1) {
            int2 t0 = tex1Dfetch<int2>(tex, c.x), t1 = tex1Dfetch<int2>(tex, c.y), t2 = tex1Dfetch<int2>(tex, c.z), t3 = tex1Dfetch<int2>(tex, c.w);
            v0 = __hiloint2double(t0.y, t0.x); v1 = __hiloint2double(t1.y, t1.x);
            v2 = __hiloint2double(t2.y, t2.x); v3 = __hiloint2double(t3.y, t3.x);
        } else if (MODE == 3) {   // half through the LSU path, half through the texture path
            int2 t0 = tex1Dfetch<int2>(tex, c.x), t2 = tex1Dfetch<int2>(tex, c.z);
            v1 = __ldg(x + c.y); v3 = __ldg(x + c.w);
            v0 = __hiloint2double(t0.y, t0.x); v2 = __hiloint2double(t2.y, t2.x);
        } else {
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v0) : "l"(x + c.x));
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v1) : "l"(x + c.y));
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v2) : "l"(x + c.z));
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v3) : "l"(x + c.w));
        }
        acc += (v0 + v1) + (v2 + v3);
    }
    if (acc == 123.456) out[0] = acc;   // keep the loads alive
}

int main()
{
    const int64_t M = 1ll << 26;   // gathers per launch (256 MB of indices, streamed)
    int4 *idx; CK(cudaMalloc(&idx, M * 4));
    double *out; CK(cudaMalloc(&out, 8));
    std::vector<int32_t> h(M);
    int dev_clock = 0; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
    for (int64_t N : {100000ll, 6250000ll, 25000000ll}) {
        uint64_t s = 88172645463325252ull;
        for (int64_t i = 0; i < M; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int32_t)(s % (uint64_t)N); }
        CK(cudaMemcpy(idx, h.data(), M * 4, cudaMemcpyHostToDevice));
        double *x; CK(cudaMalloc(&x, N * 8)); CK(cudaMemset(x, 0, N * 8));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
        rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindSigned); rd.res.linear.sizeInBytes = N * 8;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.addressMode[0] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
        cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        for (int occ : {4, 8}) {
            for (int mode = 0; mode < 4; ++mode) {
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                float best = 1e30f;
                for (int rep = 0; rep < 5; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) gather_kernel<0><<<148 * occ, 256>>>(idx, M / 4, x, tex, out);
                    if (mode == 1) gather_kernel<1><<<148 * occ, 256>>>(idx, M / 4, x, tex, out);
                    if (mode == 2) gather_kernel<2><<<148 * occ, 256>>>(idx, M / 4, x, tex, out);
                    if (mode == 3) gather_kernel<3><<<148 * occ, 256>>>(idx, M / 4, x, tex, out);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                }
                const double gps = (double)M / (best * 1e-3);
                printf("x = %8lld doubles (%6.1f MB)  CTAs/SM %d  %-22s %7.3f ms  %6.1f G gathers/s  %.2f gathers/clk/SM @1965 MHz\n",
                       (long long)N, N * 8 / 1e6, occ, mode == 0 ? "ld.global.nc" : mode == 1 ? "tex1Dfetch<int2>" : mode == 2 ? "ld.nc.L1::no_allocate" : "half LSU + half TEX",
                       best, gps / 1e9, gps / 148 / 1.965e9);
            }
        }
        cudaDestroyTextureObject(tex); cudaFree(x);
    }
    return 0;
}
