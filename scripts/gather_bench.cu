// Microbenchmark: how many random 8-byte gathers per second does a B200 sustain, and through which path?
//   mode 0  ld.global.nc (LSU, L1 allocating)                     -- what the SpMV kernel does
//   mode 1  tex1Dfetch<int2> (texture path)
//   mode 2  ld.global.nc.L1::no_allocate
//   mode 3  half LSU + half TEX
//   mode 4  cp.async.bulk 16 B global -> shared (TMA path, UBLKCP), one mbarrier per warp, then LDS
//   mode 5  cp.async.ca 8 B global -> shared (LDGSTS), then LDS
//   mode 6  16-byte loads of aligned pairs (x[2c], x[2c+1]): the bound for a format whose gathers come in adjacent pairs
//   mode 7  random 8-byte reads of a 32 KB window staged in SHARED memory (what the windowed SpMV path does)
// The SpMV kernels are bound by this rate on matrices with random columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/gather_bench.cu -o build/gather_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

template <int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const int4 *__restrict__ idx, int64_t nquads, const double *__restrict__ x,
                                                      cudaTextureObject_t tex, double *out, int64_t nx)
{
    __shared__ __align__(16) double s_buf[(MODE == 4) ? 8 * 32 * 8 : (MODE == 5) ? 8 * 32 * 4 : 2];   // per warp staging
    __shared__ uint64_t s_bar[8];
    extern __shared__ double s_win[];                                                              // MODE 7
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (MODE == 4) {
        if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar + wib)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    if (MODE == 7) {
        for (int i = threadIdx.x; i < 4096; i += 256) s_win[i] = x[i % nx];
        __syncthreads();
    }
    uint32_t phase = 0;
    double acc = 0.0;
    // (every warp runs the same number of steps: the loop bound is warp-uniform)
    for (int64_t q0 = ((int64_t)blockIdx.x * 8 + wib) * 32; q0 < nquads; q0 += (int64_t)gridDim.x * 256) {
        const int64_t q = q0 + lane;
        int4 c = make_int4(0, 0, 0, 0);
        if (q < nquads)
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
        } else if (MODE == 2) {
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v0) : "l"(x + c.x));
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v1) : "l"(x + c.y));
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v2) : "l"(x + c.z));
            asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v3) : "l"(x + c.w));
        } else if (MODE == 4) {
            // 4 bulk copies of 16 bytes per lane (the aligned pair that holds the wanted double), one barrier per warp
            double *dst = s_buf + (wib * 32 + lane) * 8;
            uint64_t *bar = s_bar + wib;
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(32u * 4u * 16u) : "memory");
            __syncwarp();
            const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                             ::"r"(smem_u32(dst + 2 * k)), "l"(x + (cc[k] & ~1)), "r"(smem_u32(bar)) : "memory");
            const long long t0 = clock64();
            while (!mbar_try_wait(bar, phase)) { if (clock64() - t0 > 4000000000ll) __trap(); }
            phase ^= 1u;
            v0 = dst[0 + (c.x & 1)]; v1 = dst[2 + (c.y & 1)]; v2 = dst[4 + (c.z & 1)]; v3 = dst[6 + (c.w & 1)];
            __syncwarp();
        } else if (MODE == 5) {
            double *dst = s_buf + (wib * 32 + lane) * 4;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + 0)), "l"(x + c.x) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + 1)), "l"(x + c.y) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + 2)), "l"(x + c.z) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + 3)), "l"(x + c.w) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            v0 = dst[0]; v1 = dst[1]; v2 = dst[2]; v3 = dst[3];
        } else if (MODE == 6) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(x + (c.x & ~1)));
            const double2 b = __ldg(reinterpret_cast<const double2 *>(x + (c.y & ~1)));
            v0 = a.x; v1 = a.y; v2 = b.x; v3 = b.y;     // 4 doubles for 2 gathers
        } else {
            v0 = s_win[c.x & 4095]; v1 = s_win[c.y & 4095]; v2 = s_win[c.z & 4095]; v3 = s_win[c.w & 4095];
        }
        acc += (v0 + v1) + (v2 + v3);
    }
    if (acc == 123.456) out[0] = acc;   // keep the loads alive
}

template <int MODE>
static float run(int grid, const int4 *idx, int64_t nquads, const double *x, cudaTextureObject_t tex, double *out, int64_t nx)
{
    const size_t dyn = MODE == 7 ? 4096 * sizeof(double) : 0;
    if (dyn) CK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        gather_kernel<MODE><<<grid, 256, dyn>>>(idx, nquads, x, tex, out, nx);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main()
{
    const int64_t M = 1ll << 26;   // gathers per launch (256 MB of indices, streamed)
    int4 *idx; CK(cudaMalloc(&idx, M * 4));
    double *out; CK(cudaMalloc(&out, 8));
    std::vector<int32_t> h(M);
    static const char *names[8] = {"ld.global.nc", "tex1Dfetch<int2>", "ld.nc.L1::no_allocate", "half LSU + half TEX",
                                   "cp.async.bulk 16B -> smem", "cp.async.ca 8B -> smem", "16B loads of pairs", "LDS from staged window"};
    for (int64_t N : {100000ll, 6250000ll, 25000000ll}) {
        uint64_t s = 88172645463325252ull;
        for (int64_t i = 0; i < M; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int32_t)(s % (uint64_t)N); }
        CK(cudaMemcpy(idx, h.data(), M * 4, cudaMemcpyHostToDevice));
        double *x; CK(cudaMalloc(&x, N * 8 + 16)); CK(cudaMemset(x, 0, N * 8 + 16));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
        rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindSigned); rd.res.linear.sizeInBytes = N * 8;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.addressMode[0] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
        cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        for (int occ : {4, 8}) {
            for (int mode = 0; mode < 8; ++mode) {
                if (mode == 7 && (occ != 4 || N != 100000ll)) continue;   // the staged window does not depend on N
                const int grid = 148 * occ;
                float best = 0.f;
                switch (mode) {
                case 0: best = run<0>(grid, idx, M / 4, x, tex, out, N); break;
                case 1: best = run<1>(grid, idx, M / 4, x, tex, out, N); break;
                case 2: best = run<2>(grid, idx, M / 4, x, tex, out, N); break;
                case 3: best = run<3>(grid, idx, M / 4, x, tex, out, N); break;
                case 4: best = run<4>(grid, idx, M / 4, x, tex, out, N); break;
                case 5: best = run<5>(grid, idx, M / 4, x, tex, out, N); break;
                case 6: best = run<6>(grid, idx, M / 4, x, tex, out, N); break;
                default: best = run<7>(grid, idx, M / 4, x, tex, out, N); break;
                }
                const double ng = mode == 6 ? (double)M / 2 : (double)M;     // mode 6 performs M/2 16-byte gathers
                const double gps = ng / (best * 1e-3);
                printf("x = %8lld doubles (%6.1f MB)  CTAs/SM %d  %-26s %7.3f ms  %6.1f G gathers/s  %.2f gathers/clk/SM @1965 MHz\n",
                       (long long)N, N * 8 / 1e6, occ, names[mode], best, gps / 1e9, gps / 148 / 1.965e9);
                fflush(stdout);
            }
        }
        cudaDestroyTextureObject(tex); cudaFree(x);
    }
    return 0;
}
