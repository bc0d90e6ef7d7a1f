// synth.cu -- K9: device-side generators of the BASELINE.json synthetic workloads
// (include/lsqr_b200_synth.h).  Bit-identical to lsqr_b200/synth.py: every entry is a pure
// function of (seed, tag, global row, slot) through a splitmix64-style counter hash, all integer
// steps are exact and the few floating-point steps are single correctly rounded operations.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "../../include/lsqr_b200_synth.h"

namespace lsqrb {
namespace {

constexpr int kT = 256;
constexpr uint64_t kM1 = 0xBF58476D1CE4E5B9ull;
constexpr uint64_t kM2 = 0x94D049BB133111EBull;
constexpr uint64_t kGold = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kRowMul = 0xD1342543DE82EF95ull;
enum { TAG_COL = 1, TAG_VAL = 2, TAG_LEN = 3 };

__host__ __device__ __forceinline__ uint64_t mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * kM1;
    z = (z ^ (z >> 27)) * kM2;
    return z ^ (z >> 31);
}
// base = mix(seed + tag*GOLD) is formed once on the host
__device__ __forceinline__ uint64_t hash_rk(uint64_t base, uint64_t r, uint64_t k)
{
    return mix(mix(base + r * kRowMul) + k);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

inline int grid_for(int64_t n)
{
    int64_t b = (n + kT - 1) / kT;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// lens[i] = #{ j : u_i <= table[j] }, table decreasing  (== numpy.searchsorted(-table, -u, 'right'))
__global__ void __launch_bounds__(kT)
powerlaw_len_kernel(uint64_t base, int64_t row0, int64_t nrows, const double *__restrict__ table, int ntable, int64_t *lens)
{
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * kT) {
        const double u = u01(hash_rk(base, (uint64_t)(row0 + i), 0));
        int lo = 0, hi = ntable;   // first j with table[j] < u
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (table[mid] < u) hi = mid; else lo = mid + 1;
        }
        lens[i] = lo;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) lens[nrows] = 0;
}

__global__ void __launch_bounds__(kT) fixed_ptr_kernel(int64_t nrows, int64_t k, int64_t *ptr)
{
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i <= nrows; i += (int64_t)gridDim.x * kT) ptr[i] = i * k;
}

__global__ void __launch_bounds__(kT)
fill_kernel(int kind, uint64_t base_col, uint64_t base_val, uint64_t m, uint64_t n, int64_t row0, int64_t nrows,
            const int64_t *__restrict__ ptr, int64_t nnz, int32_t *__restrict__ irow, int32_t *__restrict__ icol,
            double *__restrict__ a)
{
    for (int64_t e = (int64_t)blockIdx.x * kT + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * kT) {
        // row of entry e: last r with ptr[r] <= e
        int64_t lo = 0, hi = nrows;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (ptr[mid] <= e) lo = mid; else hi = mid;
        }
        const int64_t p0 = ptr[lo];
        const uint64_t slot = (uint64_t)(e - p0);
        const uint64_t grow = (uint64_t)(row0 + lo);
        const uint64_t hc = hash_rk(base_col, grow, slot);
        const uint64_t hv = hash_rk(base_val, grow, slot);
        uint64_t col;
        double val;
        if (kind == LSQR_B200_SYNTH_BANDED) {
            const uint64_t center = (grow * n) / m;
            const uint64_t off = ((hc >> 32) * 201ull) >> 32;   // 0..200
            col = (center + n + off - 100ull) % n;
            val = u01(hv);
        } else {
            col = ((hc >> 32) * n) >> 32;
            val = 2.0 * u01(hv) - 1.0;
            if (kind == LSQR_B200_SYNTH_POWERLAW) val = val / sqrt((double)(ptr[lo + 1] - p0));
        }
        irow[e] = (int32_t)(lo + 1);
        icol[e] = (int32_t)(col + 1);
        a[e] = val;
    }
}

__global__ void __launch_bounds__(kT)
vector_kernel(uint64_t base, double coef, int64_t offset, int64_t count, double *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < count; i += (int64_t)gridDim.x * kT) {
        const double u = u01(hash_rk(base, (uint64_t)(offset + i), 0));
        out[i] = coef * (2.0 * u - 1.0);
    }
}

inline uint64_t base_of(uint64_t seed, int tag) { return mix(seed + (uint64_t)tag * kGold); }

}  // namespace
}  // namespace lsqrb

using namespace lsqrb;

extern "C" {

int lsqr_b200_synth_row_ptr(int32_t kind, uint64_t seed, int64_t row0, int64_t nrows, int32_t k,
                            const double *table_host, int32_t ntable,
                            int64_t *ptr_dev, int64_t *nnz_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nrows < 0 || !ptr_dev || !nnz_out || kind < 0 || kind > 2) return LSQR_B200_ERR_ARG;
    if (lsqr_b200_device_count() == 0) return LSQR_B200_ERR_NO_DEVICE;
    if (kind != LSQR_B200_SYNTH_POWERLAW) {
        fixed_ptr_kernel<<<grid_for(nrows + 1), kT, 0, stream>>>(nrows, k, ptr_dev);
        LSQRB_CUDA(cudaGetLastError());
        LSQRB_CUDA(cudaStreamSynchronize(stream));
        *nnz_out = nrows * (int64_t)k;
        return LSQR_B200_OK;
    }
    if (!table_host || ntable < 1) return LSQR_B200_ERR_ARG;
    double *d_table = nullptr;
    LSQRB_CUDA(cudaMalloc(&d_table, sizeof(double) * (size_t)ntable));
    LSQRB_CUDA(cudaMemcpyAsync(d_table, table_host, sizeof(double) * (size_t)ntable, cudaMemcpyHostToDevice, stream));
    powerlaw_len_kernel<<<grid_for(nrows), kT, 0, stream>>>(base_of(seed, TAG_LEN), row0, nrows, d_table, ntable, ptr_dev);
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    LSQRB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, ptr_dev, ptr_dev, nrows + 1, stream));
    LSQRB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
    LSQRB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, ptr_dev, ptr_dev, nrows + 1, stream));
    LSQRB_CUDA(cudaMemcpyAsync(nnz_out, ptr_dev + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    LSQRB_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_tmp);
    cudaFree(d_table);
    return LSQR_B200_OK;
}

int lsqr_b200_synth_fill(int32_t kind, uint64_t seed, int64_t m, int64_t n, int64_t row0, int64_t nrows,
                         const int64_t *ptr_dev, int64_t nnz,
                         int32_t *irow_dev, int32_t *icol_dev, double *a_dev, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (kind < 0 || kind > 2 || m < 1 || n < 1 || nrows < 0 || nnz < 0 || !ptr_dev) return LSQR_B200_ERR_ARG;
    if (nnz > 0 && (!irow_dev || !icol_dev || !a_dev)) return LSQR_B200_ERR_ARG;
    if (lsqr_b200_device_count() == 0) return LSQR_B200_ERR_NO_DEVICE;
    if (nnz == 0) return LSQR_B200_OK;
    fill_kernel<<<grid_for(nnz), kT, 0, stream>>>(kind, base_of(seed, TAG_COL), base_of(seed, TAG_VAL), (uint64_t)m, (uint64_t)n,
                                                  row0, nrows, ptr_dev, nnz, irow_dev, icol_dev, a_dev);
    LSQRB_CUDA(cudaGetLastError());
    LSQRB_CUDA(cudaStreamSynchronize(stream));
    return LSQR_B200_OK;
}

int lsqr_b200_synth_vector(uint64_t seed, int32_t tag, double coef, int64_t offset, int64_t count,
                           double *out_dev, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (count < 0 || (count > 0 && !out_dev)) return LSQR_B200_ERR_ARG;
    if (lsqr_b200_device_count() == 0) return LSQR_B200_ERR_NO_DEVICE;
    if (count == 0) return LSQR_B200_OK;
    vector_kernel<<<grid_for(count), kT, 0, stream>>>(base_of(seed, tag), coef, offset, count, out_dev);
    LSQRB_CUDA(cudaGetLastError());
    LSQRB_CUDA(cudaStreamSynchronize(stream));
    return LSQR_B200_OK;
}

}  // extern "C"
