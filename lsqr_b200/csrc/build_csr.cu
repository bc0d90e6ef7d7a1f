// build_csr.cu -- K0/K1/K2: on-device validation of the COO triplets and their conversion to
// CSR (A) and a precomputed CSR of A' (so Atprod needs no atomics).
//
// The reference keeps the matrix as unsorted COO (src/lsqr.f90:113-118) and accumulates in COO
// order (:168-172, :188-192).  The conversion is a STABLE key sort (LSD radix sort of
// (key, COO position) pairs): inside a row / column the entries keep their COO order and
// duplicates are kept, so every per-row / per-column sum visits the same terms in the same order
// as the reference.  The result is bit-exact against oracle/csr_oracle.c (integer work + copies).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "build_csr.h"

namespace lsqrb {

namespace {

constexpr int kT = 256;

inline int grid_for(int64_t n, int per_sm = 8)
{
    int64_t blocks = (n + kT - 1) / kT;
    int64_t cap = (int64_t)kNumSMs * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// K0: min / max of an index array (bounds checks of src/lsqr.f90:110-111, plus the lower bound)
__global__ void __launch_bounds__(kT) minmax_kernel(int64_t nnz, const int32_t *__restrict__ a, int32_t *out /*[2]*/)
{
    int32_t lo = INT32_MAX, hi = INT32_MIN;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * kT) {
        const int32_t v = a[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out + 0, lo);
        atomicMax(out + 1, hi);
    }
}

// is key[] already non-decreasing?  (row-sorted COO needs no sort for A)
__global__ void __launch_bounds__(kT) unsorted_flag_kernel(int64_t nnz, const int32_t *__restrict__ key, int *flag)
{
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i + 1 < nnz; i += (int64_t)gridDim.x * kT)
        bad |= ((uint32_t)key[i] > (uint32_t)key[i + 1]);   // (composite keys of the blocked transpose use all 32 bits)
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// counts[key-1] += 1   (integer atomics: order-independent, hence deterministic)
__global__ void __launch_bounds__(kT) histogram_kernel(int64_t nnz, const int32_t *__restrict__ key, uint32_t *counts)
{
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * kT)
        atomicAdd(counts + ((uint32_t)key[i] - 1u), 1u);
}

// composite key of the row-blocked transpose: ((other-1)/block_rows)*nkeys + key   (1-based like key)
__global__ void __launch_bounds__(kT)
composite_key_kernel(int64_t nnz, const int32_t *__restrict__ key, const int32_t *__restrict__ other,
                     uint32_t block_rows, uint32_t nkeys, int32_t *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * kT)
        out[i] = (int32_t)(((uint32_t)(other[i] - 1) / block_rows) * nkeys + (uint32_t)key[i]);
}

__global__ void __launch_bounds__(kT) iota_kernel(int64_t nnz, uint32_t *p)
{
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * kT) p[i] = (uint32_t)i;
}

// idx[p] = other[perm[p]] - 1 ; val[p] = a[perm[p]]
__global__ void __launch_bounds__(kT)
gather_kernel(int64_t nnz, const uint32_t *__restrict__ perm, const int32_t *__restrict__ other,
              const double *__restrict__ a, int32_t *__restrict__ idx, double *__restrict__ val)
{
    for (int64_t p = (int64_t)blockIdx.x * kT + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * kT) {
        const uint32_t i = perm[p];
        idx[p] = other[i] - 1;
        val[p] = a[i];
    }
}

// identity permutation: idx = other - 1 ; val = a
__global__ void __launch_bounds__(kT)
copy_kernel(int64_t nnz, const int32_t *__restrict__ other, const double *__restrict__ a,
            int32_t *__restrict__ idx, double *__restrict__ val)
{
    for (int64_t p = (int64_t)blockIdx.x * kT + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * kT) {
        idx[p] = other[p] - 1;
        val[p] = a[p];
    }
}

// Temporaries of the build: freed on every exit path (after the stream has drained, so that no kernel still
// reads them); an early LSQRB_CUDA return therefore leaks nothing.
struct TempPool {
    cudaStream_t stream;
    void *ptrs[8];
    int n = 0;
    explicit TempPool(cudaStream_t s) : stream(s) {}
    template <typename T> cudaError_t alloc(T **out, size_t bytes)
    {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaSuccess) ptrs[n++] = p;
        *out = static_cast<T *>(p);
        return e;
    }
    ~TempPool()
    {
        if (n == 0) return;
        cudaStreamSynchronize(stream);
        for (int i = 0; i < n; ++i) cudaFree(ptrs[i]);
    }
};

int bits_for(int64_t nkeys)
{
    int b = 1;
    while (((int64_t)1 << b) <= nkeys) ++b;   // keys are 1..nkeys
    return b;
}

}  // namespace

int coo_validate(cudaStream_t stream, int32_t m, int32_t n, int64_t nnz,
                 const int32_t *d_irow, const int32_t *d_icol)
{
    if (nnz == 0) return LSQR_B200_OK;
    TempPool tmp(stream);
    int32_t *d_mm = nullptr;
    LSQRB_CUDA(tmp.alloc(&d_mm, 4 * sizeof(int32_t)));
    const int32_t init[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
    LSQRB_CUDA(cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, stream));
    minmax_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, d_irow, d_mm);
    minmax_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, d_icol, d_mm + 2);
    int32_t h[4];
    LSQRB_CUDA(cudaMemcpyAsync(h, d_mm, sizeof h, cudaMemcpyDeviceToHost, stream));
    LSQRB_CUDA(cudaStreamSynchronize(stream));
    // same order as the reference: irow first (:110), then icol (:111)
    if (h[1] > m) return LSQR_B200_ERR_IROW;
    if (h[3] > n) return LSQR_B200_ERR_ICOL;
    if (h[0] < 1 || h[2] < 1) {
        set_last_error("irow/icol contain an index < 1 (undefined behaviour in the reference)");
        return LSQR_B200_ERR_INDEX_LOW;
    }
    return LSQR_B200_OK;
}

int coo_to_csr_device(cudaStream_t stream, int64_t nkeys_one, int64_t nnz,
                      const int32_t *d_key_in, const int32_t *d_other, const double *d_a, Csr *out,
                      int64_t block_rows, int64_t nother)
{
    out->nkeys = nkeys_one;
    out->nblocks = 1;
    out->block_rows = 0;
    int64_t nkeys = nkeys_one;
    const int32_t *d_key = d_key_in;
    TempPool tmp(stream);
    int32_t *d_comp = nullptr;
    if (block_rows > 0 && nother > block_rows && nnz > 0) {
        const int64_t nb = (nother + block_rows - 1) / block_rows;
        if (nb * nkeys_one >= (int64_t)0xFFFFFFF0ll || block_rows >= (int64_t)0xFFFFFFFFll) {
            set_last_error("row-blocked transpose: blocks * columns exceeds 32 bits");
            return LSQR_B200_ERR_TOO_LARGE;
        }
        out->nblocks = nb;
        out->block_rows = block_rows;
        nkeys = nb * nkeys_one;
        LSQRB_CUDA(tmp.alloc(&d_comp, sizeof(int32_t) * (size_t)nnz));
        composite_key_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, d_key_in, d_other, (uint32_t)block_rows, (uint32_t)nkeys_one, d_comp);
        LSQRB_CUDA(cudaGetLastError());
        d_key = d_comp;
    }
    out->nrows = nkeys;
    out->nnz = nnz;
    out->was_sorted = 1;
    LSQRB_CUDA(cudaMalloc(&out->ptr, sizeof(uint32_t) * (size_t)(nkeys + 1 + 8)));   // + slack for aligned bulk reads
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    const size_t nzp = nz + 8;   // the tile-streamed SpMV reads 16-byte aligned supersets of a tile (spmv_stream.cuh)
    LSQRB_CUDA(cudaMalloc(&out->idx, sizeof(int32_t) * nzp));
    LSQRB_CUDA(cudaMalloc(&out->val, sizeof(double) * nzp));
    LSQRB_CUDA(cudaMemsetAsync(out->idx + nz - 1, 0, sizeof(int32_t) * 9, stream));
    LSQRB_CUDA(cudaMemsetAsync(out->val + nz - 1, 0, sizeof(double) * 9, stream));
    LSQRB_CUDA(cudaMalloc(&out->perm, sizeof(uint32_t) * nz));
    LSQRB_CUDA(cudaMemsetAsync(out->ptr, 0, sizeof(uint32_t) * (size_t)(nkeys + 1 + 8), stream));
    if (nnz == 0) return LSQR_B200_OK;

    // ptr: histogram of the keys, then an exclusive prefix sum
    histogram_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, d_key, out->ptr);
    {
        void *d_tmp = nullptr;
        size_t tmp_bytes = 0;
        LSQRB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, out->ptr, out->ptr, nkeys + 1, stream));
        LSQRB_CUDA(tmp.alloc(&d_tmp, tmp_bytes));
        LSQRB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, out->ptr, out->ptr, nkeys + 1, stream));
    }

    // already grouped by key?  then the stable sort is the identity
    int *d_flag = nullptr, h_flag = 0;
    LSQRB_CUDA(tmp.alloc(&d_flag, sizeof(int)));
    LSQRB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
    unsorted_flag_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, d_key, d_flag);
    LSQRB_CUDA(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    LSQRB_CUDA(cudaStreamSynchronize(stream));

    iota_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, out->perm);
    if (!h_flag) {
        copy_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, d_other, d_a, out->idx, out->val);
        LSQRB_CUDA(cudaGetLastError());
        return LSQR_B200_OK;
    }
    out->was_sorted = 0;

    // stable LSD radix sort of (key, COO position); only the low bits_for(nkeys) bits are sorted
    uint32_t *d_keys_out = nullptr, *d_iota = nullptr;
    LSQRB_CUDA(tmp.alloc(&d_keys_out, sizeof(uint32_t) * nz));
    LSQRB_CUDA(tmp.alloc(&d_iota, sizeof(uint32_t) * nz));
    LSQRB_CUDA(cudaMemcpyAsync(d_iota, out->perm, sizeof(uint32_t) * nz, cudaMemcpyDeviceToDevice, stream));
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    const uint32_t *keys_in = reinterpret_cast<const uint32_t *>(d_key);
    LSQRB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys_in, d_keys_out, d_iota, out->perm,
                                               nnz, 0, bits_for(nkeys), stream));
    LSQRB_CUDA(tmp.alloc(&d_tmp, tmp_bytes));
    LSQRB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys_in, d_keys_out, d_iota, out->perm,
                                               nnz, 0, bits_for(nkeys), stream));
    gather_kernel<<<grid_for(nnz), kT, 0, stream>>>(nnz, out->perm, d_other, d_a, out->idx, out->val);
    LSQRB_CUDA(cudaGetLastError());
    return LSQR_B200_OK;   // (the pool drains the stream and frees the temporaries)
}

void csr_free(Csr *c)
{
    if (!c) return;
    cudaFree(c->ptr);
    cudaFree(c->idx);
    cudaFree(c->val);
    cudaFree(c->perm);
    *c = Csr{};
}

}  // namespace lsqrb
