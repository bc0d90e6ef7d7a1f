// build_csr.h -- device COO -> CSR / CSR' build (K0-K2), see build_csr.cu
#pragma once
#include "common.cuh"

namespace lsqrb {

struct Csr {
    uint32_t *ptr = nullptr;    // [nrows+1]   row starts (0-based)
    int32_t  *idx = nullptr;    // [nnz]       0-based column (A) or row (A') index
    double   *val = nullptr;    // [nnz]
    uint32_t *perm = nullptr;   // [nnz]       0-based COO position of each stored entry
    int64_t   nrows = 0, nnz = 0;
    int       was_sorted = 0;   // the COO keys were already non-decreasing (no sort needed)
};

// Bounds checks of initialize_ez (src/lsqr.f90:110-111) on device arrays.
int coo_validate(cudaStream_t stream, int32_t m, int32_t n, int64_t nnz,
                 const int32_t *d_irow, const int32_t *d_icol);

// Stable sort by `key` (1-based, values 1..nkeys); `other` is the other coordinate (1-based).
int coo_to_csr_device(cudaStream_t stream, int64_t nkeys, int64_t nnz,
                      const int32_t *d_key, const int32_t *d_other, const double *d_a, Csr *out);

void csr_free(Csr *c);

}  // namespace lsqrb
