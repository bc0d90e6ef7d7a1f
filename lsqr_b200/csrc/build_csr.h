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
    // Row-blocked transpose: the entries are grouped by blocks of `block_rows` values of the OTHER coordinate
    // first; block b is an ordinary CSR over `nkeys` keys whose pointer array is ptr + b*nkeys.  nrows = nblocks*nkeys.
    int64_t   nkeys = 0, nblocks = 1, block_rows = 0;
};

// Bounds checks of initialize_ez (src/lsqr.f90:110-111) on device arrays.
int coo_validate(cudaStream_t stream, int32_t m, int32_t n, int64_t nnz,
                 const int32_t *d_irow, const int32_t *d_icol);

// Stable sort by `key` (1-based, values 1..nkeys); `other` is the other coordinate (1-based).
// block_rows > 0 (with nother = number of values of `other`): stable sort by ((other-1)/block_rows, key) instead,
// i.e. one CSR per block of `block_rows` consecutive values of the other coordinate, stored back to back.
int coo_to_csr_device(cudaStream_t stream, int64_t nkeys, int64_t nnz,
                      const int32_t *d_key, const int32_t *d_other, const double *d_a, Csr *out,
                      int64_t block_rows = 0, int64_t nother = 0);

void csr_free(Csr *c);

}  // namespace lsqrb
