/*
 * csr_oracle.c -- CPU ORACLE (test infrastructure; see lsqr_oracle.h).
 *
 * Host reference for the COO -> CSR / CSR-transpose conversion.  The reference
 * (src/lsqr.f90:113-118) keeps the matrix as unsorted COO for its whole life and
 * walks it in COO order (:168-172, :188-192); it has no conversion code, so this
 * stable counting sort DEFINES the bit-exact target of the device build: within a
 * row (column) entries stay in COO order and duplicates are kept, which preserves
 * the reference's per-row (per-column) accumulation order.
 */
#include "lsqr_oracle.h"

#include <stdlib.h>
#include <string.h>

void oracle_coo_to_csr(int64_t nkeys, int64_t nnz,
                       const int32_t *irow, const int32_t *icol, const double *a,
                       int by_col,
                       int64_t *ptr, int32_t *idx, double *val, int64_t *perm)
{
    const int32_t *key = by_col ? icol : irow;
    const int32_t *oth = by_col ? irow : icol;

    for (int64_t k = 0; k <= nkeys; ++k) ptr[k] = 0;
    for (int64_t i = 0; i < nnz; ++i) ptr[(int64_t)key[i]] += 1;   /* key is 1-based: counts land in ptr[1..] */
    for (int64_t k = 0; k < nkeys; ++k) ptr[k + 1] += ptr[k];

    int64_t *next = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nkeys > 0 ? nkeys : 1));
    memcpy(next, ptr, sizeof(int64_t) * (size_t)nkeys);
    for (int64_t i = 0; i < nnz; ++i) {
        int64_t k = (int64_t)key[i] - 1;
        int64_t p = next[k]++;
        idx[p] = oth[i] - 1;
        val[p] = a[i];
        perm[p] = i;
    }
    free(next);
}
