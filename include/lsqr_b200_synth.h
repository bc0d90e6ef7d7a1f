/*
 * lsqr_b200_synth.h -- device-side generators of the BASELINE.json synthetic workloads (K9).
 *
 * No reference counterpart: jacobwilliams/LSQR ships no benchmark inputs.  These entry points
 * produce, directly in HBM, exactly the triplets / vectors that lsqr_b200/synth.py produces on
 * the host (a pure counter hash of (seed, global row, slot)), so that the full-size
 * configurations (up to 2e9 stored entries) never have to cross PCIe and any row block can be
 * generated independently by the rank that owns it.  All pointers are DEVICE pointers unless
 * stated otherwise; `stream` is a cudaStream_t (NULL = default stream); calls are synchronous.
 */
#ifndef LSQR_B200_SYNTH_H
#define LSQR_B200_SYNTH_H

#include "lsqr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { LSQR_B200_SYNTH_UNIFORM = 0, LSQR_B200_SYNTH_BANDED = 1, LSQR_B200_SYNTH_POWERLAW = 2 };

/* Row pointer of rows [row0, row0 + nrows) of the workload: ptr_dev[0..nrows] (int64, 0-based,
 * ptr_dev[0] = 0).  k = entries per row (uniform / banded).  For the power-law kind the row
 * length is #{ j : u <= table[j] } with `table_host` the HOST array of ntable decreasing
 * thresholds (synth.powerlaw_table()).  *nnz_out (host) receives ptr_dev[nrows]. */
LSQR_B200_API int lsqr_b200_synth_row_ptr(int32_t kind, uint64_t seed, int64_t row0, int64_t nrows, int32_t k,
                                          const double *table_host, int32_t ntable,
                                          int64_t *ptr_dev, int64_t *nnz_out, void *stream);

/* Fills the COO triplets of the block: irow (1-based inside the block), icol (1-based, global),
 * a.  m, n are the GLOBAL dimensions; ptr_dev comes from lsqr_b200_synth_row_ptr. */
LSQR_B200_API int lsqr_b200_synth_fill(int32_t kind, uint64_t seed, int64_t m, int64_t n, int64_t row0, int64_t nrows,
                                       const int64_t *ptr_dev, int64_t nnz,
                                       int32_t *irow_dev, int32_t *icol_dev, double *a_dev, void *stream);

/* out_dev[i] = coef * (2 u(seed, tag, offset + i) - 1), i = 0..count-1  (synth.vector). */
LSQR_B200_API int lsqr_b200_synth_vector(uint64_t seed, int32_t tag, double coef, int64_t offset, int64_t count,
                                         double *out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LSQR_B200_SYNTH_H */
