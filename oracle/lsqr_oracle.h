/*
 * lsqr_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference algorithm of jacobwilliams/LSQR
 * (src/lsqr.f90, src/lsqrblas.f90, test/lsqrtest_module.f90).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the reported CPU baseline.
 * The shipped GPU path (lsqr_b200/) never links or calls it.
 *
 * Parity status: PINNED against the reference's own known answers
 *   - README.md:55-58 and test/lsqrtest_ez.f90:18-52,54-104 (ez KATs),
 *   - test/LSQR.LIS (18 LSTP problems: generator outputs, early iteration rows,
 *     istop) with the single-precision fourpi constant that log was made with.
 * UNPINNED (no golden data exists in the reference): se() values (wantse) and the
 * COO->CSR conversion (the reference never converts; csr_oracle.c *defines* it).
 *
 * The reference itself is Fortran 2008 and cannot be compiled in this image
 * (no gfortran/flang/nvfortran), so there is no oracle/_ref build.
 *
 * All index arrays follow the reference: 1-based, 32-bit.
 * Compile with -ffp-contract=off (see Makefile) so no FMA contraction changes
 * the rounding sequence of the restated loops.
 */
#ifndef LSQR_ORACLE_H
#define LSQR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/lsqrblas.f90 ---------------------------------------------------- */
void   oracle_dcopy(int n, const double *dx, int incx, double *dy, int incy); /* :25-67   */
double oracle_ddot (int n, const double *dx, int incx, const double *dy, int incy); /* :74-116 */
double oracle_dnrm2(int n, const double *x, int incx);                        /* :123-159 */
void   oracle_dscal(int n, double da, double *dx, int incx);                  /* :166-201 */

/* ---- src/lsqr.f90:1164-1179 ---------------------------------------------- */
double oracle_d2norm(double a, double b);

/* ---- the operator hook, src/lsqr.f90:67-82 -------------------------------
 * mode 1: y = y + A*x   (x has n entries, y has m entries)
 * mode 2: x = x + A'*y
 */
typedef void (*oracle_aprod_fn)(void *user, int mode, int m, int n, double *x, double *y);

/* Receives each line the reference would write to unit nout (without newline). */
typedef void (*oracle_log_fn)(void *user, const char *line);

/* Per-iteration scalar trace (optional; for parity tests of the device recurrence). */
typedef struct {
    int    itn;
    double x1, rnorm, test1, test2, anorm, acond, phi, dknorm, dxk, alfopt;
    double alpha, beta, xnorm, arnorm;
} oracle_iter_rec;
typedef void (*oracle_iter_fn)(void *user, const oracle_iter_rec *rec);

/* ---- src/lsqr.f90:432-882 (LSQR) ----------------------------------------- */
void oracle_lsqr(oracle_aprod_fn aprod, void *aprod_user,
                 int m, int n, double damp, int wantse,
                 double *u, double *v, double *w, double *x, double *se,
                 double atol, double btol, double conlim, int itnlim,
                 oracle_log_fn log, void *log_user,
                 oracle_iter_fn iter_cb, void *iter_user,
                 int *istop, int *itn, double *anorm, double *acond,
                 double *rnorm, double *arnorm, double *xnorm);

/* ---- src/lsqr.f90:908-994 (acheck) --------------------------------------- */
void oracle_acheck(oracle_aprod_fn aprod, void *aprod_user, int m, int n,
                   oracle_log_fn log, void *log_user, double eps,
                   double *v, double *w, double *x, double *y,
                   int *inform, double *relerr);

/* ---- src/lsqr.f90:1015-1154 (xcheck) ------------------------------------- */
void oracle_xcheck(oracle_aprod_fn aprod, void *aprod_user, int m, int n,
                   oracle_log_fn log, void *log_user,
                   double anorm, double damp, double eps,
                   const double *b, double *u, double *v, double *w, const double *x,
                   int *inform, double *test1, double *test2, double *test3,
                   double *norms /* [6]: bnorm,xnorm,rho1,sigma1,rho2,sigma2 or NULL */);

/* ---- class lsqr_solver_ez, src/lsqr.f90:32-65 ----------------------------- */
typedef struct oracle_ez oracle_ez;

enum {
    ORACLE_OK = 0,
    ORACLE_ERR_SIZES = 1,   /* 'invalid a,icol,irow sizes in initialize_ez'       :109 */
    ORACLE_ERR_IROW  = 2,   /* 'invalid irow or m in initialize_ez'               :110 */
    ORACLE_ERR_ICOL  = 3,   /* 'invalid icol or n in initialize_ez'               :111 */
    ORACLE_ERR_NOINIT= 4,   /* 'lsqr_solver_ez class not properly initialized'    :152 */
    ORACLE_ERR_MODE  = 5    /* 'invalid mode in aprod_ez'                         :197 */
};
const char *oracle_error_message(int code);

/* initialize_ez, src/lsqr.f90:91-127.  size_a/size_irow/size_icol are the three array
 * lengths the Fortran code compares; opts may be NULL for the defaults :46-51.
 * Returns ORACLE_OK or the code of the `error stop` the reference would raise. */
typedef struct {
    double atol, btol, conlim;
    int    itnlim;
    int    has_log;       /* nout /= 0 */
} oracle_ez_opts;
int  oracle_ez_initialize(oracle_ez **out, int m, int n,
                          int64_t size_a, const double *a,
                          int64_t size_irow, const int32_t *irow,
                          int64_t size_icol, const int32_t *icol,
                          const oracle_ez_opts *opts);
void oracle_ez_destroy(oracle_ez *me);
/* aprod_ez, src/lsqr.f90:134-200 (returns an error code instead of error stop) */
int  oracle_ez_aprod(oracle_ez *me, int mode, int m, int n, double *x, double *y);
/* solve_ez, src/lsqr.f90:207-259; se and the scalar outputs may be NULL (optional). */
void oracle_ez_solve(oracle_ez *me, const double *b, double damp, double *x, int *istop,
                     double *se, int *itn, double *anorm, double *acond,
                     double *rnorm, double *arnorm, double *xnorm,
                     oracle_log_fn log, void *log_user,
                     oracle_iter_fn iter_cb, void *iter_user);

/* ---- LSTP test problems, test/lsqrtest_module.f90 ------------------------- */
typedef struct {
    int     m, n, maxmn, minmn;
    double *d, *hy, *hz, *w;    /* rw(locd..), :162-168 */
} oracle_lstp;

#define ORACLE_FOURPI_F64  0   /* 4*acos(-1) in wp: current source, lsqrtest_module.f90:433 */
#define ORACLE_FOURPI_F32  1   /* real32 4.0*3.141592 = 12.566368103027344: what LSQR.LIS used */

oracle_lstp *oracle_lstp_new(int m, int n);
void oracle_lstp_free(oracle_lstp *p);
/* lstp, :422-505.  x (n) is in/out, b (m) out. */
void oracle_lstp_generate(oracle_lstp *p, int nduplc, int npower, double damp, int fourpi_mode,
                          double *x, double *b, double *acond, double *rnorm);
/* aprod_test_solver, :283-309 (an oracle_aprod_fn; user = oracle_lstp*) */
void oracle_lstp_aprod(void *user, int mode, int m, int n, double *x, double *y);
void oracle_hprod(int n, const double *hz, const double *x, double *y); /* :385-403 */

/* subroutine test, :119-272: one LSTP problem end-to-end.  Outputs are optional. */
typedef struct {
    double gen_acond, gen_rnorm;      /* header: Condition no., Residual function */
    int    acheck_inform; double acheck_relerr;
    int    istop, itn;
    double anorm, acond, rnorm, arnorm, xnorm;
    int    xcheck_inform; double xtest1, xtest2, xtest3;
    double xnorms[6];
    double enorm;                     /* ||x-xtrue||/(1+||xtrue||) */
    double x_head[8];
} oracle_lstp_result;
void oracle_lstp_test(int m, int n, int nduplc, int npower, double damp, int fourpi_mode,
                      oracle_log_fn log, void *log_user,
                      oracle_iter_fn iter_cb, void *iter_user,
                      oracle_lstp_result *res, double *x_out /* n or NULL */);

/* ---- host reference for the COO -> CSR conversion (defines K1/K2 parity) ---
 * Stable counting sort of the triplets by row (by_col=0) or by column (by_col=1).
 * Inputs are the reference's 1-based COO arrays.  Outputs are 0-based:
 *   ptr[nkeys+1], idx[nnz] (the other coordinate, 0-based), val[nnz],
 *   perm[nnz] (0-based position of each output entry in the COO input).
 * Inside one key the COO order is kept and duplicates are not merged, which is
 * exactly the accumulation order of src/lsqr.f90:168-172 / :188-192 per row/column.
 */
void oracle_coo_to_csr(int64_t nkeys, int64_t nnz,
                       const int32_t *irow, const int32_t *icol, const double *a,
                       int by_col,
                       int64_t *ptr, int32_t *idx, double *val, int64_t *perm);

#ifdef __cplusplus
}
#endif
#endif
