/*
 * lstp_oracle.c -- CPU ORACLE (test infrastructure; see lsqr_oracle.h).
 *
 * Restates the reference's LSTP test-problem family A = HY*D*HZ and its driver:
 *   test/lsqrtest_module.f90:119-272 (test), :283-309 (aprod dispatch),
 *   :319-343 (aprod1), :353-377 (aprod2), :385-403 (hprod), :422-505 (lstp).
 */
#include "lsqr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const double zero = 0.0, one = 1.0;

/* x**k for an integer variable k: gfortran lowers this to libgcc's __powidf2
 * (binary exponentiation), restated here so d(i) rounds the same way. */
static double powi(double x, int m)
{
    unsigned n = (m < 0) ? (unsigned)(-m) : (unsigned)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return (m < 0) ? 1.0 / y : y;
}

oracle_lstp *oracle_lstp_new(int m, int n)
{
    oracle_lstp *p = (oracle_lstp *)calloc(1, sizeof *p);
    p->m = m;
    p->n = n;
    p->maxmn = m > n ? m : n;
    p->minmn = m < n ? m : n;
    p->d = (double *)calloc((size_t)p->minmn + 1, sizeof(double));
    p->hy = (double *)calloc((size_t)m + 1, sizeof(double));
    p->hz = (double *)calloc((size_t)n + 1, sizeof(double));
    p->w = (double *)calloc((size_t)p->maxmn + 1, sizeof(double));
    return p;
}

void oracle_lstp_free(oracle_lstp *p)
{
    if (!p) return;
    free(p->d); free(p->hy); free(p->hz); free(p->w);
    free(p);
}

/* hprod, :385-403:  y = (I - 2 hz hz') x ; the dot product is accumulated as
 * s = hz(i)*x(i) + s.  x and y may alias (aprod1 calls hprod(m,hy,w,w)). */
void oracle_hprod(int n, const double *hz, const double *x, double *y)
{
    double s = zero;
    for (int i = 0; i < n; ++i) s = hz[i] * x[i] + s;
    s = s + s;
    for (int i = 0; i < n; ++i) y[i] = x[i] - s * hz[i];
}

/* aprod1, :319-343:  y = y + HY*D*HZ*x */
static void lstp_aprod1(oracle_lstp *p, double *x, double *y)
{
    const int m = p->m, n = p->n;
    double *w = p->w;
    oracle_hprod(n, p->hz, x, w);
    for (int i = 0; i < p->minmn; ++i) w[i] = p->d[i] * w[i];
    for (int i = n; i < m; ++i) w[i] = zero;
    oracle_hprod(m, p->hy, w, w);
    for (int i = 0; i < m; ++i) y[i] = y[i] + w[i];
}

/* aprod2, :353-377:  x = x + HZ*D*HY*y */
static void lstp_aprod2(oracle_lstp *p, double *x, double *y)
{
    const int m = p->m, n = p->n;
    double *w = p->w;
    oracle_hprod(m, p->hy, y, w);
    for (int i = 0; i < p->minmn; ++i) w[i] = p->d[i] * w[i];
    for (int i = m; i < n; ++i) w[i] = zero;
    oracle_hprod(n, p->hz, w, w);
    for (int i = 0; i < n; ++i) x[i] = x[i] + w[i];
}

/* aprod_test_solver, :283-309 */
void oracle_lstp_aprod(void *user, int mode, int m, int n, double *x, double *y)
{
    oracle_lstp *p = (oracle_lstp *)user;
    (void)m; (void)n;
    if (mode == 1) lstp_aprod1(p, x, y);
    else           lstp_aprod2(p, x, y);
}

/* lstp, :422-505 */
void oracle_lstp_generate(oracle_lstp *p, int nduplc, int npower, double damp, int fourpi_mode,
                          double *x, double *b, double *acond, double *rnorm)
{
    const int m = p->m, n = p->n, minmn = p->minmn;
    double *d = p->d, *hy = p->hy, *hz = p->hz, *w = p->w;

    /* :433.  The current source evaluates 4*acos(-1) in working precision; the committed
     * LSQR.LIS was produced with the single-precision product noted in the trailing comment. */
    double fourpi;
    if (fourpi_mode == ORACLE_FOURPI_F32) fourpi = (double)(4.0f * 3.141592f);
    else                                  fourpi = 4.0 * acos(-1.0);

    double dampsq = damp * damp;
    double alfa = fourpi / m;
    double beta = fourpi / n;

    for (int i = 1; i <= m; ++i) hy[i - 1] = sin(i * alfa); /* :441-447 */
    for (int i = 1; i <= n; ++i) hz[i - 1] = cos(i * beta);

    alfa = oracle_dnrm2(m, hy, 1); /* :449-452 */
    beta = oracle_dnrm2(n, hz, 1);
    oracle_dscal(m, -one / alfa, hy, 1);
    oracle_dscal(n, -one / beta, hz, 1);

    for (int i = 1; i <= minmn; ++i) { /* :457-462: singular values, nduplc copies of each */
        int j = (i - 1 + nduplc) / nduplc;
        double t = (double)(j * nduplc);
        t = t / minmn;
        d[i - 1] = powi(t, npower);
    }

    *acond = (d[minmn - 1] * d[minmn - 1] + dampsq) / (d[0] * d[0] + dampsq); /* :464-465 */
    *acond = sqrt(*acond);

    /* true solution of the form Z(w;0), :472-478 */
    oracle_hprod(n, hz, x, w);
    for (int i = m; i < n; ++i) w[i] = zero;
    oracle_hprod(n, hz, w, x);

    for (int i = 0; i < minmn; ++i) w[i] = dampsq * w[i] / d[i]; /* :483-485 */
    for (int i = minmn; i < m; ++i) w[i] = one;                  /* :490-492 */
    oracle_hprod(m, hy, w, w);                                   /* :494 */

    *rnorm = oracle_dnrm2(m, w, 1); /* :498-500:  b = r + A x */
    oracle_dcopy(m, w, 1, b, 1);
    lstp_aprod1(p, x, b);
}

static void emit(oracle_log_fn fn, void *user, const char *s) { if (fn) fn(user, s); }

/* subroutine test, :119-272 */
void oracle_lstp_test(int m, int n, int nduplc, int npower, double damp, int fourpi_mode,
                      oracle_log_fn log, void *log_user,
                      oracle_iter_fn iter_cb, void *iter_user,
                      oracle_lstp_result *res, double *x_out)
{
    const double eps = DBL_EPSILON; /* :127 */
    char line[256];
    oracle_lstp *p = oracle_lstp_new(m, n);
    const int maxmn = p->maxmn;
    double *b = (double *)calloc((size_t)m + 1, sizeof(double));
    double *u = (double *)calloc((size_t)m + 1, sizeof(double));
    double *v = (double *)calloc((size_t)n + 1, sizeof(double));
    double *w = (double *)calloc((size_t)maxmn + 1, sizeof(double));
    double *x = (double *)calloc((size_t)n + 1, sizeof(double));
    double *se = (double *)calloc((size_t)n + 1, sizeof(double));
    double *xtrue = (double *)calloc((size_t)n + 1, sizeof(double));
    double *y = (double *)calloc((size_t)maxmn + 1, sizeof(double));

    for (int j = 1; j <= n; ++j) xtrue[j - 1] = j * 0.1; /* :151-154 */

    double acond, rnorm;
    oracle_lstp_generate(p, nduplc, npower, damp, fourpi_mode, xtrue, b, &acond, &rnorm); /* :173-175 */
    res->gen_acond = acond;
    res->gen_rnorm = rnorm;

    if (log) { /* format 1000, :243-248 */
        emit(log, log_user, ""); emit(log, log_user, "");
        emit(log, log_user, " --------------------------------------------------------------------");
        snprintf(line, sizeof line, " Least-Squares Test Problem      P(%5d%5d%5d%5d%12.2E )", m, n, nduplc, npower, damp);
        emit(log, log_user, line);
        emit(log, log_user, "");
        snprintf(line, sizeof line, " Condition no. =%12.4E     Residual function =%17.9E", acond, rnorm);
        emit(log, log_user, line);
        emit(log, log_user, " --------------------------------------------------------------------");
    }

    /* :183-188 */
    oracle_acheck(oracle_lstp_aprod, p, m, n, log, log_user, eps, v, w, x, y,
                  &res->acheck_inform, &res->acheck_relerr);

    if (res->acheck_inform > 0) {
        emit(log, log_user, " Check eps and power in subroutine acheck");
        goto done; /* the reference executes `stop` here */
    }

    { /* :195-206 */
        oracle_dcopy(m, b, 1, u, 1);
        int wantse = 0;
        double atol = pow(eps, 0.99);
        double btol = atol;
        double conlim = 1000.0 * acond;
        int itnlim = 4 * (m + n + 50);

        oracle_lsqr(oracle_lstp_aprod, p, m, n, damp, wantse, u, v, w, x, se,
                    atol, btol, conlim, itnlim, log, log_user, iter_cb, iter_user,
                    &res->istop, &res->itn, &res->anorm, &res->acond, &res->rnorm,
                    &res->arnorm, &res->xnorm);

        /* :216-218 */
        oracle_xcheck(oracle_lstp_aprod, p, m, n, log, log_user, res->anorm, damp, eps,
                      b, u, v, w, x, &res->xcheck_inform, &res->xtest1, &res->xtest2,
                      &res->xtest3, res->xnorms);

        int nprint = m < n ? m : n; /* :222-226 */
        if (nprint > 8) nprint = 8;
        for (int j = 0; j < 8; ++j) res->x_head[j] = (j < nprint) ? x[j] : 0.0;
        if (log) {
            emit(log, log_user, ""); emit(log, log_user, "");
            emit(log, log_user, " Solution  x:");
            for (int j0 = 0; j0 < nprint; j0 += 4) {
                int pos = 0;
                for (int j = j0; j < nprint && j < j0 + 4; ++j)
                    pos += snprintf(line + pos, sizeof line - (size_t)pos, "%6d%14.6G", j + 1, x[j]);
                emit(log, log_user, line);
            }
        }

        /* :230-241 */
        for (int j = 0; j < n; ++j) w[j] = x[j] - xtrue[j];
        double wnorm = oracle_dnrm2(n, w, 1);
        double xn = oracle_dnrm2(n, xtrue, 1);
        double enorm = wnorm / (one + xn);
        double etol = 0.001;
        res->enorm = enorm;
        if (log) {
            emit(log, log_user, "");
            if (enorm <= etol)
                snprintf(line, sizeof line, " LSQR  appears to be successful.     Relative error in  x  =%10.2E", enorm);
            else
                snprintf(line, sizeof line, " LSQR  appears to have failed.       Relative error in  x  =%10.2E", enorm);
            emit(log, log_user, line);
        }
        if (x_out) memcpy(x_out, x, sizeof(double) * (size_t)n);
    }

done:
    free(b); free(u); free(v); free(w); free(x); free(se); free(xtrue); free(y);
    oracle_lstp_free(p);
}
