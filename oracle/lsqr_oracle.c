/*
 * lsqr_oracle.c -- CPU ORACLE (test infrastructure; see lsqr_oracle.h).
 *
 * Restates, operation by operation, the solver side of the reference:
 *   src/lsqrblas.f90  (dcopy, ddot, dnrm2, dscal)
 *   src/lsqr.f90      (initialize_ez, aprod_ez, solve_ez, LSQR, acheck, xcheck, d2norm)
 * Every function cites the reference lines it follows.  Floating-point
 * expressions keep the reference's association order; build with
 * -ffp-contract=off so the compiler cannot fuse a*b+c.
 */
#include "lsqr_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const double zero = 0.0, one = 1.0; /* src/lsqr_kinds.F90:27-28 */

/* ------------------------------------------------------------------------ */
/* Fortran-style formatted output helpers (for the nout log)                 */
/* ------------------------------------------------------------------------ */

/* 1PEw.d : one digit before the point, d after, exponent E+XX (or +XXX). */
static void fe(char *out, int w, int d, double v)
{
    char tmp[64];
    snprintf(tmp, sizeof tmp, "%.*E", d, v);
    char *e = strchr(tmp, 'E');
    if (e && strlen(e + 2) >= 3) {        /* three exponent digits: Fortran drops the 'E' */
        memmove(e, e + 1, strlen(e + 1) + 1);
    }
    int len = (int)strlen(tmp);
    if (len > w) {                        /* field overflow prints asterisks */
        memset(out, '*', (size_t)w);
        out[w] = 0;
        return;
    }
    memset(out, ' ', (size_t)(w - len));
    memcpy(out + (w - len), tmp, (size_t)len + 1);
}

typedef struct { oracle_log_fn fn; void *user; } logger;

static void emit(const logger *lg, const char *line)
{
    if (lg->fn) lg->fn(lg->user, line);
}

/* ------------------------------------------------------------------------ */
/* src/lsqrblas.f90                                                          */
/* ------------------------------------------------------------------------ */

/* dcopy, src/lsqrblas.f90:25-67.  Copies x to y; unit strides take the
 * clean-up-then-blocks-of-7 route, anything else walks both index streams. */
void oracle_dcopy(int n, const double *dx, int incx, double *dy, int incy)
{
    if (n <= 0) return;
    if (incx == 1 && incy == 1) {
        int m = n % 7;
        for (int i = 0; i < m; ++i) dy[i] = dx[i];
        if (n < 7) return;
        for (int i = m; i < n; i += 7)
            for (int k = 0; k < 7; ++k) dy[i + k] = dx[i + k];
    } else {
        long ix = 0, iy = 0;
        if (incx < 0) ix = (long)(-n + 1) * incx;
        if (incy < 0) iy = (long)(-n + 1) * incy;
        for (int i = 0; i < n; ++i) {
            dy[iy] = dx[ix];
            ix += incx;
            iy += incy;
        }
    }
}

/* ddot, src/lsqrblas.f90:74-116.  The unit-stride path adds five products to the
 * running sum in one left-to-right expression per block (:103). */
double oracle_ddot(int n, const double *dx, int incx, const double *dy, int incy)
{
    double dtemp = zero;
    if (n <= 0) return zero;
    if (incx == 1 && incy == 1) {
        int m = n % 5;
        for (int i = 0; i < m; ++i) dtemp = dtemp + dx[i] * dy[i];
        if (n < 5) return dtemp;
        for (int i = m; i < n; i += 5)
            dtemp = dtemp + dx[i] * dy[i] + dx[i + 1] * dy[i + 1] + dx[i + 2] * dy[i + 2]
                          + dx[i + 3] * dy[i + 3] + dx[i + 4] * dy[i + 4];
    } else {
        long ix = 0, iy = 0;
        if (incx < 0) ix = (long)(-n + 1) * incx;
        if (incy < 0) iy = (long)(-n + 1) * incy;
        for (int i = 0; i < n; ++i) {
            dtemp = dtemp + dx[ix] * dy[iy];
            ix += incx;
            iy += incy;
        }
    }
    return dtemp;
}

/* dnrm2, src/lsqrblas.f90:123-159.  Scaled sum of squares (the old dlassq loop):
 * zeros are skipped, a new maximum rescales ssq, one division per element. */
double oracle_dnrm2(int n, const double *x, int incx)
{
    if (n < 1 || incx < 1) return zero;
    if (n == 1) return fabs(x[0]);
    double scale = zero, ssq = one;
    for (long ix = 0; ix <= (long)(n - 1) * incx; ix += incx) {
        if (x[ix] != zero) {
            double absxi = fabs(x[ix]);
            if (scale < absxi) {
                double q = scale / absxi;
                ssq = one + ssq * (q * q);
                scale = absxi;
            } else {
                double q = absxi / scale;
                ssq = ssq + q * q;
            }
        }
    }
    return scale * sqrt(ssq);
}

/* dscal, src/lsqrblas.f90:166-201. */
void oracle_dscal(int n, double da, double *dx, int incx)
{
    if (n <= 0 || incx <= 0) return;
    if (incx == 1) {
        int m = n % 5;
        for (int i = 0; i < m; ++i) dx[i] = da * dx[i];
        if (n < 5) return;
        for (int i = m; i < n; i += 5)
            for (int k = 0; k < 5; ++k) dx[i + k] = da * dx[i + k];
    } else {
        long nincx = (long)n * incx;
        for (long i = 0; i < nincx; i += incx) dx[i] = da * dx[i];
    }
}

/* d2norm, src/lsqr.f90:1164-1179: sqrt(a^2+b^2) scaled by |a|+|b|. */
double oracle_d2norm(double a, double b)
{
    double scale = fabs(a) + fabs(b);
    if (scale == zero) return zero;
    double p = a / scale, q = b / scale;
    return scale * sqrt(p * p + q * q);
}

/* ------------------------------------------------------------------------ */
/* LSQR, src/lsqr.f90:432-882                                                */
/* ------------------------------------------------------------------------ */

static const char *const lsqr_msg[6] = { /* :581-586 */
    "The exact solution is x = 0                          ",
    "A solution to Ax = b was found, given atol, btol     ",
    "A least-squares solution was found, given atol       ",
    "A damped least-squares solution was found, given atol",
    "Cond(Abar) seems to be too large, given conlim       ",
    "The iteration limit was reached                      "};

static void log_iter_line(const logger *lg, int itn, int nvals, const double *vals)
{
    /* '(1P, I6, 2E17.9, 4E10.2, E9.1, 3E8.1)', :671,828-829 */
    static const int w[10] = {17, 17, 10, 10, 10, 10, 9, 8, 8, 8};
    static const int d[10] = {9, 9, 2, 2, 2, 2, 1, 1, 1, 1};
    char line[256], f[32];
    int pos = snprintf(line, sizeof line, "%6d", itn);
    for (int k = 0; k < nvals; ++k) {
        fe(f, w[k], d[k], vals[k]);
        pos += snprintf(line + pos, sizeof line - (size_t)pos, "%s", f);
    }
    emit(lg, line);
}

void oracle_lsqr(oracle_aprod_fn aprod, void *aprod_user,
                 int m, int n, double damp, int wantse,
                 double *u, double *v, double *w, double *x, double *se,
                 double atol, double btol, double conlim, int itnlim,
                 oracle_log_fn log, void *log_user,
                 oracle_iter_fn iter_cb, void *iter_user,
                 int *istop_out, int *itn_out, double *anorm_out, double *acond_out,
                 double *rnorm_out, double *arnorm_out, double *xnorm_out)
{
    const logger lg = {log, log_user};
    const int nout = (log != NULL);
    char line[256], f1[40], f2[40];

    int damped, i, maxdx, nconv, nstop, istop, itn;
    double alfopt, alpha, beta, bnorm = zero, cs, cs1, cs2, ctol, delta, dknorm, dnorm, dxk, dxmax,
           gamma, gambar, phi, phibar = zero, psi, res2, rho, rhobar = zero, rhbar1, rhs, rtol, sn,
           sn1, sn2, t, tau, temp, test1, test2, test3, theta, t1, t2, t3, xnorm1, z, zbar;
    double anorm, acond, rnorm = zero, arnorm, xnorm;

    if (nout) { /* :588-595 */
        emit(&lg, ""); emit(&lg, "");
        emit(&lg, " Enter LSQR.       Least-squares solution of  Ax = b");
        snprintf(line, sizeof line, " The matrix  A  has%7d rows   and%7d columns", m, n);
        emit(&lg, line);
        fe(f1, 22, 14, damp);
        snprintf(line, sizeof line, " damp   =%s   wantse =%10s", f1, wantse ? "T" : "F");
        emit(&lg, line);
        fe(f1, 10, 2, atol); fe(f2, 10, 2, conlim);
        snprintf(line, sizeof line, " atol   =%s               conlim =%s", f1, f2);
        emit(&lg, line);
        fe(f1, 10, 2, btol);
        snprintf(line, sizeof line, " btol   =%s               itnlim =%10d", f1, itnlim);
        emit(&lg, line);
    }

    /* :597-617 */
    damped = damp > zero;
    itn = 0;
    istop = 0;
    nstop = 0;
    maxdx = 0;
    if (conlim > zero) ctol = one / conlim; else ctol = zero;
    anorm = zero;
    acond = zero;
    dnorm = zero;
    dxmax = zero;
    res2 = zero;
    psi = zero;
    xnorm = zero;
    xnorm1 = zero;
    cs2 = -one;
    sn2 = zero;
    z = zero;

    /* :621-630  first vectors of the bidiagonalization */
    for (i = 0; i < n; ++i) { v[i] = zero; x[i] = zero; }
    if (wantse) for (i = 0; i < n; ++i) se[i] = zero;

    /* :632-644 */
    alpha = zero;
    beta = oracle_dnrm2(m, u, 1);
    if (beta > zero) {
        oracle_dscal(m, one / beta, u, 1);
        aprod(aprod_user, 2, m, n, v, u);
        alpha = oracle_dnrm2(n, v, 1);
    }
    if (alpha > zero) {
        oracle_dscal(n, one / alpha, v, 1);
        oracle_dcopy(n, v, 1, w, 1);
    }

    arnorm = alpha * beta; /* :646 */

    if (arnorm != zero) {
        rhobar = alpha; /* :650-653 */
        phibar = beta;
        bnorm = beta;
        rnorm = beta;

        if (nout) { /* :655-672 */
            emit(&lg, ""); emit(&lg, "");
            if (damped)
                emit(&lg, "   Itn       x(1)           Function     Compatible   LS     Norm Abar Cond Abar");
            else
                emit(&lg, "   Itn       x(1)           Function     Compatible   LS        Norm A    Cond A");
            test1 = one;
            test2 = alpha / beta;
            snprintf(line, sizeof line, "%80s%s", "", "    phi    dknorm   dxk  alfa_opt");
            emit(&lg, line);
            double vals[4] = {x[0], rnorm, test1, test2};
            log_iter_line(&lg, itn, 4, vals);
            emit(&lg, "");
        }

        for (;;) { /* main iteration loop, :673-852 */
            itn = itn + 1;

            /* bidiagonalization step, :681-699 */
            oracle_dscal(m, -alpha, u, 1);
            aprod(aprod_user, 1, m, n, v, u);
            beta = oracle_dnrm2(m, u, 1);

            /* anorm accumulates sqrt(sum alpha^2+beta^2+damp^2), :687-689 */
            temp = oracle_d2norm(alpha, beta);
            temp = oracle_d2norm(temp, damp);
            anorm = oracle_d2norm(anorm, temp);

            if (beta > zero) {
                oracle_dscal(m, one / beta, u, 1);
                oracle_dscal(n, -beta, v, 1);
                aprod(aprod_user, 2, m, n, v, u);
                alpha = oracle_dnrm2(n, v, 1);
                if (alpha > zero) oracle_dscal(n, one / alpha, v, 1);
            }

            /* rotation that removes damp, :703-710 */
            rhbar1 = rhobar;
            if (damped) {
                rhbar1 = oracle_d2norm(rhobar, damp);
                cs1 = rhobar / rhbar1;
                sn1 = damp / rhbar1;
                psi = sn1 * phibar;
                phibar = cs1 * phibar;
            }

            /* rotation that removes the subdiagonal beta, :714-721 */
            rho = oracle_d2norm(rhbar1, beta);
            cs = rhbar1 / rho;
            sn = beta / rho;
            theta = sn * alpha;
            rhobar = -cs * alpha;
            phi = cs * phibar;
            phibar = sn * phibar;
            tau = sn * phi;

            /* x, w (and se) update, :724-745 */
            t1 = phi / rho;
            t2 = -theta / rho;
            t3 = one / rho;
            dknorm = zero;
            if (wantse) {
                for (i = 0; i < n; ++i) {
                    t = w[i];
                    x[i] = t1 * t + x[i];
                    w[i] = t2 * t + v[i];
                    t = (t3 * t) * (t3 * t);
                    se[i] = t + se[i];
                    dknorm = t + dknorm;
                }
            } else {
                for (i = 0; i < n; ++i) {
                    t = w[i];
                    x[i] = t1 * t + x[i];
                    w[i] = t2 * t + v[i];
                    dknorm = (t3 * t) * (t3 * t) + dknorm;
                }
            }

            /* norms of the update, :751-757 */
            dknorm = sqrt(dknorm);
            dnorm = oracle_d2norm(dnorm, dknorm);
            dxk = fabs(phi * dknorm);
            if (dxmax < dxk) {
                dxmax = dxk;
                maxdx = itn;
            }

            /* right rotation, estimate of norm(x), :762-771 */
            delta = sn2 * rho;
            gambar = -cs2 * rho;
            rhs = phi - delta * z;
            zbar = rhs / gambar;
            xnorm = oracle_d2norm(xnorm1, zbar);
            gamma = oracle_d2norm(gambar, theta);
            cs2 = gambar / gamma;
            sn2 = theta / gamma;
            z = rhs / gamma;
            xnorm1 = oracle_d2norm(xnorm1, z);

            /* estimates, :776-790 */
            acond = anorm * dnorm;
            res2 = oracle_d2norm(res2, psi);
            rnorm = oracle_d2norm(res2, phibar);
            arnorm = alpha * fabs(tau);

            alfopt = sqrt(rnorm / (dnorm * xnorm));
            test1 = rnorm / bnorm;
            test2 = zero;
            if (rnorm > zero) test2 = arnorm / (anorm * rnorm);
            test3 = one / acond;
            t1 = test1 / (one + anorm * xnorm / bnorm);
            rtol = btol + atol * anorm * xnorm / bnorm;

            /* stopping tests, :798-810 (later assignments win) */
            t3 = one + test3;
            t2 = one + test2;
            t1 = one + t1;
            if (itn >= itnlim) istop = 5;
            if (t3 <= one) istop = 4;
            if (t2 <= one) istop = 2;
            if (t1 <= one) istop = 1;

            if (test3 <= ctol) istop = 4;
            if (test2 <= atol) istop = 2;
            if (test1 <= rtol) istop = 1;

            if (iter_cb) {
                oracle_iter_rec rec;
                rec.itn = itn; rec.x1 = x[0]; rec.rnorm = rnorm; rec.test1 = test1;
                rec.test2 = test2; rec.anorm = anorm; rec.acond = acond; rec.phi = phi;
                rec.dknorm = dknorm; rec.dxk = dxk; rec.alfopt = alfopt;
                rec.alpha = alpha; rec.beta = beta; rec.xnorm = xnorm; rec.arnorm = arnorm;
                iter_cb(iter_user, &rec);
            }

            if (nout) { /* :813-837 */
                int print_iter = (n <= 40) || (itn <= 10) || (itn >= itnlim - 10) ||
                                 (itn % 10 == 0) || (test3 <= 2.0 * ctol) ||
                                 (test2 <= 10.0 * atol) || (test1 <= 10.0 * rtol) || (istop != 0);
                if (print_iter) {
                    double vals[10] = {x[0], rnorm, test1, test2, anorm, acond, phi, dknorm, dxk, alfopt};
                    log_iter_line(&lg, itn, 10, vals);
                }
            }

            /* nconv gate, :843-850 */
            if (istop == 0) {
                nstop = 0;
            } else {
                nconv = 1;
                nstop = nstop + 1;
                if (nstop < nconv && itn < itnlim) istop = 0;
            }
            if (istop != 0) break;
        }

        /* standard errors, :857-865 */
        if (wantse) {
            t = one;
            if (m > n) t = (double)(m - n);
            if (damped) t = (double)m;
            t = rnorm / sqrt(t);
            for (i = 0; i < n; ++i) se[i] = t * sqrt(se[i]);
        }
    }

    if (damped && istop == 2) istop = 3; /* :871 */
    if (nout) { /* :872-880 */
        emit(&lg, ""); emit(&lg, "");
        snprintf(line, sizeof line, " Exit  LSQR.       istop  =%2d               itn    =%8d", istop, itn);
        emit(&lg, line);
        fe(f1, 12, 5, anorm); fe(f2, 12, 5, acond);
        snprintf(line, sizeof line, " Exit  LSQR.       anorm  =%s     acond  =%s", f1, f2);
        emit(&lg, line);
        fe(f1, 12, 5, bnorm); fe(f2, 12, 5, xnorm);
        snprintf(line, sizeof line, " Exit  LSQR.       bnorm  =%s     xnorm  =%s", f1, f2);
        emit(&lg, line);
        fe(f1, 12, 5, rnorm); fe(f2, 12, 5, arnorm);
        snprintf(line, sizeof line, " Exit  LSQR.       rnorm  =%s     arnorm =%s", f1, f2);
        emit(&lg, line);
        fe(f1, 8, 1, dxmax);
        snprintf(line, sizeof line, " Exit  LSQR.       max dx =%s occurred at itn %8d", f1, maxdx);
        emit(&lg, line);
        fe(f1, 8, 1, dxmax / (xnorm + 1.0e-20));
        snprintf(line, sizeof line, " Exit  LSQR.              =%s*xnorm", f1);
        emit(&lg, line);
        snprintf(line, sizeof line, " Exit  LSQR.       %s", lsqr_msg[istop]);
        emit(&lg, line);
    }

    *istop_out = istop;
    *itn_out = itn;
    *anorm_out = anorm;
    *acond_out = acond;
    *rnorm_out = rnorm;   /* NOTE: the reference leaves rnorm unassigned when alpha*beta = 0 (:646-653);
                             the oracle reports 0.0 there. */
    *arnorm_out = arnorm;
    *xnorm_out = xnorm;
}

/* ------------------------------------------------------------------------ */
/* acheck, src/lsqr.f90:908-994                                              */
/* ------------------------------------------------------------------------ */
void oracle_acheck(oracle_aprod_fn aprod, void *aprod_user, int m, int n,
                   oracle_log_fn log, void *log_user, double eps,
                   double *v, double *w, double *x, double *y,
                   int *inform, double *relerr)
{
    const logger lg = {log, log_user};
    char line[128], f1[32];
    const double power = 0.5; /* :927 */
    double tol = pow(eps, power);
    if (log) { emit(&lg, ""); emit(&lg, ""); emit(&lg, "Enter acheck. Test of aprod for LSQR and CRAIG"); }

    /* "unlikely" unit vectors, :946-961 */
    double t = one;
    for (int j = 0; j < n; ++j) { t = t + one; x[j] = sqrt(t); }
    t = one;
    for (int i = 0; i < m; ++i) { t = t + one; y[i] = one / sqrt(t); }

    double alfa = oracle_dnrm2(n, x, 1);
    double beta = oracle_dnrm2(m, y, 1);
    oracle_dscal(n, one / alfa, x, 1);
    oracle_dscal(m, one / beta, y, 1);

    /* w = y + A x,  v = x + A'y, :969-972 */
    oracle_dcopy(m, y, 1, w, 1);
    oracle_dcopy(n, x, 1, v, 1);
    aprod(aprod_user, 1, m, n, x, w);
    aprod(aprod_user, 2, m, n, v, y);

    alfa = oracle_ddot(m, y, 1, w, 1); /* :976-980 */
    beta = oracle_ddot(n, x, 1, v, 1);
    double test1 = fabs(alfa - beta);
    double test2 = one + fabs(alfa) + fabs(beta);
    double test3 = test1 / test2;

    if (test3 <= tol) { /* :984-992 */
        *inform = 0;
        if (log) { fe(f1, 10, 1, test3); snprintf(line, sizeof line, "aprod seems OK. Relative error = %s", f1); emit(&lg, line); }
    } else {
        *inform = 1;
        if (log) { fe(f1, 10, 1, test3); snprintf(line, sizeof line, "aprod seems incorrect. Relative error = %s", f1); emit(&lg, line); }
    }
    if (relerr) *relerr = test3;
}

/* ------------------------------------------------------------------------ */
/* xcheck, src/lsqr.f90:1015-1154                                            */
/* ------------------------------------------------------------------------ */
void oracle_xcheck(oracle_aprod_fn aprod, void *aprod_user, int m, int n,
                   oracle_log_fn log, void *log_user,
                   double anorm, double damp, double eps,
                   const double *b, double *u, double *v, double *w, const double *x,
                   int *inform, double *test1, double *test2, double *test3, double *norms)
{
    const logger lg = {log, log_user};
    char line[160], f1[40];
    const double power = 0.5;
    double dampsq = damp * damp;
    double tol = pow(eps, power);
    double *xtmp = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    memcpy(xtmp, x, sizeof(double) * (size_t)n); /* :1064 */

    /* u = b - A x via u = -b + A x, u = -u, :1069-1076 */
    oracle_dcopy(m, b, 1, u, 1);
    oracle_dscal(m, -one, u, 1);
    aprod(aprod_user, 1, m, n, xtmp, u);
    oracle_dscal(m, -one, u, 1);

    /* v = A'u, :1080-1083 */
    for (int j = 0; j < n; ++j) v[j] = zero;
    aprod(aprod_user, 2, m, n, v, u);

    /* w = A'u - damp^2 x, :1089-1094 */
    oracle_dcopy(n, v, 1, w, 1);
    if (damp != zero)
        for (int j = 0; j < n; ++j) w[j] = w[j] - dampsq * x[j];

    double bnorm = oracle_dnrm2(m, b, 1); /* :1098-1101 */
    double xnorm = oracle_dnrm2(n, x, 1);
    double rho1 = oracle_dnrm2(m, u, 1);
    double sigma1 = oracle_dnrm2(n, v, 1);
    double rho2, sigma2;
    if (log) {
        emit(&lg, ""); emit(&lg, "");
        emit(&lg, "Enter xcheck. Does x solve Ax = b, etc?");
        fe(f1, 10, 3, damp);   snprintf(line, sizeof line, " damp            =%s", f1); emit(&lg, line);
        fe(f1, 10, 3, xnorm);  snprintf(line, sizeof line, " norm(x)         =%s", f1); emit(&lg, line);
        fe(f1, 15, 8, rho1);   snprintf(line, sizeof line, " norm(r)         =%s = rho1", f1); emit(&lg, line);
        fe(f1, 10, 3, sigma1); snprintf(line, sizeof line, " norm(A'r)       =%s      = sigma1", f1); emit(&lg, line);
    }

    if (damp == zero) { /* :1110-1124 */
        rho2 = rho1;
        sigma2 = sigma1;
    } else {
        rho2 = sqrt(rho1 * rho1 + dampsq * (xnorm * xnorm));
        sigma2 = oracle_dnrm2(n, w, 1);
        double snorm = rho1 / damp;
        double xsnorm = rho2 / damp;
        if (log) {
            emit(&lg, "");
            fe(f1, 10, 3, snorm);  snprintf(line, sizeof line, " norm(s)         =%s", f1); emit(&lg, line);
            fe(f1, 10, 3, xsnorm); snprintf(line, sizeof line, " norm(x,s)       =%s", f1); emit(&lg, line);
            fe(f1, 15, 8, rho2);   snprintf(line, sizeof line, " norm(rbar)      =%s = rho2", f1); emit(&lg, line);
            fe(f1, 10, 3, sigma2); snprintf(line, sizeof line, " norm(Abar'rbar) =%s      = sigma2", f1); emit(&lg, line);
        }
    }

    if (bnorm == zero && xnorm == zero) { /* :1129-1144 */
        *inform = 0;
        *test1 = zero;
        *test2 = zero;
        *test3 = zero;
    } else {
        *inform = 4;
        *test1 = rho1 / (bnorm + anorm * xnorm);
        *test2 = zero;
        if (rho1 > zero) *test2 = sigma1 / (anorm * rho1);
        *test3 = *test2;
        if (rho2 > zero) *test3 = sigma2 / (anorm * rho2);
        if (*test3 <= tol) *inform = 3;
        if (*test2 <= tol) *inform = 2;
        if (*test1 <= tol) *inform = 1;
    }

    if (log) { /* :1146-1152 */
        emit(&lg, "");
        snprintf(line, sizeof line, " inform          =%2d", *inform); emit(&lg, line);
        fe(f1, 10, 3, tol);    snprintf(line, sizeof line, " tol             =%s", f1); emit(&lg, line);
        fe(f1, 10, 3, *test1); snprintf(line, sizeof line, " test1           =%s (Ax = b)", f1); emit(&lg, line);
        fe(f1, 10, 3, *test2); snprintf(line, sizeof line, " test2           =%s (least-squares)", f1); emit(&lg, line);
        fe(f1, 10, 3, *test3); snprintf(line, sizeof line, " test3           =%s (damped least-squares)", f1); emit(&lg, line);
    }
    if (norms) {
        norms[0] = bnorm; norms[1] = xnorm; norms[2] = rho1;
        norms[3] = sigma1; norms[4] = rho2; norms[5] = sigma2;
    }
    free(xtmp);
}

/* ------------------------------------------------------------------------ */
/* class lsqr_solver_ez, src/lsqr.f90:32-65,91-259                           */
/* ------------------------------------------------------------------------ */
struct oracle_ez {
    int      m, n;
    int64_t  num_nonzero_elements;
    int32_t *irow, *icol;
    double  *a;
    double   atol, btol, conlim;
    int      itnlim;
    double  *Ax, *Aty, *v, *w;   /* allocated lazily, persist across solves (:166,186,239-240) */
    int      last_error;
};

const char *oracle_error_message(int code)
{
    switch (code) {
    case ORACLE_OK:         return "";
    case ORACLE_ERR_SIZES:  return "invalid a,icol,irow sizes in initialize_ez";
    case ORACLE_ERR_IROW:   return "invalid irow or m in initialize_ez";
    case ORACLE_ERR_ICOL:   return "invalid icol or n in initialize_ez";
    case ORACLE_ERR_NOINIT: return "lsqr_solver_ez class not properly initialized";
    case ORACLE_ERR_MODE:   return "invalid mode in aprod_ez";
    default:                return "unknown";
    }
}

int oracle_ez_initialize(oracle_ez **out, int m, int n,
                         int64_t size_a, const double *a,
                         int64_t size_irow, const int32_t *irow,
                         int64_t size_icol, const int32_t *icol,
                         const oracle_ez_opts *opts)
{
    *out = NULL;
    /* :109-111 -- only upper bounds are checked, like the reference */
    if (size_a != size_irow || size_a != size_icol) return ORACLE_ERR_SIZES;
    for (int64_t k = 0; k < size_irow; ++k) if (irow[k] > m) return ORACLE_ERR_IROW;
    for (int64_t k = 0; k < size_icol; ++k) if (icol[k] > n) return ORACLE_ERR_ICOL;

    oracle_ez *me = (oracle_ez *)calloc(1, sizeof *me);
    me->num_nonzero_elements = size_irow; /* :113-118: deep copies */
    me->m = m;
    me->n = n;
    size_t nz = (size_t)(size_a > 0 ? size_a : 1);
    me->irow = (int32_t *)malloc(nz * sizeof(int32_t));
    me->icol = (int32_t *)malloc(nz * sizeof(int32_t));
    me->a = (double *)malloc(nz * sizeof(double));
    memcpy(me->irow, irow, (size_t)size_a * sizeof(int32_t));
    memcpy(me->icol, icol, (size_t)size_a * sizeof(int32_t));
    memcpy(me->a, a, (size_t)size_a * sizeof(double));
    /* defaults :46-51, optionals :121-125 */
    me->atol = zero; me->btol = zero; me->conlim = zero; me->itnlim = 100;
    if (opts) {
        me->atol = opts->atol; me->btol = opts->btol; me->conlim = opts->conlim;
        me->itnlim = opts->itnlim;
    }
    *out = me;
    return ORACLE_OK;
}

void oracle_ez_destroy(oracle_ez *me)
{
    if (!me) return;
    free(me->irow); free(me->icol); free(me->a);
    free(me->Ax); free(me->Aty); free(me->v); free(me->w);
    free(me);
}

/* aprod_ez, src/lsqr.f90:134-200.  The product is accumulated in COO order into a
 * zeroed workspace and then added to the in/out vector. */
int oracle_ez_aprod(oracle_ez *me, int mode, int m, int n, double *x, double *y)
{
    if (m != me->m || n != me->n) return ORACLE_ERR_NOINIT; /* :152 */
    const int64_t nnz = me->num_nonzero_elements;
    switch (mode) {
    case 1: /* y = y + A*x, :156-174 */
        if (!me->Ax) me->Ax = (double *)malloc(sizeof(double) * (size_t)(me->m > 0 ? me->m : 1));
        for (int i = 0; i < me->m; ++i) me->Ax[i] = zero;
        for (int64_t i = 0; i < nnz; ++i) {
            int r = me->irow[i] - 1;
            int c = me->icol[i] - 1;
            me->Ax[r] = me->Ax[r] + me->a[i] * x[c];
        }
        for (int i = 0; i < me->m; ++i) y[i] = y[i] + me->Ax[i];
        return ORACLE_OK;
    case 2: /* x = x + A'*y, :176-194 */
        if (!me->Aty) me->Aty = (double *)malloc(sizeof(double) * (size_t)(me->n > 0 ? me->n : 1));
        for (int i = 0; i < me->n; ++i) me->Aty[i] = zero;
        for (int64_t i = 0; i < nnz; ++i) {
            int r = me->irow[i] - 1;
            int c = me->icol[i] - 1;
            me->Aty[c] = me->Aty[c] + me->a[i] * y[r];
        }
        for (int i = 0; i < me->n; ++i) x[i] = x[i] + me->Aty[i];
        return ORACLE_OK;
    default:
        return ORACLE_ERR_MODE; /* :197 */
    }
}

static void ez_aprod_thunk(void *user, int mode, int m, int n, double *x, double *y)
{
    oracle_ez *me = (oracle_ez *)user;
    int rc = oracle_ez_aprod(me, mode, m, n, x, y);
    if (rc != ORACLE_OK) me->last_error = rc;
}

/* solve_ez, src/lsqr.f90:207-259 */
void oracle_ez_solve(oracle_ez *me, const double *b, double damp, double *x, int *istop,
                     double *se, int *itn, double *anorm, double *acond,
                     double *rnorm, double *arnorm, double *xnorm,
                     oracle_log_fn log, void *log_user,
                     oracle_iter_fn iter_cb, void *iter_user)
{
    int wantse = (se != NULL); /* :232-237 */
    double *se_ = (double *)malloc(sizeof(double) * (size_t)(wantse ? (me->n > 0 ? me->n : 1) : 1));
    double *u = (double *)malloc(sizeof(double) * (size_t)(me->m > 0 ? me->m : 1));
    if (!me->v) me->v = (double *)malloc(sizeof(double) * (size_t)(me->n > 0 ? me->n : 1));
    if (!me->w) me->w = (double *)malloc(sizeof(double) * (size_t)(me->n > 0 ? me->n : 1));
    memcpy(u, b, sizeof(double) * (size_t)me->m); /* :242 */

    int itn_;
    double anorm_, acond_, rnorm_, arnorm_, xnorm_;
    oracle_lsqr(ez_aprod_thunk, me, me->m, me->n, damp, wantse, u, me->v, me->w, x, se_,
                me->atol, me->btol, me->conlim, me->itnlim, log, log_user, iter_cb, iter_user,
                istop, &itn_, &anorm_, &acond_, &rnorm_, &arnorm_, &xnorm_);

    if (wantse) memcpy(se, se_, sizeof(double) * (size_t)me->n); /* :251-257 */
    if (itn) *itn = itn_;
    if (anorm) *anorm = anorm_;
    if (acond) *acond = acond_;
    if (rnorm) *rnorm = rnorm_;
    if (arnorm) *arnorm = arnorm_;
    if (xnorm) *xnorm = xnorm_;
    free(u);
    free(se_);
}
