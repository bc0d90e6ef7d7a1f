// lsqr_b200.hpp -- C++ host-side mirror of the reference's public interface (module lsqr_module,
// src/lsqr.f90:16-82) over the C ABI in lsqr_b200.h.  Header only.
//
// The reference is compiled Fortran 2008 and this build image has no Fortran compiler, so the host
// layer that is actually compiled and tested here is this C++ one (the Fortran shim that binds the
// same C ABI through iso_c_binding lives in fortran/ and INTEGRATION.md).  Names, argument order,
// argument meaning and error behaviour follow the reference:
//
//   reference (Fortran 2008)                                 here (namespace lsqr_module)
//   -------------------------------------------------------  ---------------------------------------------
//   type(lsqr_solver_ez) :: s                                lsqr_solver_ez s;
//   call s%initialize(m,n,a,irow,icol,atol,btol,conlim,      s.initialize(m,n,a,irow,icol,atol,btol,conlim,
//                     itnlim,nout)                :91                     itnlim,nout);
//   call s%solve(b,damp,x,istop,se,itn,anorm,acond,          s.solve(b,damp,x,istop,&se,&itn,&anorm,&acond,
//                rnorm,arnorm,xnorm)              :207                   &rnorm,&arnorm,&xnorm);
//   call s%aprod(mode,m,n,x,y)                    :134       s.aprod(mode,m,n,x,y);        (host vectors)
//   type,extends(lsqr_solver) :: my ; aprod => .. :16-30     struct my : lsqr_solver { void aprod(...) override; };
//   call my%lsqr(m,n,damp,wantse,u,v,w,x,se,atol,btol,       my.lsqr(m,n,damp,wantse,u,v,w,x,se,atol,btol,
//                conlim,itnlim,nout,istop,itn,anorm,                 conlim,itnlim,nout,istop,itn,anorm,
//                acond,rnorm,arnorm,xnorm)        :432               acond,rnorm,arnorm,xnorm);  (DEVICE vectors)
//   call my%acheck(m,n,nout,eps,v,w,x,y,inform)   :908       my.acheck(m,n,nout,eps,v,w,x,y,inform);
//   call my%xcheck(m,n,nout,anorm,damp,eps,b,u,v,w,x,        my.xcheck(m,n,nout,anorm,damp,eps,b,u,v,w,x,
//                  inform,test1,test2,test3)      :1015                inform,test1,test2,test3);
//
// `nout` is a std::FILE* (nullptr = the reference's nout = 0: silent).  Where the reference executes
// `error stop '<message>'` this layer throws lsqr_error carrying the same message; set
// lsqr_module::error_stop_aborts = true to get the reference's behaviour (message to stderr, abort).
// Index arrays are 1-based 32-bit integers like the reference's.  In the low-level class the vectors
// are DEVICE arrays and aprod receives device pointers plus the CUDA stream to enqueue on.
#ifndef LSQR_B200_HPP
#define LSQR_B200_HPP

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "lsqr_b200.h"

namespace lsqr_module {

// lsqr_kinds (src/lsqr_kinds.F90:16-28)
using wp = double;
constexpr wp zero = 0.0;
constexpr wp one = 1.0;

struct lsqr_error : std::runtime_error {
    int code;
    lsqr_error(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

inline bool error_stop_aborts = false;

inline void check(int rc)
{
    if (rc == LSQR_B200_OK) return;
    std::string msg = lsqr_b200_error_message(rc);
    const char *detail = lsqr_b200_last_error();
    if (rc > 5 && detail && *detail) msg += std::string(" [") + detail + "]";
    if (error_stop_aborts) {   // `error stop '<message>'`
        std::fprintf(stderr, "ERROR STOP %s\n", msg.c_str());
        std::abort();
    }
    throw lsqr_error(rc, msg);
}

namespace detail {
inline void log_to_file(void *user, const char *line) { std::fprintf(static_cast<std::FILE *>(user), "%s\n", line); }
inline lsqr_b200_options options_with_nout(std::FILE *nout, void *stream)
{
    lsqr_b200_options o;
    lsqr_b200_default_options(&o);
    o.stream = stream;
    if (nout) { o.log = &log_to_file; o.log_user = nout; }
    return o;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------
// type,abstract :: lsqr_solver  (src/lsqr.f90:16-30); aprod_func (:67-82)
// ------------------------------------------------------------------------------------------------
class lsqr_solver {
public:
    virtual ~lsqr_solver() = default;

    // deferred procedure aprod (:26).  mode 1: y(m) = y + A*x(n); mode 2: x(n) = x + A'*y(m).
    // x and y are DEVICE pointers; enqueue the work on `stream` (a cudaStream_t) and do not synchronize.
    virtual void aprod(int mode, int m, int n, wp *x, wp *y, void *stream) = 0;

    // LSQR (:432-882).  u(m) [in: b, destroyed], v(n), w(n), x(n) [out], se(n) [only if wantse]: DEVICE arrays.
    void lsqr(int m, int n, wp damp, bool wantse, wp *u, wp *v, wp *w, wp *x, wp *se,
              wp atol, wp btol, wp conlim, int itnlim, std::FILE *nout,
              int &istop, int &itn, wp &anorm, wp &acond, wp &rnorm, wp &arnorm, wp &xnorm, void *stream = nullptr)
    {
        lsqr_b200_options o = detail::options_with_nout(nout, stream);
        int32_t is = 0, it = 0;
        check(lsqr_b200_lsqr(&lsqr_solver::trampoline, this, m, n, damp, wantse ? 1 : 0, u, v, w, x, se,
                             atol, btol, conlim, itnlim, &o, &is, &it, &anorm, &acond, &rnorm, &arnorm, &xnorm));
        istop = is;
        itn = it;
    }

    // acheck (:908-994): v(n), w(m), x(n), y(m) are DEVICE work arrays; inform = 0 if aprod looks consistent.
    void acheck(int m, int n, std::FILE *nout, wp eps, wp *v, wp *w, wp *x, wp *y, int &inform, void *stream = nullptr)
    {
        lsqr_b200_options o = detail::options_with_nout(nout, stream);
        int32_t inf = 0;
        check(lsqr_b200_acheck(&lsqr_solver::trampoline, this, m, n, eps, v, w, x, y, &o, &inf, nullptr));
        inform = inf;
    }

    // xcheck (:1015-1154): b(m), x(n) DEVICE inputs; u(m), v(n), w(n) DEVICE outputs (r, A'r, A'r - damp^2 x).
    void xcheck(int m, int n, std::FILE *nout, wp anorm, wp damp, wp eps, const wp *b, wp *u, wp *v, wp *w, const wp *x,
                int &inform, wp &test1, wp &test2, wp &test3, void *stream = nullptr)
    {
        lsqr_b200_options o = detail::options_with_nout(nout, stream);
        int32_t inf = 0;
        check(lsqr_b200_xcheck(&lsqr_solver::trampoline, this, m, n, anorm, damp, eps, b, u, v, w, x, &o,
                               &inf, &test1, &test2, &test3, nullptr));
        inform = inf;
    }

private:
    static int trampoline(void *user, int32_t mode, int32_t m, int32_t n, double *x, double *y, void *stream)
    {
        try {
            static_cast<lsqr_solver *>(user)->aprod(mode, m, n, x, y, stream);
            return 0;
        } catch (const lsqr_error &e) {
            return e.code ? e.code : 1;
        } catch (...) {
            return 1;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// The same abstract class with the reference's OWN signatures: every vector is a HOST array and aprod is host code
// (aprod_func, src/lsqr.f90:67-82).  Each product makes a host round trip; all vector arithmetic and the scalar
// recurrence run on the GPU.  What an unmodified type,extends(lsqr_solver) of the reference maps to.
// ------------------------------------------------------------------------------------------------
class lsqr_solver_host {
public:
    virtual ~lsqr_solver_host() = default;
    virtual void aprod(int mode, int m, int n, wp *x, wp *y) = 0;          // host arrays x(n), y(m), both inout

    void lsqr(int m, int n, wp damp, bool wantse, wp *u, wp *v, wp *w, wp *x, wp *se,
              wp atol, wp btol, wp conlim, int itnlim, std::FILE *nout,
              int &istop, int &itn, wp &anorm, wp &acond, wp &rnorm, wp &arnorm, wp &xnorm)
    {
        lsqr_b200_options o = detail::options_with_nout(nout, nullptr);
        int32_t is = 0, it = 0;
        check(lsqr_b200_lsqr_host(&lsqr_solver_host::trampoline, this, m, n, damp, wantse ? 1 : 0, u, v, w, x, se,
                                  atol, btol, conlim, itnlim, &o, &is, &it, &anorm, &acond, &rnorm, &arnorm, &xnorm));
        istop = is;
        itn = it;
    }
    void acheck(int m, int n, std::FILE *nout, wp eps, wp *v, wp *w, wp *x, wp *y, int &inform)
    {
        lsqr_b200_options o = detail::options_with_nout(nout, nullptr);
        int32_t inf = 0;
        check(lsqr_b200_acheck_host(&lsqr_solver_host::trampoline, this, m, n, eps, v, w, x, y, &o, &inf, nullptr));
        inform = inf;
    }
    void xcheck(int m, int n, std::FILE *nout, wp anorm, wp damp, wp eps, const wp *b, wp *u, wp *v, wp *w, const wp *x,
                int &inform, wp &test1, wp &test2, wp &test3)
    {
        lsqr_b200_options o = detail::options_with_nout(nout, nullptr);
        int32_t inf = 0;
        check(lsqr_b200_xcheck_host(&lsqr_solver_host::trampoline, this, m, n, anorm, damp, eps, b, u, v, w, x, &o,
                                    &inf, &test1, &test2, &test3, nullptr));
        inform = inf;
    }

private:
    static int trampoline(void *user, int32_t mode, int32_t m, int32_t n, double *x, double *y)
    {
        try {
            static_cast<lsqr_solver_host *>(user)->aprod(mode, m, n, x, y);
            return 0;
        } catch (...) {
            return 1;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// type,extends(lsqr_solver) :: lsqr_solver_ez  (src/lsqr.f90:32-65)
// ------------------------------------------------------------------------------------------------
class lsqr_solver_ez : public lsqr_solver {
public:
    lsqr_solver_ez() = default;
    lsqr_solver_ez(const lsqr_solver_ez &) = delete;
    lsqr_solver_ez &operator=(const lsqr_solver_ez &) = delete;
    ~lsqr_solver_ez() override { lsqr_b200_ez_destroy(h_); }

    // initialize_ez (:91-127).  a, irow, icol may be host or device arrays; they are copied (:113-118).
    void initialize(int m, int n, const std::vector<wp> &a, const std::vector<int32_t> &irow, const std::vector<int32_t> &icol,
                    wp atol = zero, wp btol = zero, wp conlim = zero, int itnlim = 100, std::FILE *nout = nullptr)
    {
        initialize(m, n, (int64_t)a.size(), a.data(), (int64_t)irow.size(), irow.data(), (int64_t)icol.size(), icol.data(),
                   atol, btol, conlim, itnlim, nout);
    }
    void initialize(int m, int n, int64_t size_a, const wp *a, int64_t size_irow, const int32_t *irow,
                    int64_t size_icol, const int32_t *icol,
                    wp atol = zero, wp btol = zero, wp conlim = zero, int itnlim = 100, std::FILE *nout = nullptr)
    {
        lsqr_b200_ez_destroy(h_);   // `me` is intent(out): the object is reset (:95)
        h_ = nullptr;
        m_ = n_ = 0;
        lsqr_b200_options o = detail::options_with_nout(nout, nullptr);
        o.atol = atol; o.btol = btol; o.conlim = conlim; o.itnlim = itnlim;
        check(lsqr_b200_ez_initialize(&h_, m, n, size_a, a, size_irow, irow, size_icol, icol, &o));
        m_ = m;
        n_ = n;
    }

    // solve_ez (:207-259).  b(m), x(n), se(n): host or device arrays.  Optional outputs may be nullptr.
    void solve(const wp *b, wp damp, wp *x, int &istop, wp *se = nullptr, int *itn = nullptr, wp *anorm = nullptr,
               wp *acond = nullptr, wp *rnorm = nullptr, wp *arnorm = nullptr, wp *xnorm = nullptr)
    {
        if (!h_) check(LSQR_B200_ERR_NOINIT);
        int32_t is = 0, it = 0;
        check(lsqr_b200_ez_solve(h_, b, damp, x, &is, se, &it, anorm, acond, rnorm, arnorm, xnorm));
        istop = is;
        if (itn) *itn = it;
    }
    void solve(const std::vector<wp> &b, wp damp, std::vector<wp> &x, int &istop)
    {
        x.assign((size_t)n_, zero);
        solve(b.data(), damp, x.data(), istop);
    }

    // aprod_ez (:134-200) on host or device vectors (blocking), as a user of the ez class would call it
    void aprod(int mode, int m, int n, wp *x, wp *y)
    {
        if (!h_) check(LSQR_B200_ERR_NOINIT);
        check(lsqr_b200_ez_aprod(h_, mode, m, n, x, y));
    }
    // ... and as the operator of the low-level solver (device pointers, stream ordered)
    void aprod(int mode, int m, int n, wp *x, wp *y, void *stream) override
    {
        if (!h_) check(LSQR_B200_ERR_NOINIT);
        check(lsqr_b200_ez_aprod_device(h_, mode, m, n, x, y, stream));
    }

    lsqr_b200_ez *handle() const { return h_; }

private:
    lsqr_b200_ez *h_ = nullptr;
    int m_ = 0, n_ = 0;
};

}  // namespace lsqr_module

// lsqpblas_module [sic] (src/lsqrblas.f90:8,16): dcopy, ddot, dnrm2, dscal on DEVICE or HOST arrays, stride 1
namespace lsqpblas_module {
inline void dcopy(int n, const double *dx, double *dy, void *stream = nullptr) { lsqr_module::check(lsqr_b200_dcopy(n, dx, dy, stream)); }
inline double ddot(int n, const double *dx, const double *dy, void *stream = nullptr)
{
    double r = 0.0;
    lsqr_module::check(lsqr_b200_ddot(n, dx, dy, &r, stream));
    return r;
}
inline double dnrm2(int n, const double *dx, void *stream = nullptr)
{
    double r = 0.0;
    lsqr_module::check(lsqr_b200_dnrm2(n, dx, &r, stream));
    return r;
}
inline void dscal(int n, double da, double *dx, void *stream = nullptr) { lsqr_module::check(lsqr_b200_dscal(n, da, dx, stream)); }
}  // namespace lsqpblas_module

#endif  // LSQR_B200_HPP
