// lsqrtest_ez.cpp -- C++ counterpart of the reference's test/lsqrtest_ez.f90, call for call, against the
// B200 engine through the host mirror include/lsqr_b200.hpp (class lsqr_solver_ez).
//   test_1 (test/lsqrtest_ez.f90:18-52): 3x3 dense system as 9 COO triplets, itnlim = 100, nout = stdout
//   test_2 (test/lsqrtest_ez.f90:54-104): 3x4 under-determined system, 12 triplets
// Each asserts  any(abs(A x - b) > 1.0e-12) -> 'TEST FAILED'  like the reference (:50, :102) and, beyond the
// reference, compares istop / itn / x with the CPU oracle (tests may use the oracle as the checker).
#include <cmath>
#include <cstdio>
#include <vector>

#include "lsqr_b200.hpp"
#include "../../oracle/lsqr_oracle.h"

using namespace lsqr_module;

static int failures = 0;

static void run(const char *name, int m, int n, const std::vector<wp> &a, const std::vector<int32_t> &irow,
                const std::vector<int32_t> &icol, const std::vector<wp> &b)
{
    const wp damp = zero;
    lsqr_solver_ez solver;
    std::vector<wp> x((size_t)n);
    int istop = -1, itn = -1;
    solver.initialize(m, n, a, irow, icol, zero, zero, zero, /*itnlim=*/100, /*nout=*/stdout);
    solver.solve(b.data(), damp, x.data(), istop, nullptr, &itn);

    // A x - b with A given column-major like reshape(a,[m,n])
    std::vector<wp> r(b.size());
    wp worst = 0;
    for (int i = 0; i < m; ++i) {
        wp s = 0;
        for (int j = 0; j < n; ++j) s += a[(size_t)j * m + i] * x[(size_t)j];
        r[(size_t)i] = s - b[(size_t)i];
        worst = std::fmax(worst, std::fabs(r[(size_t)i]));
    }
    std::printf("\n %s: istop = %d  itn = %d\n x       =", name, istop, itn);
    for (wp v : x) std::printf("%16.6E", v);
    std::printf("\n A*x - b =");
    for (wp v : r) std::printf("%16.6E", v);
    std::printf("\n");
    if (worst > 1.0e-12) { std::printf("TEST FAILED (%s): max |A x - b| = %g\n", name, worst); ++failures; }

    // the oracle on the same input
    oracle_ez *o = nullptr;
    oracle_ez_opts oo{0.0, 0.0, 0.0, 100, 0};
    oracle_ez_initialize(&o, m, n, (int64_t)a.size(), a.data(), (int64_t)irow.size(), irow.data(), (int64_t)icol.size(), icol.data(), &oo);
    std::vector<wp> xo((size_t)n);
    int istop_o = -1, itn_o = -1;
    oracle_ez_solve(o, b.data(), damp, xo.data(), &istop_o, nullptr, &itn_o, nullptr, nullptr, nullptr, nullptr, nullptr,
                    nullptr, nullptr, nullptr, nullptr);
    oracle_ez_destroy(o);
    wp dx = 0, nx = 0;
    for (int j = 0; j < n; ++j) { dx += (x[(size_t)j] - xo[(size_t)j]) * (x[(size_t)j] - xo[(size_t)j]); nx += xo[(size_t)j] * xo[(size_t)j]; }
    const wp rel = std::sqrt(dx / nx);
    std::printf(" oracle: istop = %d itn = %d  rel diff in x = %.2e\n", istop_o, itn_o, rel);
    if (istop != istop_o || std::abs(itn - itn_o) > 2 || rel > 1e-10) { std::printf("PARITY FAILED (%s)\n", name); ++failures; }
}

int main()
{
    try {
        // test_1: README example, istop = 1, x = 1.242424E+00 -6.060606E-02 -4.040404E-02 (README.md:55-58)
        run("test_1", 3, 3, {1, 4, 7, 2, 5, 88, 3, 66, 9}, {1, 2, 3, 1, 2, 3, 1, 2, 3}, {1, 1, 1, 2, 2, 2, 3, 3, 3}, {1, 2, 3});
        // test_2 (n > m)
        run("test_2", 3, 4, {4.1, 1.1, 11.1, 5.1, -3.1, 3.1, 66.1, 8.1, -87.1, 0.1, -9.1, 2.1},
            {1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3}, {1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4}, {1, 2, 3});
        // the reference's error stops surface as lsqr_error with the same text
        try {
            lsqr_solver_ez bad;
            bad.initialize(2, 2, {1.0, 2.0}, {1, 3}, {1, 2});
            std::printf("ERROR STOP TEST FAILED: no error for irow > m\n");
            ++failures;
        } catch (const lsqr_error &e) {
            std::printf(" error stop text: '%s'\n", e.what());
            if (std::string(e.what()) != "invalid irow or m in initialize_ez") ++failures;
        }
    } catch (const std::exception &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "\nFAILED (%d)\n" : "\nALL EZ TESTS PASSED\n", failures);
    return failures ? 1 : 0;
}
