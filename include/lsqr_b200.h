/*
 * lsqr_b200.h -- C ABI of the B200-native LSQR engine (the drop-in boundary).
 *
 * Every entry point replaces one procedure of jacobwilliams/LSQR's hot path
 * (src/lsqr.f90); the reference line range is cited at each declaration.  The ABI is
 * plain C (pointers + sizes, no C++/torch types) so that the reference's Fortran 2008
 * host layer can bind it with iso_c_binding (see INTEGRATION.md and fortran/).
 *
 * Conventions
 *  - All index arrays are the reference's: 32-bit, 1-BASED (Fortran default integer).
 *  - All floating point is IEEE binary64 (wp = real64, src/lsqr_kinds.F90:23).
 *  - Vector / matrix pointers may be HOST or DEVICE pointers unless stated otherwise;
 *    the library classifies them with cudaPointerGetAttributes.  Host data is staged
 *    through pinned buffers; device data is used in place / copied device-to-device.
 *  - Every function returns LSQR_B200_OK (0) or an error code.  Codes 1..5 stand for the
 *    reference's `error stop '<message>'` statements and lsqr_b200_error_message() returns
 *    the reference's literal message so a Fortran shim can `error stop` with it.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails
 *    with LSQR_B200_ERR_NO_DEVICE.
 *  - One handle = one solve at a time (the reference object is not re-entrant either:
 *    aprod_ez mutates me%Ax / me%Aty, src/lsqr.f90:166-167,186-187).  Different handles
 *    may be driven from different host threads; the library keeps no global mutable state
 *    (the error detail of lsqr_b200_last_error() is per thread).
 *  - Stream ordering.  options.stream == NULL: the handle works on a library-owned BLOCKING
 *    stream, which CUDA orders against the legacy default stream (stream 0, the one PyTorch
 *    uses by default): device data the caller produced on stream 0 is complete before the
 *    library reads it, and the library's results are complete before the caller's later
 *    stream-0 work.  A caller working on any other stream passes that stream.  Where a
 *    function takes a `stream` ARGUMENT (lsqr_b200_ez_aprod_device, the BLAS-1 functions, the
 *    aprod callback), NULL means the legacy default stream itself, as everywhere in CUDA;
 *    cudaStreamLegacy / cudaStreamPerThread handles are accepted too.
 */
#ifndef LSQR_B200_H
#define LSQR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSQR_B200_VERSION 200

/* The shared library is built with -fvisibility=hidden; only these entry points are exported. */
#if defined(__GNUC__)
#define LSQR_B200_API __attribute__((visibility("default")))
#else
#define LSQR_B200_API
#endif

/* ------------------------------------------------------------------ status codes */
enum {
    LSQR_B200_OK             = 0,
    LSQR_B200_ERR_SIZES      = 1,  /* 'invalid a,icol,irow sizes in initialize_ez'     src/lsqr.f90:109 */
    LSQR_B200_ERR_IROW       = 2,  /* 'invalid irow or m in initialize_ez'             src/lsqr.f90:110 */
    LSQR_B200_ERR_ICOL       = 3,  /* 'invalid icol or n in initialize_ez'             src/lsqr.f90:111 */
    LSQR_B200_ERR_NOINIT     = 4,  /* 'lsqr_solver_ez class not properly initialized'  src/lsqr.f90:152 */
    LSQR_B200_ERR_MODE       = 5,  /* 'invalid mode in aprod_ez'                       src/lsqr.f90:197 */
    LSQR_B200_ERR_INDEX_LOW  = 6,  /* irow < 1 or icol < 1: undefined behaviour in the reference
                                      (no lower-bound check, :110-111); rejected here */
    LSQR_B200_ERR_NO_DEVICE  = 10, /* no usable CUDA device: the engine has no CPU path */
    LSQR_B200_ERR_CUDA       = 11, /* a CUDA runtime call failed; see lsqr_b200_last_error() */
    LSQR_B200_ERR_NCCL       = 12, /* NCCL could not be loaded or a collective failed */
    LSQR_B200_ERR_ARG        = 13, /* NULL / negative / inconsistent argument */
    LSQR_B200_ERR_TOO_LARGE  = 14, /* more than 2^32-2 stored entries on one GPU */
    LSQR_B200_ERR_CALLBACK   = 15  /* the user's aprod callback returned non-zero */
};

/* The reference's literal `error stop` text for codes 1..5, a description otherwise. */
LSQR_B200_API const char *lsqr_b200_error_message(int code);
/* Detail of the most recent failure on the calling thread (CUDA error string, etc.). */
LSQR_B200_API const char *lsqr_b200_last_error(void);
LSQR_B200_API int lsqr_b200_version(void);
/* Number of CUDA devices visible (0 when there is no driver / GPU).  Never fails. */
LSQR_B200_API int lsqr_b200_device_count(void);

/* ------------------------------------------------------------------ callbacks */
/* One line of the `nout` log (src/lsqr.f90:589-595,655-671,813-837,872-880), no newline.
 * The Fortran shim writes it to unit nout. */
typedef void (*lsqr_b200_log_fn)(void *user, const char *line);

/* Per-iteration scalar record, read by the host from pinned mapped memory. */
typedef struct lsqr_b200_iter_record {
    double itn;      /* iteration number (stored as double so the record is 16 doubles) */
    double istop;    /* stopping code decided in this iteration (0 = continue) */
    double x1;       /* x(1) after the update                      */
    double rnorm, test1, test2, anorm, acond;
    double phi, dknorm, dxk, alfopt;
    double alpha, beta, xnorm, arnorm;
} lsqr_b200_iter_record;
typedef void (*lsqr_b200_iter_fn)(void *user, const lsqr_b200_iter_record *rec);

/* ------------------------------------------------------------------ options */
/* Optional arguments of initialize_ez (src/lsqr.f90:101-106, defaults :46-51) plus engine knobs.
 * Obtain defaults with lsqr_b200_default_options(). */
typedef struct lsqr_b200_options {
    double  atol;            /* default 0                                       */
    double  btol;            /* default 0                                       */
    double  conlim;          /* default 0                                       */
    int32_t itnlim;          /* default 100                                     */
    int32_t device;          /* CUDA device ordinal; -1 = the current device    */
    void   *stream;          /* cudaStream_t to run on; NULL = library-owned    */
    lsqr_b200_log_fn  log;   /* nout /= 0 equivalent; NULL = silent             */
    void   *log_user;
    lsqr_b200_iter_fn iter;  /* optional per-iteration record callback          */
    void   *iter_user;
    int32_t engine;          /* 0 = fused kernels (default), 1 = reference pass structure
                                (dscal / aprod / dnrm2 as separate kernels; for A/B parity) */
    int32_t use_graph;       /* 1 (default) = CUDA-graph the iteration, 0 = plain launches */
    int32_t profile;         /* 1 = time every kernel class with CUDA events (slower)  */
    int32_t spmv_variant;    /* 0 or 3: the warp-autonomous segmented kernel (round 1's variants 1 and 2
                                measured slower on every workload and were removed; other values are
                                rejected with LSQR_B200_ERR_ARG) */
    /* --- multi-GPU: A is row-partitioned, one process per GPU ----------- */
    int32_t world_size;      /* 1 (default) = single GPU                         */
    int32_t rank;
    const void *nccl_unique_id;  /* 128 bytes from lsqr_b200_nccl_unique_id() on rank 0,
                                    broadcast by the launcher (torch.distributed) */
    int64_t m_global;        /* total rows over all ranks (for se() and logs)     */
} lsqr_b200_options;
LSQR_B200_API void lsqr_b200_default_options(lsqr_b200_options *opts);

/* Fills 128 bytes with a fresh ncclUniqueId (rank 0 calls it; others receive the bytes). */
LSQR_B200_API int lsqr_b200_nccl_unique_id(void *out128);

/* ------------------------------------------------------------------ class lsqr_solver_ez */
typedef struct lsqr_b200_ez lsqr_b200_ez;

/* Replaces  lsqr_solver_ez%initialize  (initialize_ez, src/lsqr.f90:91-127).
 * size_a/size_irow/size_icol are size(a), size(irow), size(icol) (the three lengths the
 * reference compares at :109).  Validates irow<=m, icol<=n (:110-111), deep-copies the
 * triplets to the GPU (:113-118) and converts them ON DEVICE into CSR for A and a
 * precomputed CSR for A' (stable by COO position; duplicates kept).
 * In a multi-GPU run each rank passes its own block of rows: m = local row count and
 * irow is 1-based inside the block; n and icol are global. */
LSQR_B200_API int lsqr_b200_ez_initialize(lsqr_b200_ez **me, int32_t m, int32_t n,
                            int64_t size_a, const double *a,
                            int64_t size_irow, const int32_t *irow,
                            int64_t size_icol, const int32_t *icol,
                            const lsqr_b200_options *opts /* NULL = defaults */);

/* Replaces  lsqr_solver_ez%solve  (solve_ez, src/lsqr.f90:207-259) and the LSQR loop it
 * drives (src/lsqr.f90:432-882).  b(m) is not modified, x(n) is fully overwritten.
 * se (n) / itn / anorm / acond / rnorm / arnorm / xnorm are the reference's optional
 * outputs: pass NULL to omit (se == NULL means wantse = .false.). */
LSQR_B200_API int lsqr_b200_ez_solve(lsqr_b200_ez *me, const double *b, double damp,
                       double *x, int32_t *istop,
                       double *se, int32_t *itn, double *anorm, double *acond,
                       double *rnorm, double *arnorm, double *xnorm);

/* Replaces  lsqr_solver_ez%aprod  (aprod_ez, src/lsqr.f90:134-200).
 * mode 1: y(m) = y + A*x(n);  mode 2: x(n) = x + A'*y(m).  m,n must equal the handle's. */
LSQR_B200_API int lsqr_b200_ez_aprod(lsqr_b200_ez *me, int32_t mode, int32_t m, int32_t n,
                       double *x, double *y);

/* The reference frees its allocatable components automatically; the shim calls this from a
 * `final` procedure.  NULL is accepted. */
LSQR_B200_API void lsqr_b200_ez_destroy(lsqr_b200_ez *me);

/* Changes atol/btol/conlim/itnlim/log/iter of an initialized handle (the reference would
 * re-run initialize; this avoids rebuilding the matrix). */
LSQR_B200_API int lsqr_b200_ez_set_options(lsqr_b200_ez *me, const lsqr_b200_options *opts);

/* Parity inspection (no reference counterpart): copies the device-built CSR (which = 0) or
 * CSR of A' (which = 1) to HOST arrays: ptr[nkeys+1] (0-based), idx[nnz] (0-based other
 * coordinate), val[nnz], perm[nnz] (0-based COO position of each stored entry).
 * Any output may be NULL. */
LSQR_B200_API int lsqr_b200_ez_get_csr(lsqr_b200_ez *me, int32_t which,
                         int64_t *ptr, int32_t *idx, double *val, int64_t *perm);
/* The same arrays in place, as DEVICE pointers owned by the handle (ptr is 32-bit on the device; any output may
 * be NULL): lets full-size property tests inspect the layout without a 10 GB host copy. */
LSQR_B200_API int lsqr_b200_ez_get_csr_device(lsqr_b200_ez *me, int32_t which, const uint32_t **ptr_dev,
                                const int32_t **idx_dev, const double **val_dev, const uint32_t **perm_dev);
LSQR_B200_API int64_t lsqr_b200_ez_nnz(const lsqr_b200_ez *me);
/* Layout of the stored matrices.  When the gathered vector of a product cannot stay in L2 the matrix is kept
 * BLOCKED along the gathered coordinate, one CSR per block, stored back to back:
 *   which = 1 (A', gathers u, 8 m bytes): blocks of block_size consecutive ROWS of A;
 *   which = 0 (A,  gathers v, 8 n bytes): blocks of block_size consecutive COLUMNS of A.
 * lsqr_b200_ez_get_csr(which) then returns ptr[nblocks*nkeys + 1]: block b is the stable (by COO position) sort by
 * key of the entries whose other coordinate lies in [b*block_size, (b+1)*block_size).
 * nblocks = 1 (block_size = 0) is the plain CSR. */
LSQR_B200_API int lsqr_b200_ez_blocks(const lsqr_b200_ez *me, int32_t which, int64_t *nblocks, int64_t *block_size);
/* Work schedule of the SpMV kernel over block `block` of the stored A (which = 0) or A' (which = 1): number of
 * row-aligned tiles, nominal stored entries per tile, whether the tiles are dealt to the warps by the balanced
 * (largest-first) schedule used for very uneven row lengths instead of round robin, and the load of the most
 * loaded warp relative to the mean under the schedule in use.  Diagnostic; no reference counterpart. */
LSQR_B200_API int lsqr_b200_ez_schedule(const lsqr_b200_ez *me, int32_t which, int64_t block, int64_t *ntiles,
                          int64_t *tile_entries, int32_t *balanced, double *imbalance);

/* The work plan of the SpMV kernel over the stored A (which = 0) or A' (which = 1): one plan covers every block of
 * the matrix.  Diagnostic; no reference counterpart. */
typedef struct lsqr_b200_plan_info {
    int64_t nblocks;            /* blocks along the gathered coordinate (1 = plain CSR)                         */
    int64_t ntiles;             /* row-aligned tiles; the same row cuts in every block                          */
    int32_t grid_ctas;          /* persistent grid                                                              */
    int32_t ctas_per_sm;        /* resident CTAs per SM of the kernel flavour: 4 (EPL 4) or 2 (EPL 8)            */
    int32_t window_doubles;     /* shared-memory gather window per warp, in doubles (0 = gathers stay global)   */
    int32_t balanced;           /* largest-first tile schedule in use (very uneven row lengths)                 */
    double  windowed_fraction;  /* fraction of the stored entries whose gathers are served from shared memory   */
    double  imbalance;          /* most loaded warp / mean load                                                 */
    int64_t span_median, span_max;  /* gather span of the pieces (entries of the dense vector)                  */
    int32_t single_launch;      /* one persistent launch walks every block of a product                         */
    int32_t peer_exchange;      /* multi-GPU: exchange over NVLink peer memory instead of one NCCL all-reduce   */
    int32_t entries_per_lane;   /* kernel flavour: 4 (gather-bound matrices) or 8 (matrices with local gathers)  */
    int32_t flavour;            /* kernel flavour: 0 = local gathers, 1 = staged gather windows, 2 = random columns */
    double  lines_per_gather;   /* 128-byte lines spanned by 32 consecutive stored entries (32 = no locality)   */
} lsqr_b200_plan_info;
LSQR_B200_API int lsqr_b200_ez_plan(const lsqr_b200_ez *me, int32_t which, lsqr_b200_plan_info *out);

/* Kernel timing of the most recent solve.  loop_ms / init_ms / total_launches are always
 * filled; the per-kernel averages only when options.profile = 1 (CUDA-event pairs around every
 * launch inside the real iteration loop, so L2 contents are those of the real loop). */
typedef struct lsqr_b200_kernel_times {
    double  aprod_ms, atprod_ms, update_ms, other_ms;  /* average per launch            */
    int64_t aprod_launches, atprod_launches, update_launches, other_launches;
    int64_t total_launches;    /* every kernel of this library launched by the last solve */
    double  loop_ms;           /* device time of the iteration loop of the last solve     */
    double  init_ms;           /* device time before the first iteration (b -> u, v, w)   */
    int64_t iteration_launches;/* kernels of this library per LSQR iteration              */
} lsqr_b200_kernel_times;
LSQR_B200_API int lsqr_b200_ez_get_kernel_times(const lsqr_b200_ez *me, lsqr_b200_kernel_times *out);

/* ------------------------------------------------------------------ class lsqr_solver (operator hook) */
/* Replaces the deferred type-bound procedure  aprod  (aprod_func, src/lsqr.f90:67-82) for
 * device-resident vectors.  mode 1: y_dev(m) += A*x_dev(n); mode 2: x_dev(n) += A'*y_dev(m).
 * The callback must enqueue its work on `stream` (a cudaStream_t) and must not synchronize.
 * Return 0 on success. */
typedef int (*lsqr_b200_aprod_fn)(void *user, int32_t mode, int32_t m, int32_t n,
                                  double *x_dev, double *y_dev, void *stream);

/* Replaces  lsqr_solver%lsqr  (LSQR, src/lsqr.f90:432-882) with a caller-supplied operator and
 * caller-supplied DEVICE storage u(m) [in: b, overwritten], v(n), w(n), x(n) [out],
 * se(n) [touched only if wantse].  Scalar outputs are host pointers (NULL to omit).
 * opts supplies stream/device/log/iter; its atol..itnlim fields are ignored here.
 * Iterations are enqueued in batches without host synchronisation (the device stops itself), so
 * the callback may be invoked a few times after the stopping iteration: x, se and every scalar are
 * those of the stopping iteration, u and v are work space on return (the reference leaves the last
 * Lanczos vectors there). */
LSQR_B200_API int lsqr_b200_lsqr(lsqr_b200_aprod_fn aprod, void *aprod_user,
                   int32_t m, int32_t n, double damp, int32_t wantse,
                   double *u_dev, double *v_dev, double *w_dev, double *x_dev, double *se_dev,
                   double atol, double btol, double conlim, int32_t itnlim,
                   const lsqr_b200_options *opts,
                   int32_t *istop, int32_t *itn, double *anorm, double *acond,
                   double *rnorm, double *arnorm, double *xnorm);

/* Replaces  lsqr_solver%acheck  (src/lsqr.f90:908-994): adjoint test of a device operator.
 * v(n), w(m), x(n), y(m) are DEVICE work vectors.  inform = 0 if consistent. */
LSQR_B200_API int lsqr_b200_acheck(lsqr_b200_aprod_fn aprod, void *aprod_user, int32_t m, int32_t n,
                     double eps, double *v_dev, double *w_dev, double *x_dev, double *y_dev,
                     const lsqr_b200_options *opts, int32_t *inform, double *relerr);

/* Replaces  lsqr_solver%xcheck  (src/lsqr.f90:1015-1154).  b(m), x(n) DEVICE inputs;
 * u(m) <- r = b - A x, v(n) <- A'r, w(n) <- A'r - damp^2 x  (DEVICE outputs).
 * norms (host, 6 doubles or NULL): bnorm, xnorm, rho1, sigma1, rho2, sigma2. */
LSQR_B200_API int lsqr_b200_xcheck(lsqr_b200_aprod_fn aprod, void *aprod_user, int32_t m, int32_t n,
                     double anorm, double damp, double eps,
                     const double *b_dev, double *u_dev, double *v_dev, double *w_dev,
                     const double *x_dev, const lsqr_b200_options *opts,
                     int32_t *inform, double *test1, double *test2, double *test3,
                     double *norms);

/* ---- the abstract class with the reference's own signatures: HOST arrays, HOST operator -------------------- */
/* Replaces the deferred  aprod  exactly as the reference declares it (aprod_func, src/lsqr.f90:67-82): x(n) and
 * y(m) are HOST arrays, both intent(inout); mode 1: y = y + A*x, mode 2: x = x + A'*y.  Return 0 on success. */
typedef int (*lsqr_b200_aprod_host_fn)(void *user, int32_t mode, int32_t m, int32_t n, double *x, double *y);

/* lsqr_solver%lsqr / %acheck / %xcheck with the reference's argument lists (src/lsqr.f90:432-435, :908, :1015):
 * every vector is a HOST (or device) array and the operator is host code, so each product makes a host round trip;
 * all vector arithmetic and the scalar recurrence still run on the GPU.  This is the form an unmodified
 * type,extends(lsqr_solver) of the reference binds to (fortran/lsqr_b200_shim.F90); supply a device operator through
 * lsqr_b200_lsqr for speed.  u(m) holds b on entry and is overwritten; v, w, x (n), se (n, if wantse) are outputs. */
LSQR_B200_API int lsqr_b200_lsqr_host(lsqr_b200_aprod_host_fn aprod, void *aprod_user,
                        int32_t m, int32_t n, double damp, int32_t wantse,
                        double *u, double *v, double *w, double *x, double *se,
                        double atol, double btol, double conlim, int32_t itnlim,
                        const lsqr_b200_options *opts,
                        int32_t *istop, int32_t *itn, double *anorm, double *acond,
                        double *rnorm, double *arnorm, double *xnorm);
LSQR_B200_API int lsqr_b200_acheck_host(lsqr_b200_aprod_host_fn aprod, void *aprod_user, int32_t m, int32_t n,
                          double eps, double *v, double *w, double *x, double *y,
                          const lsqr_b200_options *opts, int32_t *inform, double *relerr);
LSQR_B200_API int lsqr_b200_xcheck_host(lsqr_b200_aprod_host_fn aprod, void *aprod_user, int32_t m, int32_t n,
                          double anorm, double damp, double eps,
                          const double *b, double *u, double *v, double *w, const double *x,
                          const lsqr_b200_options *opts,
                          int32_t *inform, double *test1, double *test2, double *test3, double *norms);

/* An lsqr_b200_aprod_fn backed by an ez handle (user = the lsqr_b200_ez*), so the ez matrix can
 * be driven through the low-level path exactly like  class(lsqr_solver_ez) -> lsqr_solver. */
LSQR_B200_API int lsqr_b200_ez_aprod_device(void *ez_handle, int32_t mode, int32_t m, int32_t n,
                              double *x_dev, double *y_dev, void *stream);

/* ------------------------------------------------------------------ device BLAS-1 (src/lsqrblas.f90) */
/* Deterministic device versions of the reference's vector kernels (stride 1 only -- the only form the hot path
 * uses; the Fortran layer packs strided arguments).  The arrays may be DEVICE arrays (used in place) or HOST arrays
 * (staged through the GPU: the arithmetic always happens on the device).  stream NULL = legacy default stream.
 * dnrm2 is the SCALED norm of the reference (it neither overflows nor underflows, :136-154), computed with Blue's
 * three accumulators. */
LSQR_B200_API int lsqr_b200_dnrm2(int64_t n, const double *x_dev, double *result_host, void *stream);  /* :123-159 */
LSQR_B200_API int lsqr_b200_ddot (int64_t n, const double *x_dev, const double *y_dev, double *result_host, void *stream); /* :74-116 */
LSQR_B200_API int lsqr_b200_dscal(int64_t n, double da, double *x_dev, void *stream);                  /* :166-201 */
LSQR_B200_API int lsqr_b200_dcopy(int64_t n, const double *x_dev, double *y_dev, void *stream);        /* :25-67 */

#ifdef __cplusplus
}
#endif
#endif /* LSQR_B200_H */
