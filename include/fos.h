/*
 * fos.h -- C ABI of libfos_b200.so, the B200 (sm_100a) solver core behind
 * FastOptSolver's Python module API.
 *
 * The reference (ElBaldo1/FastOptSolver) has no FFI layer: its boundary is the
 * flat Python modules iterative_solvers.py / prox_operators.py /
 * objective_functions.py / lbfgs.py.  Each entry point below names the reference
 * interface it replaces (file:line into the reference tree).  The Python drop-in
 * (the modules under fastoptsolver_b200/dropin) binds these with ctypes; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns FOS_OK (0) or a negative fos_status; the message of
 *     the last failure on the calling thread is fos_last_error();
 *   - all pointers are HOST pointers unless the name ends in _dev;
 *   - vectors are float64; the design matrix A is float64 or float32 *storage*
 *     (arithmetic is always float64, as in the reference where numpy up-casts:
 *     iterative_solvers.py:150);
 *   - a design handle is NOT re-entrant (the reference keeps module-global metric
 *     lists and is single-threaded too: iterative_solvers.py:16-18).
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with FOS_ERR_CUDA.
 */
#ifndef FOS_B200_H
#define FOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOS_ABI_VERSION 1

typedef enum fos_status {
    FOS_OK = 0,
    FOS_ERR_INVALID = -1,  /* bad argument -> ValueError in the drop-in */
    FOS_ERR_CUDA = -2,     /* CUDA runtime / launch failure -> RuntimeError */
    FOS_ERR_NOMEM = -3,    /* device or pinned allocation failed */
    FOS_ERR_UNSUPPORTED = -4,
    FOS_ERR_COMM = -5      /* multi-GPU exchange failed */
} fos_status;

typedef enum fos_dtype { FOS_F64 = 0, FOS_F32 = 1 } fos_dtype;

/* momentum scheme of the proximal-gradient engine */
typedef enum fos_scheme {
    FOS_SCHEME_NESTEROV = 0, /* fista        iterative_solvers.py:132-245 */
    FOS_SCHEME_DELTA = 1,    /* fista_delta  iterative_solvers.py:251-344 */
    FOS_SCHEME_ISTA = 2      /* ista         iterative_solvers.py:65-125 (framework-owned callables) */
} fos_scheme;

typedef enum fos_stop {
    FOS_STOP_MAXITER = 0,
    FOS_STOP_GRADNORM = 1, /* iterative_solvers.py:179 */
    FOS_STOP_STEP = 2,     /* iterative_solvers.py:238, :337, :122 */
    FOS_STOP_RATIO = 3     /* iterative_solvers.py:242, :341 */
} fos_stop;

typedef struct fos_design fos_design; /* opaque: A and b resident in HBM + workspaces */

/* ---- library / device -------------------------------------------------------------- */
int fos_abi_version(void);
const char* fos_last_error(void);
/* number of visible CUDA devices (0 on a CPU box; never fails) */
int fos_device_count(void);
/* sm_count, total/free HBM bytes of `device` */
int fos_device_info(int device, int* sm_count, size_t* total_bytes, size_t* free_bytes);
/* release device memory the library keeps between calls: the split workspace of the upload-time
 * Gram accumulation (up to 1.5 GB per device) and the recycled matrix block -- the HBM block of the
 * last destroyed design (one per device, >= 256 MB), kept because the reference's callers upload the
 * same design again on every solver call (iterative_solvers.py:133-134 re-reads A) and a fresh
 * allocation + free of tens of GB costs ~0.2 s; FOS_KEEP_BLOCK=0 disables the recycling */
int fos_trim(void);

/* ---- design: A (n x d) and b (n) resident on one GPU ---------------------------------
 * Replaces the numpy arrays every reference solver takes as (A, b)
 * (iterative_solvers.py:133-134, lbfgs.py:41, objective_functions.py:3).
 * The matrix is copied into a row-major device layout whose leading dimension is
 * padded to 16 bytes; row_stride/col_stride are the host strides in ELEMENTS, so
 * C-order is (d, 1) and Fortran order is (1, n).  The host arrays are never written. */
int fos_design_create(const void* A, const double* b, int64_t n, int64_t d, int dtype,
                      int64_t row_stride, int64_t col_stride, int device, fos_design** out);
/* Two-step form of fos_design_create: _begin allocates the design (matrix undefined), _upload copies the
 * host arrays into it (same layouts and staging as fos_design_create).  Between the two calls the handle
 * may already be wired to its peers (fos_comm_window_alloc* / fos_comm_attach* from another host thread):
 * row-sharded callers overlap the window hand-off with the PCIe copy.  Solver entry points must not be
 * called before _upload has returned. */
int fos_design_create_begin(int64_t n, int64_t d, int dtype, int device, fos_design** out);
int fos_design_upload(fos_design* h, const void* A, const double* b, int64_t row_stride, int64_t col_stride);
/* Borrow a row-major matrix already resident on `device` (lda in elements, lda*elem
 * multiple of 16 bytes, base 16-byte aligned); b_dev is float64.  Not freed by destroy. */
int fos_design_create_device(const void* A_dev, const double* b_dev, int64_t n, int64_t d,
                             int dtype, int64_t lda, int device, fos_design** out);
/* Generate the correlated-column synthetic design directly in HBM with a counter-based
 * (Philox4x32-10) generator: population-standardised version of the recipe in
 * easy_boston_data.py:23-43 generalised to d columns; rows [row0, row0+n) of a virtual
 * n_total-row design, so that R ranks generate disjoint shards of the same matrix. */
int fos_design_create_synthetic(int64_t n, int64_t d, int dtype, uint64_t seed, double noise_std,
                                double rho1, double rho2, int64_t row0, int device,
                                fos_design** out);
int fos_design_destroy(fos_design* h);
int fos_design_shape(const fos_design* h, int64_t* n, int64_t* d, int* dtype, int64_t* lda);
/* copy rows [row0, row0+rows) back to the host as a dense C-order block (for the CPU
 * baseline, which must see the same numbers) */
int fos_design_download(fos_design* h, int64_t row0, int64_t rows, void* A_out, double* b_out);
/* device pointers of the resident arrays (for zero-copy wrappers) */
int fos_design_pointers(fos_design* h, void** A_dev, double** b_dev);
/* Gram matrix accumulated under the upload.  fos_design_create copies a dense float64 C-order
 * design with n >= 16 d, d a multiple of 128 and <= 4096, >= 1 GB (FOS_UPLOAD_GRAM=1/0 forces it
 * on/off) in 512 MB row chunks and runs the tensor-core SYRK of fos_gram_create on every chunk
 * that has arrived, on a second stream: G = A^T A of the local rows is complete a few ms after
 * the last byte.  fos_power_iter (estimate_lipschitz, iterative_solvers.py:45-60) then iterates
 * w = G v instead of w = A^T(A v): the same numbers up to rounding (relative 1e-15 on L), 100 x
 * d^2 instead of 100 x n d doubles of traffic.  fos_gram_create reuses the matrix as well.
 * state: 0 none, 1 local rows, 2 local rows and confirmed on every rank.  Row-sharded designs
 * never sum the matrices: A^T A v = sum_r G_r v, so each step is a local product whose result goes
 * through the same fused peer-memory all-reduce as a streaming pass (32 KB per step).  All ranks
 * must take the same branch, so the caller first agrees that EVERY rank holds its matrix and then
 * calls fos_design_upload_gram_set(h, 2) (else (h, 0) on all ranks, which discards it); until then
 * a sharded design keeps the streaming power iteration.  fos_gram_create copies the local matrix
 * in either state (its caller sums the copies). */
int fos_design_upload_gram(fos_design* h, double** G_dev, int* state, float* copy_ms, float* tail_ms);
int fos_design_upload_gram_set(fos_design* h, int state);
/* Column statistics of the resident design (one pass over A): out[c] = sum_i (A[i][c] - center[c])^p
 * with p = 2 if `squared` else 1 (center may be NULL = 0; it has d+1 entries, the last for b);
 * *b_out gets the same statistic of b.  Local rows only: sharded callers add the ranks' results. */
int fos_design_column_sums(fos_design* h, const double* center, int squared, double* out, double* b_out);
/* In place: A[i][c] = (A[i][c] - shift[c]) / scale[c], b[i] -= b_shift -- the z-scoring the reference's
 * notebook applied to its data before calling the solvers. */
int fos_design_affine(fos_design* h, const double* shift, const double* scale, double b_shift);
/* enable/disable per-launch CUDA-event timing of the gradient kernel inside the solver loop */
int fos_design_set_profile(fos_design* h, int enable);
/* Diagnostic: average duration (CUDA events, solver stream) of `reps` back-to-back launches of
 * the gradient kernel in a given mode: 1 = gradient, 3 = gradient + second residual norm,
 * 2 = residual norm only, 8 = streaming probe (bulk-copy ring only, no arithmetic: the sustained
 * HBM read ceiling of this pipeline). */
int fos_time_grad_kernel(fos_design* h, int mode, int reps, float* ms_avg);
/* Debug: per-CTA start/end timestamps (ns) of one gradient-kernel launch; out[2*n_parts]. */
int fos_debug_cta_times(fos_design* h, int mode, long long* out, int cap, int* n_parts);
/* Row-sharded designs: may this rank run its solves in the persistent solve kernel (one launch per solve,
 * push-model slice exchange)?  All ranks must take the same path, so the binding gathers this flag over
 * the ranks when the windows are attached and calls _disable everywhere unless every rank said yes. */
int fos_design_solve_kernel_ok(const fos_design* h, int world, int* ok);
int fos_design_solve_kernel_disable(fos_design* h);
/* Debug: phase profile of the persistent solve kernel (CTA 0's %globaltimer stamps, ns, accumulated
 * since the design was created): out[0..7] = streaming loop, wait at barrier 1, slice sums, peer exchange,
 * elementwise 1, barrier 2, scalars + decision + elementwise 2 + commit, barrier 3; out[8] = passes.
 * reset != 0 zeroes the accumulators after reading.  All zeros for designs that never ran it. */
int fos_debug_solve_profile(fos_design* h, unsigned long long* out9, int reset);
/* lambda_max = ||A^T b||_inf (one fused pass), the usual scale for alpha1 */
int fos_design_lambda_max(fos_design* h, double* out);

/* ---- multi-GPU: rows of A sharded over `world` processes, one per GPU -----------------
 * The exchange step is one all-reduce of the d-length A^T r partial (+2 scalars) per pass.
 * Peer-memory mode: every rank allocates an exchange window with fos_comm_window_alloc,
 * publishes its cudaIpcMemHandle (64 bytes) out of band (the drop-in uses
 * torch.distributed.all_gather), and maps the peers' windows with fos_comm_attach; the
 * epilogue kernel then reduces across NVLink in fixed rank order inside the same launch. */
int fos_comm_window_alloc(fos_design* h, int rank, int world, void* ipc_handle_out64);
int fos_comm_attach(fos_design* h, const void* ipc_handles /* world x 64 bytes */, int world);
int fos_comm_info(const fos_design* h, int* rank, int* world);
/* Preferred variant: the window is a cuMemCreate allocation exported as a POSIX file descriptor
 * (*fd_out, owned by the design); the caller passes the descriptors around (unix socket,
 * SCM_RIGHTS) and hands fos_comm_attach_fd the ones it RECEIVED (fds[r] = rank r's window as a
 * descriptor valid in this process; own entry ignored; the caller closes them afterwards).  A peer
 * maps the window for its own device only, so -- unlike cudaIpcOpenMemHandle, which enables peer
 * access for the device pair and makes every later cudaMalloc of the process peer-visible (200-260
 * ms for a fresh 8 GB block with 3 peers) -- nothing else is affected.  FOS_ERR_UNSUPPORTED: use
 * the cudaIpc pair above. */
int fos_comm_window_alloc_fd(fos_design* h, int rank, int world, int* fd_out);
int fos_comm_attach_fd(fos_design* h, const int* fds, int world);
/* drop a window that has not been attached yet (either kind), e.g. to fall back to the other variant */
int fos_comm_window_free(fos_design* h);

/* ---- one-shot operators ----------------------------------------------------------------
 * fos_grad: loss = 0.5||Ax-b||^2 (+0.5 a2 ||x||^2), g = A^T(Ax-b) (+a2 x), A read ONCE.
 *   replaces the inlined gradient iterative_solvers.py:173-175, :292-294 and `fg`
 *   lbfgs.py:46-51. */
int fos_grad(fos_design* h, const double* x, double alpha2, double* g_out, double* loss_out);
/* fos_objective: objective_functions.py:3-30.  reg: bit0 = add alpha1*||x||_1,
 * bit1 = add 0.5*alpha2*||x||^2 (lasso=1, ridge=2, elasticnet=3). */
int fos_objective(fos_design* h, const double* x, int reg_bits, double alpha1, double alpha2,
                  double* out);
/* fos_power_iter: estimate_lipschitz, iterative_solvers.py:45-60.  v0 is the caller's
 * normalised start vector (the drop-in draws it from numpy's legacy global RNG so the
 * stream advances exactly as in the reference, :50). */
int fos_power_iter(fos_design* h, const double* v0, int n_iter, double tol, double* L_out,
                   int* iters_out, float* gpu_ms_out);
/* prox_operators.py:3-8 and :10-16 on a flat host buffer (any shape flattened).
 * scale = 1/(1+tau*alpha2) for the elastic-net prox, 1.0 for prox_l1. */
int fos_prox_l1(const double* v, int64_t len, double thresh, double* out, int device);
int fos_prox_elastic_net(const double* v, int64_t len, double tau, double alpha1, double alpha2,
                         double* out, int device);

/* ---- the proximal-gradient engine: fista / fista_delta / ista ---------------------------- */
typedef struct fos_pg_params {
    int scheme;            /* fos_scheme */
    double alpha1, alpha2; /* L1 weight (prox), L2 weight (smooth part) */
    int obj_terms;         /* bits as fos_objective.reg_bits: terms in the recorded objective
                              (fista: from alpha>0, :227-230; fista_delta: from reg_type, :321) */
    double delta;          /* FISTA-delta momentum parameter (> 2) */
    int backtracking;      /* Armijo rule of the reference, :183-197 */
    double eta;
    double armijo_c;       /* module global C, iterative_solvers.py:11 */
    double step0;          /* initial step t_init_factor / L */
    int max_iter;
    double tol, tol_ratio;
    int adaptive_restart;
    double restart_threshold;
    int want_history;      /* record iterates (and objectives unless scheme == ISTA) */
    const double* x0;      /* start point (ISTA only; NULL = zeros) */
} fos_pg_params;

typedef struct fos_pg_result {
    /* caller-allocated outputs (NULL = not wanted) */
    double* x;        /* d                     final iterate */
    double* x_hist;   /* (max_iter+1) x d      row 0 = x0, row k = x_k   (want_history) */
    double* obj_hist; /* max_iter              objective of x_1..         (want_history) */
    double* t_hist;   /* max_iter+1            step size after each iteration (t_hist[0]=step0) */
    double* step_hist;/* max_iter              ||x_{k+1}-x_k|| */
    int* ls_iters;    /* max_iter              Armijo shrink count per iteration */
    float* grad_ms;   /* max_iter+1            device time of each gradient pass */
    float* ls_ms;     /* max_iter              device time of each line search */
    /* scalars written by the call */
    int n_iters;      /* completed iterations */
    int n_grad_calls; /* gradient passes (== iterations, +1 when the gradient-norm stop fired) */
    int n_passes;     /* passes over A, everything included */
    int stop_reason;  /* fos_stop */
    float loop_ms;    /* device time of the whole loop (CUDA events on the solver stream) */
    int64_t kernel_launches;
    /* filled when fos_design_set_profile(h, 1): CUDA events around every gradient-kernel
     * launch of the loop (on the solver stream), summed */
    float grad_kernel_ms;
    int grad_kernel_launches;
    /* device-side (%globaltimer) totals over the loop: time inside the epilogue kernel, and the
     * part of it spent publishing to / waiting for the peer ranks (0 on one GPU) */
    float epilogue_ms;
    float exchange_ms;
    /* host wall clock of the call: workspace + state upload, launching and waiting for the loop,
     * downloading the results */
    float host_setup_ms, host_loop_ms, host_finish_ms;
} fos_pg_result;

int fos_prox_grad(fos_design* h, const fos_pg_params* p, fos_pg_result* r);

/* ---- device-resident L-BFGS ------------------------------------------------------------------
 * The unconstrained path of L-BFGS-B as scipy.optimize.fmin_l_bfgs_b runs it for the reference
 * (lbfgs.py:64-70: m = 10, factr = 1e7, pgtol = tol, maxfun = 15000, maxls = 20), with the
 * (S, Y) history, the two-loop recursion and the More'-Thuente line search on the device.
 * Minimises 0.5||Ax-b||^2 + 0.5 alpha2 ||x||^2 (the reference's `fg`, lbfgs.py:43-54: the L1 term
 * is never in it); obj_hist records the full objective of every accepted iterate (the reference's
 * callback, lbfgs.py:56-61; obj_terms as in fos_objective).  stop_reason: 1 ||g||_inf <= pgtol,
 * 2 relative reduction <= factr*eps, 3 max_iter, 4 maxfun, 5 abnormal line-search termination. */
typedef struct fos_lbfgs_params {
    int m, max_iter, maxfun, maxls, obj_terms;
    double alpha1, alpha2, pgtol, factr;
    const double* x0; /* NULL = zeros (lbfgs.py:63) */
} fos_lbfgs_params;
typedef struct fos_lbfgs_result {
    double* x;        /* d */
    double* obj_hist; /* max_iter */
    double f_final;   /* smooth loss at x (res[1] of fmin_l_bfgs_b, lbfgs.py:72) */
    int n_iters, n_fg, n_skipped, stop_reason;
    float loop_ms;
    int64_t kernel_launches;
} fos_lbfgs_result;
int fos_lbfgs(fos_design* h, const fos_lbfgs_params* p, fos_lbfgs_result* r);

/* ---- Gram-matrix mode (n >> d) and the batched regularisation path --------------------------
 * New capability (north_star item 4; no counterpart in the reference, which re-reads A every
 * iteration): G = A^T A, c = A^T b, b^T b are built once with fp64 tensor-core MMA, after
 * which a FISTA iteration for Lambda penalties at once is one d x d x Lambda contraction
 *   Grad = G Y - c 1^T (+alpha2 Y)   (== A^T(A y - b) + alpha2 y of iterative_solvers.py:173-175)
 * followed by the same prox / Nesterov update as fista (:200-221), fixed step, per column.
 * Needs float64 storage and d a multiple of 128. */
typedef struct fos_gram fos_gram;
int fos_gram_create(fos_design* h, fos_gram** out);
int fos_gram_destroy(fos_gram* g);
int fos_gram_info(const fos_gram* g, int* d, double* btb, float* build_ms, int* nsplit);
int fos_gram_pointers(fos_gram* g, double** G_dev, double** c_dev);
int fos_gram_download(fos_gram* g, double* G_out, double* c_out);
int fos_gram_set_btb(fos_gram* g, double btb);
/* Debug: how many tile-product (SYRK) launches of this process were staged by the TMA unit (tensor maps, SASS UTMALDG)
 * and how many by the cp.async kernel (FOS_GRAM_TMA=0, an odd row pitch, or no tensor-map encoder in the driver). */
int fos_debug_gram_staging(long long* tma_launches, long long* cp_async_launches);
/* Debug: the same count for fos_gram_path_fista calls (FOS_PATH_TMA=0 selects the cp.async ring; results are
 * bit-identical between the two). */
int fos_debug_path_staging(long long* tma_solves, long long* cp_async_solves);
/* Host logic only (runs without a GPU): the schedule fos_gram_path_fista picks for d features (multiple of 128) and
 * n_lambda penalties on a part with sm_count SMs -- padded penalty count, tile shape, stream-K or one tile per CTA,
 * number of 128-row tiles.  Honours FOS_PATH_SK / FOS_PATH_TN like the solver does. */
int fos_debug_path_plan(int d, int n_lambda, int sm_count, int* padded_lambdas, int* tile_rows, int* tile_cols,
                        int* stream_k, long long* n_tiles);
/* Strong-rule screening on the regularisation path (the per-column
 * semantics stay those of fista, iterative_solvers.py:199-221).
 * fos_gram_subset: the Gram system restricted to the strictly increasing feature indices idx[0..n_idx):
 *   G_S = G[idx][:, idx], c_S = c[idx], zero padded to a multiple of 128 columns (padded features stay 0);
 *   b^T b is inherited, so objectives of the restricted problem equal those of the full problem at x_{S^c} = 0.
 * fos_gram_apply: out[l] = G X[l] - c for n_cols host vectors of length d (row-major n_cols x d): the gradient of
 *   the smooth part A^T(A x - b) (iterative_solvers.py:173) without the alpha2 term -- the quantity the strong
 *   rule thresholds and the KKT re-check bounds by alpha1. */
int fos_gram_subset(fos_gram* g, const int* idx, int n_idx, fos_gram** out);
int fos_gram_apply(fos_gram* g, const double* X, int n_cols, double* out);
typedef struct fos_path_params {
    const double* alphas1; /* n_lambda L1 weights */
    int n_lambda;
    double alpha2;         /* shared L2 weight (smooth part) */
    double step;           /* t_init_factor / L, shared by all columns */
    int max_iter;
    double tol;            /* > 0: stop once every column's step norm is below tol (iterative_solvers.py:238) */
    int check_every;       /* evaluate the stop test every this many iterations */
    const double* X0;      /* n_lambda x d warm start, NULL = zeros (iterative_solvers.py:150) */
} fos_path_params;
typedef struct fos_path_result {
    double* X;   /* n_lambda x d: row l = solution for alphas1[l] */
    double* obj; /* n_lambda: 0.5 x^T G x - c^T x + 0.5 b^T b (+0.5 alpha2 |x|^2) (+alpha1 |x|_1) */
    int n_iters;
    double last_max_step;
    int tile_rows;
    float loop_ms;
    int64_t kernel_launches;
} fos_path_result;
int fos_gram_path_fista(fos_gram* g, const fos_path_params* p, fos_path_result* r);

/* ---- streaming multi-RHS mode: the same batched fixed-step FISTA WITHOUT the Gram matrix ----------
 * For designs where d^2 does not fit or n is not >> d.  Every iteration streams A once per batch of 8
 * penalties; U = A Y - b 1^T and A^T U (iterative_solvers.py:173-175, batched over the columns) both
 * run on the fp64 tensor pipe inside one kernel, a cluster of 4 CTAs sharing a row block (csrc/
 * mrhs_kernels.cu).  Column l reproduces fista(A, b, ..., alpha1 = alphas1[l], alpha2, backtracking =
 * False) (iterative_solvers.py:132-245).  r->obj = 0.5 |A x - b|^2 (+0.5 alpha2 |x|^2) (+alpha1 |x|_1),
 * from one norms-only pass per batch.  Dense float64 designs with d in {1024, 2048, 4096}, one GPU;
 * FOS_ERR_UNSUPPORTED otherwise. */
int fos_mrhs_fista(fos_design* h, const fos_path_params* p, fos_path_result* r);

#ifdef __cplusplus
}
#endif
#endif /* FOS_B200_H */
