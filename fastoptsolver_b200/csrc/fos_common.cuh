// fos_common.cuh -- shared host/device definitions of libfos_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fos.h"

// ------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------
void fos_set_error(const char* fmt, ...);

#define FOS_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            fos_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                    \
                          cudaGetErrorString(_e));                                         \
            return (_e == cudaErrorMemoryAllocation) ? FOS_ERR_NOMEM : FOS_ERR_CUDA;       \
        }                                                                                  \
    } while (0)

#define FOS_REQUIRE(cond, ...)                                                             \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            fos_set_error(__VA_ARGS__);                                                    \
            return FOS_ERR_INVALID;                                                        \
        }                                                                                  \
    } while (0)

#define FOS_TRY(expr)                                                                      \
    do {                                                                                   \
        int _s = (expr);                                                                   \
        if (_s != FOS_OK) return _s;                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------
// pass modes of the gradient kernel (bit set, lives in FosCtrl::g_mode)
// ------------------------------------------------------------------------------------------
enum : int {
    GM_SKIP = 0,
    GM_GRAD = 1,  // r1 = A v1 - b ; partial_g += A^T r1 ; s1 += r1^2
    GM_DOT2 = 2,  // r2 = A v2 - b ; s2 += r2^2
    GM_NOB = 4,   // treat b as zero (power iteration)
    GM_PROBE = 8, // diagnostic: run the bulk-copy ring only, consumers discard the data
    GM_QREC = 16, // with GM_GRAD: s2 = |A x_k - b|^2 from the residual recurrence instead of a second dot:
                  // y_k = x_k + beta (x_k - x_{k-1})  =>  A x_k - b = (r_y + beta q_{k-1}) / (1 + beta), q stored per row
};

// phases of the proximal-gradient state machine (FosCtrl::phase)
enum : int {
    PH_GRAD = 0,      // next pass: gradient at y (+ lagged objective of x_k)
    PH_TRIAL = 1,     // next pass: smooth value of the Armijo candidate
    PH_FINALOBJ = 2,  // next pass: objective of the last iterate only
    PH_DONE = 3,
};

// epilogue operations
enum : int {
    EOP_PG = 0,     // proximal-gradient state machine (fista / fista_delta / ista)
    EOP_POWER = 1,  // power iteration step
    EOP_FG = 2,     // g (+a2 x), loss
    EOP_OBJ = 3,    // objective value
    EOP_ATB = 4,    // |A^T b|_inf  (lambda_max)
};

constexpr int FOS_EPI_CLUSTER = 8;    // CTAs in the epilogue cluster (portable maximum)
constexpr int FOS_EPI_THREADS = 256;  // threads per epilogue CTA
constexpr int FOS_NSCAL = 8;          // scalars reduced per epilogue

// Control block, one per design, resident in device memory.  The host writes the
// configuration before a solve; afterwards only the epilogue kernel writes it and the
// gradient kernel reads g_mode.
struct FosCtrl {
    // ---- configuration
    int scheme, backtracking, adaptive_restart, want_hist, want_obj, max_iter, obj_terms;
    int pad0;
    double alpha1, alpha2, eta, tol, tol_ratio, restart_thr, delta, armijo_c;
    // ---- dynamic state
    int g_mode, phase, k, shrinks, obj_pending, stop_reason, n_grad_calls, n_passes;
    double tau, t_mom, prev_step, trial_t, gy, gd, cand_xx, pend_l2, pend_l1;
    unsigned long long pass_t0;  // %globaltimer at the start of the current pass
    unsigned long long epi_ns, xchg_ns;  // accumulated epilogue time / time spent waiting for peers
    unsigned long long grad_ns;          // persistent solve kernel: accumulated time from the start of a pass until
                                         // every CTA has published its partials (the gradient phase proper)
    int use_qrec, pad_q;                 // the recorded objective comes from the residual recurrence (GM_QREC)
    double beta_y;                       // momentum coefficient the current y was formed with (0: y == x)
    // ---- power iteration
    double L, L_prev, ptol;
    int pit, pit_max;
    // ---- one-shot outputs
    double out[4];
};

// Per-solve device arrays referenced by the epilogue.
struct FosHist {
    double* x_hist;     // (max_iter+1) x d   (may be null)
    double* obj_hist;   // max_iter
    double* t_hist;     // max_iter+1
    double* step_hist;  // max_iter
    int* ls_iters;      // max_iter
    float* grad_ms;     // max_iter+1
    float* ls_ms;       // max_iter
};

// Peer-memory exchange window (multi-GPU).  Every rank owns one window in its own HBM,
// mapped into all peers through CUDA IPC:  [2 slots][ (ldv + FOS_NSCAL) doubles ] payload,
// followed by [2 slots][world] 64-bit arrival flags.
struct FosPeer {
    double* win[8];                // win[r] = rank r's window as mapped in this process
    unsigned long long* flag[8];   // flag[r] = rank r's arrival counters, one per source rank
    unsigned long long* epoch;     // device-resident exchange counter (identical on all ranks)
    // second region of the same window, used by the persistent solve kernel (push model): every rank
    // WRITES its column slices into all windows, [2 slots][FOS_MAX_WORLD source ranks][ldv + FOS_WIN_PAD],
    // and signals per (source rank, CTA): fflag[r][src * FOS_MAX_PARTS + cta] in rank r's window
    double* fwin[8];
    unsigned long long* fflag[8];
};
constexpr int FOS_WIN_PAD = 8;     // doubles after the ldv payload: [s1, s2, spare...]
constexpr int FOS_MAX_WORLD = 8;
constexpr int FOS_MAX_PARTS = 160; // >= CTAs of the streaming kernel (one per SM)

// Grid-wide synchronisation state of the persistent solve kernel (device memory, zeroed once when the
// design is created, never reset: all counters are monotonic).
struct FosGridSync {
    unsigned long long arrive0, pad0[15];  // one 128-byte line per counter
    unsigned long long arrive1, pad1[15];
    unsigned long long arrive2, pad2[15];
    unsigned long long gen, pad3[15];      // passes completed (written by the leader before the last barrier of a pass)
    int abort, pad4[31];                   // set when a wait timed out: every CTA leaves
    double scal[FOS_MAX_PARTS][FOS_NSCAL]; // per-CTA column-slice sums of the current pass
    // phase profile of CTA 0 (ns, accumulated over passes): [0] streaming loop, [1] wait at barrier 1,
    // [2] slice sums, [3] peer exchange, [4] elementwise 1 + slice scalars, [5] barrier 2, [6] scalar totals +
    // decision + elementwise 2 + commit, [7] barrier 3, [8] passes
    unsigned long long prof[16];
};

// Everything the gradient kernel needs.
struct GradArgs {
    const void* A;
    const double* b;
    const double* v1;
    const double* v2;
    double* partial_g;   // [n_parts][ldv]
    double* partial_s;   // [n_parts][2]
    FosCtrl* ctrl;
    long long n;
    int d, lda, ldv;
    int mode_override;   // >= 0: use instead of ctrl->g_mode
    const long long* row_lo;        // [n_parts + 1] row partition of the streaming kernel
    const int* sm_slot;             // [256] SM id -> preferred partition slot (nullable: slot = blockIdx.x)
    unsigned* slot_claim;           // [n_parts] pass number that last claimed each slot (with sm_slot)
    unsigned pass_no;               // number of this launch (monotonic per design): the claim token
    unsigned long long* cta_times;  // debug: [n_parts][2] start/end %globaltimer per CTA (nullable)
    double* qres;                   // [2][n] residual A x_k - b of the iterate, row by row; the two copies alternate (GM_QREC; nullable)
};

// Everything the epilogue kernel needs.
struct EpiArgs {
    FosCtrl* ctrl;
    FosHist hist;
    const double* partial_g;
    const double* partial_s;
    int n_parts;
    int d, ldv;
    int op;
    int g_mode_ran;      // for one-shot ops: the mode the preceding pass ran with
    double* y;           // v1 of the next pass
    double* xc;          // v2 of the next pass (candidate / newest iterate)
    double* xk;          // current iterate
    double* g;           // reduced gradient
    double op_a1, op_a2; // scalars of the one-shot ops
    int op_bits;
    // multi-GPU
    int world, rank;
    FosPeer peer;
};

// ------------------------------------------------------------------------------------------
// design handle (host side)
// ------------------------------------------------------------------------------------------
struct fos_design {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    long long n = 0;
    int d = 0, lda = 0, ldv = 0;
    int dtype = FOS_F64;
    void* A = nullptr;
    size_t A_cap = 0;  // bytes of the block A sits in (>= the matrix: blocks are recycled, fos_api.cu)
    double* b = nullptr;
    bool owns_A = false, owns_b = false;
    // workspaces (carved out of work_block / pin_block)
    void* work_block = nullptr;
    void* pin_block = nullptr;
    int n_parts = 0;
    double* partial_g = nullptr;
    double* partial_s = nullptr;
    double *y = nullptr, *xc = nullptr, *xk = nullptr, *g = nullptr;
    FosCtrl* ctrl = nullptr;
    FosCtrl* ctrl_host = nullptr;  // pinned mirror
    double* vec_host = nullptr;    // pinned staging, 2*ldv + 8 doubles
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::atomic<long long> launches{0};  // also bumped by the upload thread of a two-step creation
    // multi-GPU
    int world = 1, rank = 0;
    void* window = nullptr;        // this rank's exchange window (cudaMalloc, IPC-exported)
    size_t window_bytes = 0;
    void* peer_base[8] = {nullptr};  // cudaIpcOpenMemHandle results (to close on destroy)
    // windows shared through the virtual-memory-management API instead (comm_vmm.cu)
    bool vmm = false;
    unsigned long long vmm_handle[8] = {0};  // CUmemGenericAllocationHandle per rank (own + imported)
    void* vmm_ptr[8] = {nullptr};            // where each rank's window is mapped in this process
    size_t vmm_size = 0, vmm_gran = 0;
    int vmm_fd = -1;                         // exported descriptor of the own window
    FosPeer peer{};
    // gradient kernel selection (chosen at creation)
    int kern_kind = 0;  // 0 generic, 1 streaming
    FosGridSync* gsync = nullptr;  // persistent solve kernel (streaming designs only)
    double* qres = nullptr;        // [2][n] residual vectors of the recurrence (allocated by the first solve that uses it)
    bool fused_ok = false;         // the whole solve may run as ONE launch of the persistent kernel
    bool lite_ok = false;         // a gradient-only kernel variant exists for this shape
    bool grad_only_hint = false;  // the running loop never asks for the second dot
    unsigned long long* cta_times = nullptr;  // debug buffer (fos_debug_cta_times)
    long long* row_lo = nullptr;              // device: [n_parts + 1] row partition (streaming kernel)
    std::vector<long long> row_lo_host;
    int* sm_slot = nullptr;                   // device: SM id -> slot, set when the partition is SM-indexed;
                                              // followed by n_parts claim words (GradArgs::slot_claim)
    bool balanced = false;
    // two-step creation: the exchange windows may be attached (another host thread) while the upload is
    // still running; the rate-weighted row partition must not be calibrated under a PCIe copy, so it is
    // deferred to the end of the upload
    std::mutex* life_mu = nullptr;
    bool uploading = false, balance_pending = false;
    bool pdl = true;  // launch passes with programmatic dependent launch (FOS_NO_PDL=1 disables)
    void* arena = nullptr;         // grow-only device workspace of the per-solve arrays (fos_arena_reserve)
    size_t arena_bytes = 0;
    void* pin_scratch = nullptr;   // FOS_PIN_SCRATCH bytes of pinned host memory for control-block snapshots
    // Gram matrix of the local rows accumulated under the host->device upload (gram_kernels.cu)
    double* G_up = nullptr;   // [d][d], owned
    int G_state = 0;          // 0 none, 1 local rows, 2 summed over all ranks by the caller
    double* up_W = nullptr;   // split workspace, alive only during the upload
    size_t up_W_bytes = 0;
    int up_nsplit = 0;
    float up_copy_ms = 0.f, up_tail_ms = 0.f;  // upload duration / Gram work left after the last byte arrived
    // optional per-launch event timing of the gradient kernel
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
    size_t prof_used = 0;
};

constexpr size_t FOS_PIN_SCRATCH = 16384;
// Recycling allocator for the per-design / per-call blocks (fos_api.cu).  cudaMalloc, cudaFree,
// cudaMallocHost and cudaFreeHost are trips through the driver's global lock and sporadically take
// 50-100 ms each on a busy host (measured around a 0.72 s upload: 'workspaces 94 ms', 'cudaFree(b)
// 95 ms'); a drop-in caller creates and destroys a design per solver call, so freed blocks are kept
// (per device, device and pinned memory apart, bounded) and handed to the next request of a similar
// size.  fos_trim() releases them.
cudaError_t fos_pool_malloc(void** p, size_t bytes);      // device memory of the current device
cudaError_t fos_pool_malloc_host(void** p, size_t bytes); // pinned host memory
void fos_pool_free(void* p);                              // either kind; nullptr is fine
void fos_pool_trim();
// Grow-only per-design device workspace: solver calls carve their per-call arrays out of it
// instead of cudaMalloc/cudaFree pairs (milliseconds each once peer mappings exist).  The
// previous contents are lost when it grows; one solver call at a time per design.
int fos_arena_reserve(fos_design* h, size_t bytes, void** base);

// launchers implemented in the .cu files
cudaError_t fos_launch_ex(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t s, void** args,
                          bool pdl, int cluster_x);
int fos_launch_grad(fos_design* h, int mode_override);
// persistent solve kernel: runs passes (gradient + in-kernel epilogue + peer exchange) until the state
// machine reports PH_DONE or max_passes have run.  FOS_ERR_UNSUPPORTED when the design does not qualify.
int fos_launch_solve(fos_design* h, const FosHist& hist, long long max_passes);
int fos_solve_stages(const fos_design* h);  // ring depth the persistent kernel would use; 0 = not eligible
size_t fos_window_bytes(const fos_design* h);              // exchange window: both regions + flags
void fos_window_bind(fos_design* h, int r, void* base);    // point peer.{win,flag,fwin,fflag}[r] into a mapped window
int fos_launch_epilogue(fos_design* h, int op, int g_mode_ran, const FosHist& hist, double a1,
                        double a2, int bits);
int fos_grad_plan(fos_design* h);
void fos_comm_vmm_release(fos_design* h);
void fos_block_cache_trim();  // releases the recycled matrix blocks (part of fos_trim)
int fos_comm_after_attach(fos_design* h);  // common tail of the attach variants (row-balance policy)
int fos_balance_rows(fos_design* h);  // SM-indexed row partition weighted by measured per-SM rates  // picks kernel + n_parts, sets smem attributes
int fos_launch_prox(const double* v_dev, double* out_dev, long long len, double thresh,
                    double scale, cudaStream_t stream);
int fos_launch_synthetic(fos_design* h, unsigned long long seed, double noise, double rho1,
                         double rho2, long long row0);
// upload-overlapped Gram accumulation and the power iteration on it (gram_kernels.cu)
bool fos_upload_gram_eligible(const fos_design* h);
long long fos_upload_gram_chunk_rows(const fos_design* h);
int fos_upload_gram_begin(fos_design* h, cudaStream_t s);
int fos_upload_gram_chunk(fos_design* h, long long row0, long long rows, cudaStream_t s);
int fos_upload_gram_finish(fos_design* h, cudaStream_t s);
void fos_upload_gram_drop(fos_design* h);
int fos_gram_power_iter(fos_design* h, const double* v0, int n_iter, double tol, double* L_out, int* iters_out,
                        float* gpu_ms_out);
int fos_launch_repack(const void* src_dev, void* dst_dev, long long rows, int d, int lda,
                      long long row_stride, long long col_stride, int dtype, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch (PDL): a kernel launched with the programmatic stream
// serialization attribute may start while its predecessor is still running; everything after
// fos_pdl_wait() sees the predecessor's completed memory.  No-ops for ordinary launches.
__device__ __forceinline__ void fos_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fos_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned long long fos_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// numpy semantics of sign(v)*maximum(|v|-thr, 0)  (prox_operators.py:8): sign(+-0) = +0,
// NaN propagates through both factors, a shrunk negative entry becomes -0.0.
__device__ __forceinline__ double fos_soft_threshold(double v, double thr) {
    double sgn = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : ((v == 0.0) ? 0.0 : v));
    double mag = __dsub_rn(fabs(v), thr);
    double mx = (mag != mag) ? mag : ((mag > 0.0) ? mag : 0.0);
    return __dmul_rn(sgn, mx);
}

__device__ __forceinline__ double fos_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
