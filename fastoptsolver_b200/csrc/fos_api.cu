// fos_api.cu -- C ABI of libfos_b200.so (see include/fos.h): design handles, one-shot
// operators, the power iteration and the proximal-gradient driver.  The host only launches
// passes and polls a pinned copy of the control block; every numerical decision is taken on
// the device (epilogue_kernels.cu).
#include <stdarg.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "fos_common.cuh"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void fos_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fos_last_error(void) { return g_err; }

cudaError_t fos_launch_ex(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t s, void** args,
                          bool pdl, int cluster_x) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attrs[2];
    unsigned n = 0;
    if (pdl) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x < 0) {  // cooperative launch (grid-wide barriers inside the kernel)
        attrs[n].id = cudaLaunchAttributeCooperative;
        attrs[n].val.cooperative = 1;
        ++n;
    }
    if (cluster_x > 0) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = cluster_x;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    return cudaLaunchKernelExC(&cfg, fn, args);
}
extern "C" int fos_abi_version(void) { return FOS_ABI_VERSION; }

extern "C" int fos_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int fos_device_info(int device, int* sm_count, size_t* total_bytes, size_t* free_bytes) {
    FOS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FOS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    size_t f = 0, t = 0;
    FOS_CUDA(cudaMemGetInfo(&f, &t));
    if (total_bytes) *total_bytes = t;
    if (free_bytes) *free_bytes = f;
    return FOS_OK;
}

// ------------------------------------------------------------------------------------------
// design lifetime
// ------------------------------------------------------------------------------------------
static inline int elem_size(int dtype) { return dtype == FOS_F64 ? 8 : 4; }


// ------------------------------------------------------------------------------------------
// recycling allocator (see fos_common.cuh)
// ------------------------------------------------------------------------------------------
namespace {
struct PoolBlock {
    void* p;
    size_t cap;
    int device;  // -1: pinned host memory
};
std::mutex g_pool_mu;
std::map<void*, PoolBlock> g_pool_live;   // blocks handed out
std::vector<PoolBlock> g_pool_idle;       // blocks waiting for reuse (oldest first)
constexpr size_t POOL_IDLE_DEVICE_MAX = 1536u << 20;  // per process, all devices
constexpr size_t POOL_IDLE_PINNED_MAX = 96u << 20;

bool pool_on() {
    static const bool on = [] {
        const char* e = getenv("FOS_POOL");
        return !(e && e[0] == '0');
    }();
    return on;
}
size_t pool_round(size_t bytes) {
    if (bytes < 4096) return 4096;
    if (bytes < (1u << 20)) return (bytes + 4095) & ~static_cast<size_t>(4095);
    return (bytes + (1u << 20) - 1) & ~static_cast<size_t>((1u << 20) - 1);
}
void pool_release(const PoolBlock& b) {
    if (b.device < 0) {
        cudaFreeHost(b.p);
    } else {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != b.device) cudaSetDevice(b.device);
        cudaFree(b.p);
        if (cur != b.device) cudaSetDevice(cur);
    }
}
// evict idle blocks (oldest first) of one kind until at most `keep` bytes of it stay; lock held
void pool_shrink(bool pinned, size_t keep) {
    size_t total = 0;
    for (const PoolBlock& b : g_pool_idle)
        if ((b.device < 0) == pinned) total += b.cap;
    for (size_t i = 0; i < g_pool_idle.size() && total > keep;) {
        if ((g_pool_idle[i].device < 0) == pinned) {
            total -= g_pool_idle[i].cap;
            pool_release(g_pool_idle[i]);
            g_pool_idle.erase(g_pool_idle.begin() + i);
        } else {
            ++i;
        }
    }
}
cudaError_t pool_get(void** out, size_t bytes, bool pinned) {
    *out = nullptr;
    if (bytes == 0) bytes = 1;
    int device = -1;
    if (!pinned) {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return e;
    }
    const size_t want = pool_round(bytes);
    std::lock_guard<std::mutex> lock(g_pool_mu);
    // best fit among the idle blocks of this kind: never hand out more than twice the request
    int best = -1;
    for (size_t i = 0; i < g_pool_idle.size(); ++i) {
        const PoolBlock& b = g_pool_idle[i];
        if (b.device != device || b.cap < want || b.cap > 2 * want) continue;
        if (best < 0 || b.cap < g_pool_idle[best].cap) best = static_cast<int>(i);
    }
    PoolBlock blk;
    if (best >= 0) {
        blk = g_pool_idle[best];
        g_pool_idle.erase(g_pool_idle.begin() + best);
    } else {
        void* p = nullptr;
        cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e == cudaErrorMemoryAllocation) {  // make room: give the idle blocks back and try once more
            cudaGetLastError();
            pool_shrink(pinned, 0);
            e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) return e;
        blk = PoolBlock{p, want, device};
    }
    g_pool_live[blk.p] = blk;
    *out = blk.p;
    return cudaSuccess;
}
}  // namespace

cudaError_t fos_pool_malloc(void** p, size_t bytes) { return pool_get(p, bytes, false); }
cudaError_t fos_pool_malloc_host(void** p, size_t bytes) { return pool_get(p, bytes, true); }

void fos_pool_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = g_pool_live.find(p);
    if (it == g_pool_live.end()) {  // not ours (never happens for pool allocations): plain free
        cudaFree(p);
        cudaGetLastError();
        return;
    }
    const PoolBlock blk = it->second;
    g_pool_live.erase(it);
    const bool pinned = blk.device < 0;
    const size_t cap_kind = pinned ? POOL_IDLE_PINNED_MAX : POOL_IDLE_DEVICE_MAX;
    if (!pool_on() || blk.cap > cap_kind) {
        pool_release(blk);
        return;
    }
    g_pool_idle.push_back(blk);
    pool_shrink(pinned, cap_kind);
}

void fos_pool_trim() {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (const PoolBlock& b : g_pool_idle) pool_release(b);
    g_pool_idle.clear();
    cudaGetLastError();
}

static int design_common_init(fos_design* h, long long n, long long d, int dtype, int device) {
    FOS_REQUIRE(n >= 1 && d >= 1, "design must have at least one row and one column (got %lld x %lld)", n, d);
    FOS_REQUIRE(dtype == FOS_F64 || dtype == FOS_F32, "dtype must be FOS_F64 or FOS_F32");
    if (d > 8192) {
        fos_set_error("d = %lld exceeds the supported maximum of 8192 columns", d);
        return FOS_ERR_UNSUPPORTED;
    }
    int ndev = fos_device_count();
    if (ndev <= 0) {
        fos_set_error("no CUDA device visible: libfos_b200 has no CPU fallback");
        return FOS_ERR_CUDA;
    }
    FOS_REQUIRE(device >= 0 && device < ndev, "device %d out of range (0..%d)", device, ndev - 1);
    FOS_CUDA(cudaSetDevice(device));
    // two attribute queries, cached per device: cudaGetDeviceProperties fills ~100 fields (some through
    // the driver's global lock) and took 12-128 ms per design on a busy host
    static std::mutex prop_mu;
    static std::map<int, std::pair<int, int>> prop_cache;  // device -> (cc major * 10 + minor, SM count)
    std::pair<int, int> pr;
    {
        std::lock_guard<std::mutex> lock(prop_mu);
        auto it = prop_cache.find(device);
        if (it == prop_cache.end()) {
            int major = 0, minor = 0, sms = 0;
            FOS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
            FOS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
            FOS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            it = prop_cache.emplace(device, std::make_pair(major * 10 + minor, sms)).first;
        }
        pr = it->second;
    }
    if (pr.first / 10 != 10) {
        fos_set_error("device %d is sm_%d; libfos_b200 contains sm_100a code only", device, pr.first);
        return FOS_ERR_UNSUPPORTED;
    }
    h->device = device;
    h->sm_count = pr.second;
    h->n = n;
    h->d = static_cast<int>(d);
    h->dtype = dtype;
    const int per16 = 16 / elem_size(dtype);
    h->lda = static_cast<int>((d + per16 - 1) / per16 * per16);
    h->ldv = static_cast<int>((d + 1) / 2 * 2);
    {
        const char* e = getenv("FOS_NO_PDL");
        h->pdl = !(e && e[0] == '1');
    }
    if (h->life_mu == nullptr) h->life_mu = new std::mutex();
    FOS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    FOS_CUDA(cudaEventCreate(&h->ev0));
    FOS_CUDA(cudaEventCreate(&h->ev1));
    return FOS_OK;
}

static int design_alloc_work(fos_design* h) {
    FOS_TRY(fos_grad_plan(h));
    // ONE device block and ONE pinned block for all per-design workspaces: every separate
    // cudaMalloc / cudaMallocHost is a trip through the driver's global lock (50 ms for the lot when
    // eight ranks create their designs at the same moment)
    const size_t vb = static_cast<size_t>(h->ldv) * sizeof(double);
    auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    const size_t o_pg = 0;
    const size_t o_ps = o_pg + up(static_cast<size_t>(h->n_parts) * vb);
    const size_t o_y = o_ps + up(static_cast<size_t>(h->n_parts) * 2 * sizeof(double));
    const size_t o_xc = o_y + up(vb), o_xk = o_xc + up(vb), o_g = o_xk + up(vb);
    const size_t o_row = o_g + up(vb);
    const size_t o_ctrl = o_row + up((h->n_parts + 1) * sizeof(long long));
    const size_t o_sync = o_ctrl + up(sizeof(FosCtrl));
    const size_t total = o_sync + (h->kern_kind == 1 ? up(sizeof(FosGridSync)) : 0);
    FOS_CUDA(fos_pool_malloc(&h->work_block, total));
    FOS_CUDA(cudaMemsetAsync(h->work_block, 0, total, h->stream));
    char* wb = static_cast<char*>(h->work_block);
    h->partial_g = reinterpret_cast<double*>(wb + o_pg);
    h->partial_s = reinterpret_cast<double*>(wb + o_ps);
    h->y = reinterpret_cast<double*>(wb + o_y);
    h->xc = reinterpret_cast<double*>(wb + o_xc);
    h->xk = reinterpret_cast<double*>(wb + o_xk);
    h->g = reinterpret_cast<double*>(wb + o_g);
    h->row_lo = reinterpret_cast<long long*>(wb + o_row);
    h->ctrl = reinterpret_cast<FosCtrl*>(wb + o_ctrl);
    h->gsync = (h->kern_kind == 1) ? reinterpret_cast<FosGridSync*>(wb + o_sync) : nullptr;  // zeroed with the block
    // row partition of the streaming kernel: equal blocks to start with
    h->row_lo_host.resize(h->n_parts + 1);
    for (int c = 0; c <= h->n_parts; ++c) h->row_lo_host[c] = (h->n * c) / h->n_parts;
    FOS_CUDA(cudaMemcpyAsync(h->row_lo, h->row_lo_host.data(), (h->n_parts + 1) * sizeof(long long),
                             cudaMemcpyHostToDevice, h->stream));
    const size_t p_ctrl = 0;
    const size_t p_vec = p_ctrl + up(4 * sizeof(FosCtrl));
    const size_t p_scr = p_vec + up((2 * static_cast<size_t>(h->ldv) + 16) * sizeof(double));
    FOS_CUDA(fos_pool_malloc_host(&h->pin_block, p_scr + FOS_PIN_SCRATCH));
    char* pb = static_cast<char*>(h->pin_block);
    h->ctrl_host = reinterpret_cast<FosCtrl*>(pb + p_ctrl);
    h->vec_host = reinterpret_cast<double*>(pb + p_vec);
    h->pin_scratch = pb + p_scr;
    memset(h->ctrl_host, 0, 4 * sizeof(FosCtrl));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    {
        // FOS_BALANCE=1 forces the rate-weighted partition on; by default it is switched on when
        // the design joins >= 4 ranks (fos_comm_attach), where the GPUs are not power-capped
        const char* e = getenv("FOS_BALANCE");
        if (e && e[0] == '1') FOS_TRY(fos_balance_rows(h));
    }
    return FOS_OK;
}

// ------------------------------------------------------------------------------------------
// SM-indexed, rate-weighted row partition
// ------------------------------------------------------------------------------------------
// With equal row blocks the CTAs of one pass finish up to 30 % apart (tools/exp_cta_balance.py):
// the SMs do not get equal shares of the memory system, and the shares are a stable property of
// the SM (correlation 0.95 between launches), not of the block index.  While all CTAs stream the
// total rate is the HBM ceiling, but once the fast SMs retire, the remaining ones cannot absorb
// the freed bandwidth (their consumer chain is latency bound), so the tail runs below the
// ceiling.  Fix: tie the row block to the SM a CTA lands on (one persistent CTA per SM) and size
// the blocks so that all SMs finish together.  The weights are measured once per process and
// device (a few passes on zero vectors) and reused for every design; the partition of a design
// never changes afterwards, so its results stay bit-reproducible for its lifetime.
// Measured (1M x 4096 fp64, B200): an isolated pass drops from 4.93 to 4.49 ms and the CTA finish
// spread from 30 % to 4 %, but 100 back-to-back FISTA steps run at 199 it/s instead of 203-209
// (1 GPU) and 386 instead of 392 (2 GPUs): the sustained loop sits at the 1 kW power cap, where
// keeping every SM busy to the end buys nothing.  With the rows spread over 8 GPUs each GPU draws
// ~380 W, is not capped, and the weighted partition wins: 1520 vs 1443 it/s (kernel 0.592 vs
// 0.640 ms).  Policy: on when a design joins >= 4 ranks, off otherwise; FOS_BALANCE=1/0 forces it.
// Without it the blocks are equal and indexed by blockIdx (bit-reproducible across processes).
static std::mutex g_bal_mutex;
static std::map<std::pair<int, int>, std::vector<double>> g_bal_weights;  // (device, n_parts) -> weights

static void apply_weights(fos_design* h, const std::vector<double>& w) {
    const int P = h->n_parts;
    double wsum = 0.0;
    for (double v : w) wsum += v;
    double cum = 0.0;
    h->row_lo_host[0] = 0;
    for (int c = 0; c < P; ++c) {
        cum += w[c];
        long long hi = static_cast<long long>(llround(static_cast<double>(h->n) * cum / wsum));
        hi = std::min(hi, static_cast<long long>(h->n));
        h->row_lo_host[c + 1] = std::max(hi, h->row_lo_host[c]);
    }
    h->row_lo_host[P] = h->n;
}

int fos_balance_rows(fos_design* h) {
    const int P = h->n_parts;
    if (h->kern_kind != 1 || P != h->sm_count || P > 256) return FOS_OK;
    if (h->n < 256LL * P) return FOS_OK;  // too few stages per CTA for the weights to matter
    // slot table: SM id -> preferred slot (identity when the SM ids are 0..sm_count-1), followed by
    // one claim word per slot.  The table is a preference only: the kernel claims its slot with an
    // atomic exchange of the launch's pass number and walks on if it is taken, so a pass covers
    // every row block exactly once whatever the placement of the CTAs (%smid is not guaranteed to
    // be contiguous, and other work may hold some SMs).
    std::vector<int> slot(256 + P, 0);
    for (int i = 0; i < 256; ++i) slot[i] = i % P;
    for (int i = 0; i < P; ++i) slot[256 + i] = -1;  // 0xFFFFFFFF: never a pass number (31 bits)
    if (h->sm_slot == nullptr) FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&h->sm_slot), slot.size() * sizeof(int)));
    FOS_CUDA(cudaMemcpy(h->sm_slot, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice));

    std::lock_guard<std::mutex> lock(g_bal_mutex);
    const auto key = std::make_pair(h->device, P);
    auto it = g_bal_weights.find(key);
    if (it == g_bal_weights.end()) {
        // calibrate on this design
        unsigned long long* buf = nullptr;
        FOS_CUDA(cudaMalloc(&buf, 2 * static_cast<size_t>(P) * sizeof(unsigned long long)));
        std::vector<unsigned long long> t(2 * P);
        std::vector<double> w(P, 1.0), best_w(P, 1.0), dur(P);
        double best_T = 1e300;
        int status = FOS_OK;
        bool usable = true;
        for (int round = 0; round < 6 && status == FOS_OK; ++round) {
            apply_weights(h, w);
            cudaMemcpy(h->row_lo, h->row_lo_host.data(), (P + 1) * sizeof(long long), cudaMemcpyHostToDevice);
            status = fos_launch_grad(h, GM_GRAD | GM_DOT2);
            cudaMemsetAsync(buf, 0, t.size() * sizeof(unsigned long long), h->stream);
            h->cta_times = buf;
            if (status == FOS_OK) status = fos_launch_grad(h, GM_GRAD | GM_DOT2);
            h->cta_times = nullptr;
            if (status != FOS_OK) break;
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e == cudaSuccess)
                e = cudaMemcpy(t.data(), buf, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                fos_set_error("row-balance calibration failed: %s", cudaGetErrorString(e));
                status = FOS_ERR_CUDA;
                break;
            }
            // every slot must have been run exactly once, by a CTA sitting on the SM the table maps
            // to it; otherwise the SM-indexed weights mean nothing on this device / in this process
            // state and the design keeps its equal, blockIdx-indexed blocks
            for (int c = 0; c < P && usable; ++c) {
                const unsigned long long smid = t[2 * c + 1] >> 48;
                if (t[2 * c] == 0 || t[2 * c + 1] == 0 || static_cast<int>(smid % P) != c) usable = false;
            }
            if (!usable) break;
            unsigned long long t0 = ~0ull, t1 = 0;
            double mean = 0.0;
            for (int c = 0; c < P; ++c) {
                const unsigned long long st = t[2 * c] & 0xFFFFFFFFFFFFull, en = t[2 * c + 1] & 0xFFFFFFFFFFFFull;
                t0 = std::min(t0, st);
                t1 = std::max(t1, en);
                dur[c] = static_cast<double>(en - st);
                mean += dur[c] / P;
            }
            const double T = static_cast<double>(t1 - t0);
            if (T < best_T) {
                best_T = T;
                best_w = w;
            }
            if (round == 5) break;
            double ws = 0.0;
            for (int c = 0; c < P; ++c) {
                const double f = (dur[c] > 0.0) ? mean / dur[c] : 1.0;
                w[c] *= pow(f, 0.8);
                ws += w[c] / P;
            }
            for (int c = 0; c < P; ++c) w[c] = std::min(1.6, std::max(0.5, w[c] / ws));
        }
        cudaFree(buf);
        cudaMemsetAsync(h->partial_g, 0, static_cast<size_t>(P) * h->ldv * sizeof(double), h->stream);
        cudaMemsetAsync(h->partial_s, 0, static_cast<size_t>(P) * 2 * sizeof(double), h->stream);
        cudaStreamSynchronize(h->stream);
        if (status != FOS_OK) return status;
        if (!usable) best_w.clear();  // remembered: this device does not get SM-indexed blocks
        it = g_bal_weights.emplace(key, best_w).first;
    }
    if (it->second.empty()) {
        // equal blocks indexed by blockIdx (the default partition)
        fos_pool_free(h->sm_slot);
        h->sm_slot = nullptr;
        for (int c = 0; c <= P; ++c) h->row_lo_host[c] = (h->n * c) / P;
        FOS_CUDA(cudaMemcpy(h->row_lo, h->row_lo_host.data(), (P + 1) * sizeof(long long), cudaMemcpyHostToDevice));
        return FOS_OK;
    }
    apply_weights(h, it->second);
    FOS_CUDA(cudaMemcpy(h->row_lo, h->row_lo_host.data(), (P + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    h->balanced = true;
    return FOS_OK;
}


static void blk_give(int device, void* p, size_t cap);  // matrix-block recycling, below

static void design_free(fos_design* h) {
    if (!h) return;
    const bool dbg = getenv("FOS_UPLOAD_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto f0 = now();
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    const auto f1 = now();
    if (h->owns_A && h->A) blk_give(h->device, h->A, h->A_cap);
    if (h->owns_b && h->b) fos_pool_free(h->b);
    const auto f2 = now();
    fos_upload_gram_drop(h);
    const auto f3 = now();
    void* bufs[] = {h->work_block, h->sm_slot, h->arena, h->qres};
    for (void* p : bufs) fos_pool_free(p);
    const auto f4 = now();
    fos_pool_free(h->pin_block);
    if (dbg)
        fprintf(stderr, "[fos] design_free: sync %.1f ms, A+b %.1f ms, gram %.1f ms, work/arena %.1f ms, pinned %.1f ms\n",
                ms(f0, f1), ms(f1, f2), ms(f2, f3), ms(f3, f4), ms(f4, now()));
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    for (int r = 0; r < FOS_MAX_WORLD; ++r)
        if (h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    if (h->vmm)
        fos_comm_vmm_release(h);
    else if (h->window)
        cudaFree(h->window);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h->life_mu;
    delete h;
}

#define FOS_TRY_FREE(h, expr)   \
    do {                        \
        int _s = (expr);        \
        if (_s != FOS_OK) {     \
            design_free(h);     \
            return _s;          \
        }                       \
    } while (0)

// Plain cudaMalloc on purpose.  Once a process has peer access enabled (the exchange windows of a
// row-sharded design) a FRESH block of this size is also mapped into the peers (measured 200-260 ms
// for an 8 GB shard with 3 peers, once: the driver then reuses the freed block in 2-25 ms).  The
// stream-ordered pool (cudaMallocAsync) avoids the peer mapping but pays 300-540 ms for the same
// 8 GB on EVERY allocation (tools/exp_e2e_multi.py), so it is not used.
// The reference re-reads its arrays on every call, so a drop-in caller uploads the same design again
// and again (19 solver variants per scenario in the notebook).  A fresh cudaMalloc of a large block
// and the cudaFree that follows cost far more than they look (page tables for tens of GB: measured
// ~0.2 s of a 0.8 s upload of 32.8 GB), so the matrix block of a destroyed design is kept -- one
// block per device, >= 256 MB -- and handed to the next design that fits in it.  fos_trim() releases
// it; FOS_KEEP_BLOCK=0 switches the cache off.
static std::mutex g_blk_mu;
static std::map<int, std::pair<void*, size_t>> g_blk_cache;  // device -> (block, bytes), not in use
constexpr size_t FOS_BLOCK_CACHE_MIN = 256u << 20;

static bool blk_cache_on() {
    const char* e = getenv("FOS_KEEP_BLOCK");
    return !(e && e[0] == '0');
}

static void* blk_take(int device, size_t bytes, size_t* cap) {
    std::lock_guard<std::mutex> lock(g_blk_mu);
    auto it = g_blk_cache.find(device);
    if (it == g_blk_cache.end()) return nullptr;
    void* p = it->second.first;
    const size_t have = it->second.second;
    g_blk_cache.erase(it);
    if (have >= bytes) {
        *cap = have;
        return p;
    }
    cudaFree(p);  // too small for the newcomer: make room before the fresh allocation
    return nullptr;
}

static void blk_give(int device, void* p, size_t cap) {
    if (!p) return;
    if (!blk_cache_on() || cap < FOS_BLOCK_CACHE_MIN) {
        cudaFree(p);
        return;
    }
    std::lock_guard<std::mutex> lock(g_blk_mu);
    auto it = g_blk_cache.find(device);
    if (it != g_blk_cache.end()) {
        if (it->second.second >= cap) {  // keep the larger one
            cudaFree(p);
            return;
        }
        cudaFree(it->second.first);
        g_blk_cache.erase(it);
    }
    g_blk_cache[device] = std::make_pair(p, cap);
}

void fos_block_cache_trim() {
    fos_pool_trim();
    std::lock_guard<std::mutex> lock(g_blk_mu);
    for (auto& kv : g_blk_cache) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.first);
    }
    g_blk_cache.clear();
    cudaGetLastError();
}

static int alloc_matrix(fos_design* h) {
    const size_t bytes = static_cast<size_t>(h->n) * h->lda * elem_size(h->dtype);
    h->A_cap = bytes;
    h->A = blk_take(h->device, bytes, &h->A_cap);
    cudaError_t e = h->A ? cudaSuccess : cudaMalloc(&h->A, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fos_set_error("cannot allocate %.2f GB of HBM for A (%lld x %d)", bytes / 1e9, h->n, h->d);
        return FOS_ERR_NOMEM;
    }
    h->owns_A = true;
    FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&h->b), static_cast<size_t>(h->n) * sizeof(double)));
    h->owns_b = true;
    return FOS_OK;
}

// ------------------------------------------------------------------------------------------
// host -> device copies of the design
// ------------------------------------------------------------------------------------------
// The reference's callers hand over plain numpy arrays, i.e. PAGEABLE memory, which
// cudaMemcpy moves through a single driver-side staging buffer at a fraction of the PCIe
// rate.  Large pageable sources are therefore copied by HostStager: T threads each memcpy their
// share of the pieces into their own pair of pinned slots and push them to the device on
// their own stream, so host memcpy and PCIe transfers of different pieces overlap.  Pinned
// (cudaHostAlloc / cudaHostRegister) sources go straight through cudaMemcpyAsync.
namespace {

struct HostStager {
    static constexpr int T = 16;                // most copy threads (FOS_UPLOAD_THREADS; default: hardware threads, 8..16)
    static constexpr int NSLOT = 2;             // pinned slots per thread (double buffer)
    static constexpr size_t SLOT = 8u << 20;    // bytes per slot
    static constexpr size_t MIN_BYTES = 128u << 20;  // below this a plain cudaMemcpyAsync is used

    // default: one copy thread per hardware thread the process may use, between 8 and 16 (measured on a 16-core
    // host, 32.8 GB pageable source: 8 threads 41 GB/s, 12: 43, 16: 45)
    static int default_threads() {
        unsigned hc = std::thread::hardware_concurrency();
        // several ranks on one host (torchrun exports LOCAL_WORLD_SIZE) share the cores
        if (const char* lw = getenv("LOCAL_WORLD_SIZE")) hc /= static_cast<unsigned>(std::max(1, atoi(lw)));
        return std::max(8, std::min(T, hc ? static_cast<int>(hc) : 8));
    }
    static int slot_threads() {
        int n = default_threads();
        if (const char* e = getenv("FOS_UPLOAD_THREADS")) n = std::max(8, std::min(T, atoi(e)));
        return n;
    }
    // process-wide pinned slots (8 threads x 2 x 8 MB = 128 MB unless more threads are asked for at
    // first use), allocated on first use and kept: pinning them costs ~70 ms
    static void* slots(int t, int s) {
        static std::mutex mu;
        static void* base = nullptr;
        static bool tried = false;
        std::lock_guard<std::mutex> lock(mu);
        if (!tried) {
            tried = true;
            if (cudaMallocHost(&base, static_cast<size_t>(slot_threads()) * NSLOT * SLOT) != cudaSuccess) {
                cudaGetLastError();
                base = nullptr;
            }
        }
        if (base && n_pinned_ref() == 0) n_pinned_ref() = slot_threads();
        return base ? static_cast<char*>(base) + (static_cast<size_t>(t) * NSLOT + s) * SLOT : nullptr;
    }
    static int& n_pinned_ref() {
        static int n = 0;
        return n;
    }
    static int pinned_threads() { return n_pinned_ref(); }

    // The pinned slots are shared by the whole process: a stager that has started staging owns
    // them (this lock) until its last device copy has left them (destructor, after drain()).
    static std::mutex& slot_owner() {
        static std::mutex mu;
        return mu;
    }
    std::unique_lock<std::mutex> own;

    int device = 0;
    int threads = default_threads();
    cudaStream_t stream[T] = {};
    cudaEvent_t slot_ev[T][NSLOT] = {};
    cudaEvent_t done_ev[T] = {};
    int next_slot[T] = {};
    bool ready = false, failed = false;

    int init(int dev) {
        device = dev;
        return FOS_OK;
    }
    // streams, events and the pinned slots are created the first time a pageable source shows up
    // (pinning the slots costs ~0.5 ms per MB: page-locked sources must not pay for it)
    int prepare() {
        if (ready || failed) return FOS_OK;
        failed = true;
        if (slots(0, 0) == nullptr) return FOS_OK;  // no pinned memory: stay on the plain path
        if (const char* e = getenv("FOS_UPLOAD_THREADS")) threads = std::max(1, std::min(T, atoi(e)));
        threads = std::min(threads, pinned_threads());
        for (int t = 0; t < threads; ++t) {
            FOS_CUDA(cudaStreamCreateWithFlags(&stream[t], cudaStreamNonBlocking));
            FOS_CUDA(cudaEventCreateWithFlags(&done_ev[t], cudaEventDisableTiming));
            for (int q = 0; q < NSLOT; ++q) FOS_CUDA(cudaEventCreateWithFlags(&slot_ev[t][q], cudaEventDisableTiming));
        }
        ready = true;
        failed = false;
        return FOS_OK;
    }
    void drain() {
        for (int t = 0; t < T; ++t)
            if (stream[t]) cudaStreamSynchronize(stream[t]);
    }
    ~HostStager() {
        drain();
        for (int t = 0; t < T; ++t) {
            if (stream[t]) cudaStreamDestroy(stream[t]);
            if (done_ev[t]) cudaEventDestroy(done_ev[t]);
            for (int q = 0; q < NSLOT; ++q)
                if (slot_ev[t][q]) cudaEventDestroy(slot_ev[t][q]);
        }
    }

    static bool pageable(const void* p) {
        const char* e = getenv("FOS_UPLOAD_STAGED");  // "0": never stage, "1": stage even pinned sources (tests)
        if (e && e[0] == '1') return true;
        if (e && e[0] == '0') return false;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return true;
        }
        return at.type == cudaMemoryTypeUnregistered;
    }

    // Equivalent of cudaMemcpyAsync(dst, src, bytes, HostToDevice, order_on): when it returns the
    // source has been read completely and `order_on` is ordered after the last device copy.
    int copy(void* dst, const void* src, size_t bytes, cudaStream_t order_on) {
        const char* e = getenv("FOS_UPLOAD_STAGED");
        const bool force = e && e[0] == '1';
        bool staged = (force || bytes >= MIN_BYTES) && pageable(src);
        if (staged) {
            FOS_TRY(prepare());
            staged = ready;
        }
        if (!staged) {
            FOS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, order_on));
            return FOS_OK;
        }
        if (!own.owns_lock()) own = std::unique_lock<std::mutex>(slot_owner());
        const size_t pieces = (bytes + SLOT - 1) / SLOT;
        std::atomic<int> err{static_cast<int>(cudaSuccess)};
        auto work = [&](int t) {
            cudaSetDevice(device);
            for (size_t p = t; p < pieces && err.load() == cudaSuccess; p += threads) {
                const int q = next_slot[t];
                next_slot[t] ^= 1;
                cudaError_t ce = cudaEventSynchronize(slot_ev[t][q]);  // the slot's previous copy has left
                const size_t off = p * SLOT, len = std::min(SLOT, bytes - off);
                void* st = slots(t, q);
                memcpy(st, static_cast<const char*>(src) + off, len);
                if (ce == cudaSuccess)
                    ce = cudaMemcpyAsync(static_cast<char*>(dst) + off, st, len, cudaMemcpyHostToDevice, stream[t]);
                if (ce == cudaSuccess) ce = cudaEventRecord(slot_ev[t][q], stream[t]);
                if (ce != cudaSuccess) err.store(static_cast<int>(ce));
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        if (err.load() != cudaSuccess) {
            fos_set_error("staged upload failed: %s", cudaGetErrorString(static_cast<cudaError_t>(err.load())));
            return FOS_ERR_CUDA;
        }
        for (int t = 0; t < threads; ++t) {
            FOS_CUDA(cudaEventRecord(done_ev[t], stream[t]));
            FOS_CUDA(cudaStreamWaitEvent(order_on, done_ev[t], 0));
        }
        return FOS_OK;
    }
};

}  // namespace

// Chunked H2D copy of a dense float64 C-order matrix with the Gram accumulation of the arrived
// rows running underneath on a second stream.  Copy stream: h->stream.
static int upload_dense(fos_design* h, const void* A, HostStager& stager, bool with_gram) {
    cudaStream_t cs = nullptr;
    FOS_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> evs;
    cudaEvent_t t_done = nullptr;
    const bool dbg = getenv("FOS_UPLOAD_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto body = [&]() -> int {
        const auto u0 = now();
        if (with_gram) FOS_TRY(fos_upload_gram_begin(h, cs));
        const auto u1 = now();
        const size_t row_bytes = static_cast<size_t>(h->lda) * (h->dtype == FOS_F64 ? 8 : 4);
        // 512 MB row chunks in every case: one giant cudaMemcpyAsync pays its whole DMA set-up
        // before the first byte moves (~70 ms for 8 GB), chunks pipeline it
        const long long chunk = with_gram ? fos_upload_gram_chunk_rows(h)
                                          : std::max<long long>(1, (512LL << 20) / static_cast<long long>(row_bytes));
        FOS_CUDA(cudaEventCreate(&t_done));
        FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
        for (long long r0 = 0, rows = 0; r0 < h->n; r0 += rows) {
            // the last 512 MB go in quarters: the Gram work left after the final byte is one small chunk
            const long long left = h->n - r0;
            rows = (with_gram && left <= chunk) ? std::min(left, std::max<long long>(chunk / 4, 1024)) : std::min(chunk, left);
            FOS_TRY(stager.copy(static_cast<char*>(h->A) + static_cast<size_t>(r0) * row_bytes,
                                static_cast<const char*>(A) + static_cast<size_t>(r0) * row_bytes,
                                static_cast<size_t>(rows) * row_bytes, h->stream));
            if (h->up_W) {
                cudaEvent_t e;
                FOS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                evs.push_back(e);
                FOS_CUDA(cudaEventRecord(e, h->stream));
                FOS_CUDA(cudaStreamWaitEvent(cs, e, 0));
                FOS_TRY(fos_upload_gram_chunk(h, r0, rows, cs));
            }
        }
        FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
        FOS_CUDA(cudaStreamWaitEvent(cs, h->ev1, 0));
        const auto u2 = now();
        FOS_TRY(fos_upload_gram_finish(h, cs));  // reduces the splits, synchronises cs
        const auto u3 = now();
        FOS_CUDA(cudaEventRecord(t_done, cs));
        FOS_CUDA(cudaStreamSynchronize(cs));
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        if (dbg)
            fprintf(stderr, "[fos] upload_dense: gram_begin %.1f ms, enqueue loop %.1f ms, finish (sync + free) %.1f ms\n",
                    ms(u0, u1), ms(u1, u2), ms(u2, u3));
        FOS_CUDA(cudaEventElapsedTime(&h->up_copy_ms, h->ev0, h->ev1));
        FOS_CUDA(cudaEventElapsedTime(&h->up_tail_ms, h->ev1, t_done));
        return FOS_OK;
    };
    const int st = body();
    if (st != FOS_OK) {
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(h->stream);
        fos_upload_gram_drop(h);
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    if (t_done) cudaEventDestroy(t_done);
    cudaStreamDestroy(cs);
    return st;
}

// Two-step creation: fos_design_create_begin allocates everything (matrix block, workspaces, control
// block -- recycled blocks in the steady state, so this takes well under a millisecond) and returns a
// handle whose matrix is still undefined; fos_design_upload then copies the host arrays.  Row-sharded
// callers run the upload on a side thread while the main thread wires the exchange windows of the same
// handle (socket hand-off + two all-gathers, ~0.1 s that used to follow the copy).
extern "C" int fos_design_create_begin(int64_t n, int64_t d, int dtype, int device, fos_design** out) {
    FOS_REQUIRE(out, "null pointer argument");
    fos_design* h = new fos_design();
    FOS_TRY_FREE(h, design_common_init(h, n, d, dtype, device));
    FOS_TRY_FREE(h, alloc_matrix(h));
    FOS_TRY_FREE(h, design_alloc_work(h));
    *out = h;
    return FOS_OK;
}

extern "C" int fos_design_upload(fos_design* h, const void* A, const double* b, int64_t row_stride, int64_t col_stride) {
    FOS_REQUIRE(h && A && b, "null pointer argument");
    FOS_REQUIRE(h->owns_A && h->owns_b, "this design borrows its arrays; it cannot be uploaded into");
    FOS_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->n, d = h->d;
    const int dtype = h->dtype;
    const int device = h->device;
    const auto t2 = std::chrono::steady_clock::now();
    const size_t es = elem_size(dtype);
    HostStager stager;
    auto body = [&]() -> int {
        const auto s0 = std::chrono::steady_clock::now();
        FOS_TRY(stager.init(device));
        if (getenv("FOS_UPLOAD_DEBUG"))
            fprintf(stderr, "[fos] stager.init %.1f ms\n",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s0).count());
        FOS_TRY(stager.copy(h->b, b, static_cast<size_t>(n) * sizeof(double), h->stream));
        if (col_stride == 1 && row_stride >= d) {
            // C order (possibly with a row pitch): strided copy straight into the padded layout
            if (h->lda == d && row_stride == d && fos_upload_gram_eligible(h)) {
                // dense float64, tall: copy in row chunks and push every chunk that has arrived
                // through the SYRK kernel on a second stream -- G = A^T A is ready a few ms after
                // the last byte (gram_kernels.cu), and estimate_lipschitz then iterates on G
                FOS_TRY(upload_dense(h, A, stager, true));
            } else if (h->lda == d && row_stride == d) {
                // dense on both sides: linear copies at the full PCIe rate
                FOS_TRY(upload_dense(h, A, stager, false));
            } else {
            if (h->lda != d)
                FOS_CUDA(cudaMemsetAsync(h->A, 0, static_cast<size_t>(n) * h->lda * es, h->stream));
            FOS_CUDA(cudaMemcpy2DAsync(h->A, static_cast<size_t>(h->lda) * es, A, static_cast<size_t>(row_stride) * es,
                                       static_cast<size_t>(d) * es, static_cast<size_t>(n), cudaMemcpyHostToDevice,
                                       h->stream));
            }
        } else if (row_stride == 1 && col_stride >= n) {
            // Fortran order: upload column-major row chunks and transpose on the device
            long long chunk = std::max<long long>(32, (256LL << 20) / (d * static_cast<long long>(es)));
            chunk = std::min<long long>(chunk, n);
            void* tmp = nullptr;
            FOS_CUDA(cudaMalloc(&tmp, static_cast<size_t>(chunk) * d * es));
            int st = FOS_OK;
            for (long long r0 = 0; r0 < n && st == FOS_OK; r0 += chunk) {
                const long long rows = std::min<long long>(chunk, n - r0);
                cudaError_t ce = cudaMemcpy2DAsync(tmp, static_cast<size_t>(rows) * es,
                                                   static_cast<const char*>(A) + static_cast<size_t>(r0) * es,
                                                   static_cast<size_t>(col_stride) * es, static_cast<size_t>(rows) * es,
                                                   static_cast<size_t>(d), cudaMemcpyHostToDevice, h->stream);
                if (ce != cudaSuccess) {
                    fos_set_error("column-major upload failed: %s", cudaGetErrorString(ce));
                    st = FOS_ERR_CUDA;
                    break;
                }
                st = fos_launch_repack(tmp, static_cast<char*>(h->A) + static_cast<size_t>(r0) * h->lda * es, rows,
                                       h->d, h->lda, 0, 0, dtype, h->stream);
            }
            cudaStreamSynchronize(h->stream);
            cudaFree(tmp);
            if (st != FOS_OK) return st;
        } else {
            fos_set_error("unsupported strides (%lld, %lld): pass a C- or Fortran-contiguous matrix",
                          static_cast<long long>(row_stride), static_cast<long long>(col_stride));
            return FOS_ERR_UNSUPPORTED;
        }
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        return FOS_OK;
    };
    {
        std::lock_guard<std::mutex> lock(*h->life_mu);
        h->uploading = true;
    }
    const int st_body = body();
    stager.drain();  // no copy may be in flight when the caller frees the design on the error path
    bool pending;
    {
        std::lock_guard<std::mutex> lock(*h->life_mu);
        h->uploading = false;
        pending = h->balance_pending;
        h->balance_pending = false;
    }
    if (st_body != FOS_OK) return st_body;
    if (pending) FOS_TRY(fos_balance_rows(h));
    if (getenv("FOS_UPLOAD_DEBUG")) {
        const auto t3 = std::chrono::steady_clock::now();
        fprintf(stderr, "[fos] design_upload %lld x %lld: copy %.1f ms (device %.1f)\n", static_cast<long long>(n),
                static_cast<long long>(d), std::chrono::duration<double, std::milli>(t3 - t2).count(), h->up_copy_ms);
    }
    return FOS_OK;
}

extern "C" int fos_design_create(const void* A, const double* b, int64_t n, int64_t d, int dtype,
                                 int64_t row_stride, int64_t col_stride, int device, fos_design** out) {
    FOS_REQUIRE(A && b && out, "null pointer argument");
    fos_design* h = nullptr;
    FOS_TRY(fos_design_create_begin(n, d, dtype, device, &h));
    FOS_TRY_FREE(h, fos_design_upload(h, A, b, row_stride, col_stride));
    *out = h;
    return FOS_OK;
}

extern "C" int fos_design_create_device(const void* A_dev, const double* b_dev, int64_t n, int64_t d, int dtype,
                                        int64_t lda, int device, fos_design** out) {
    FOS_REQUIRE(A_dev && b_dev && out, "null pointer argument");
    fos_design* h = new fos_design();
    FOS_TRY_FREE(h, design_common_init(h, n, d, dtype, device));
    const int es = elem_size(dtype);
    if (lda < d || (lda * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(A_dev) % 16) != 0) {
        fos_set_error("borrowed matrix needs lda >= d, lda*elem %% 16 == 0 and a 16-byte aligned base");
        design_free(h);
        return FOS_ERR_INVALID;
    }
    h->lda = static_cast<int>(lda);
    h->A = const_cast<void*>(A_dev);
    h->b = const_cast<double*>(b_dev);
    FOS_TRY_FREE(h, design_alloc_work(h));
    *out = h;
    return FOS_OK;
}

extern "C" int fos_design_create_synthetic(int64_t n, int64_t d, int dtype, uint64_t seed, double noise_std,
                                           double rho1, double rho2, int64_t row0, int device, fos_design** out) {
    FOS_REQUIRE(out, "null pointer argument");
    FOS_REQUIRE(fabs(rho1) <= 1.0 && fabs(rho2) <= 1.0, "correlations must lie in [-1, 1]");
    fos_design* h = new fos_design();
    FOS_TRY_FREE(h, design_common_init(h, n, d, dtype, device));
    FOS_TRY_FREE(h, alloc_matrix(h));
    FOS_TRY_FREE(h, fos_launch_synthetic(h, seed, noise_std, rho1, rho2, row0));
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        fos_set_error("synthetic generator failed: %s", cudaGetErrorString(e));
        design_free(h);
        return FOS_ERR_CUDA;
    }
    FOS_TRY_FREE(h, design_alloc_work(h));
    *out = h;
    return FOS_OK;
}

extern "C" int fos_design_destroy(fos_design* h) {
    design_free(h);
    return FOS_OK;
}

extern "C" int fos_design_shape(const fos_design* h, int64_t* n, int64_t* d, int* dtype, int64_t* lda) {
    FOS_REQUIRE(h, "null design");
    if (n) *n = h->n;
    if (d) *d = h->d;
    if (dtype) *dtype = h->dtype;
    if (lda) *lda = h->lda;
    return FOS_OK;
}

extern "C" int fos_design_upload_gram(fos_design* h, double** G_dev, int* state, float* copy_ms, float* tail_ms) {
    FOS_REQUIRE(h, "null design");
    if (G_dev) *G_dev = h->G_up;
    if (state) *state = h->G_up ? h->G_state : 0;
    if (copy_ms) *copy_ms = h->up_copy_ms;
    if (tail_ms) *tail_ms = h->up_tail_ms;
    return FOS_OK;
}

extern "C" int fos_design_upload_gram_set(fos_design* h, int state) {
    FOS_REQUIRE(h, "null design");
    FOS_REQUIRE(state == 0 || state == 2, "state must be 0 (discard) or 2 (held by every rank)");
    if (state == 0) {
        FOS_CUDA(cudaSetDevice(h->device));
        fos_upload_gram_drop(h);
    } else {
        FOS_REQUIRE(h->G_up != nullptr, "this design holds no Gram matrix");
        h->G_state = 2;
    }
    return FOS_OK;
}

extern "C" int fos_design_pointers(fos_design* h, void** A_dev, double** b_dev) {
    FOS_REQUIRE(h, "null design");
    if (A_dev) *A_dev = h->A;
    if (b_dev) *b_dev = h->b;
    return FOS_OK;
}

extern "C" int fos_design_download(fos_design* h, int64_t row0, int64_t rows, void* A_out, double* b_out) {
    FOS_REQUIRE(h, "null design");
    FOS_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= h->n, "row range out of bounds");
    FOS_CUDA(cudaSetDevice(h->device));
    const size_t es = elem_size(h->dtype);
    if (A_out && rows > 0)
        FOS_CUDA(cudaMemcpy2DAsync(A_out, static_cast<size_t>(h->d) * es,
                                   static_cast<const char*>(h->A) + static_cast<size_t>(row0) * h->lda * es,
                                   static_cast<size_t>(h->lda) * es, static_cast<size_t>(h->d) * es,
                                   static_cast<size_t>(rows), cudaMemcpyDeviceToHost, h->stream));
    if (b_out && rows > 0)
        FOS_CUDA(cudaMemcpyAsync(b_out, h->b + row0, static_cast<size_t>(rows) * sizeof(double),
                                 cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    return FOS_OK;
}

// ------------------------------------------------------------------------------------------
// helpers: vector upload/download through the pinned staging buffer
// ------------------------------------------------------------------------------------------
static int upload_vec(fos_design* h, double* dst_dev, const double* src_host, int slot) {
    double* st = h->vec_host + static_cast<size_t>(slot) * h->ldv;
    if (src_host)
        memcpy(st, src_host, static_cast<size_t>(h->d) * sizeof(double));
    else
        memset(st, 0, static_cast<size_t>(h->d) * sizeof(double));
    for (int c = h->d; c < h->ldv; ++c) st[c] = 0.0;
    FOS_CUDA(cudaMemcpyAsync(dst_dev, st, static_cast<size_t>(h->ldv) * sizeof(double), cudaMemcpyHostToDevice,
                             h->stream));
    return FOS_OK;
}

static int run_oneshot(fos_design* h, int g_mode, int eop, double a1, double a2, int bits) {
    FosHist none{};
    FOS_TRY(fos_launch_grad(h, g_mode));
    FOS_TRY(fos_launch_epilogue(h, eop, g_mode, none, a1, a2, bits));
    return FOS_OK;
}

extern "C" int fos_grad(fos_design* h, const double* x, double alpha2, double* g_out, double* loss_out) {
    FOS_REQUIRE(h && x, "null pointer argument");
    FOS_CUDA(cudaSetDevice(h->device));
    FOS_TRY(upload_vec(h, h->y, x, 0));
    FOS_TRY(run_oneshot(h, GM_GRAD, EOP_FG, 0.0, alpha2, alpha2 != 0.0 ? 2 : 0));
    FOS_CUDA(cudaMemcpyAsync(h->vec_host, h->g, static_cast<size_t>(h->ldv) * sizeof(double),
                             cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(FosCtrl), cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    if (g_out) memcpy(g_out, h->vec_host, static_cast<size_t>(h->d) * sizeof(double));
    if (loss_out) *loss_out = h->ctrl_host->out[0];
    return FOS_OK;
}

extern "C" int fos_objective(fos_design* h, const double* x, int reg_bits, double alpha1, double alpha2,
                             double* out) {
    FOS_REQUIRE(h && x && out, "null pointer argument");
    FOS_REQUIRE(reg_bits >= 0 && reg_bits <= 3, "reg_bits must be in 0..3");
    FOS_CUDA(cudaSetDevice(h->device));
    FOS_TRY(upload_vec(h, h->xc, x, 0));
    FOS_TRY(run_oneshot(h, GM_DOT2, EOP_OBJ, alpha1, alpha2, reg_bits));
    FOS_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(FosCtrl), cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    *out = h->ctrl_host->out[0];
    return FOS_OK;
}

extern "C" int fos_design_set_profile(fos_design* h, int enable) {
    FOS_REQUIRE(h, "null design");
    h->profile = enable != 0;
    return FOS_OK;
}

extern "C" int fos_time_grad_kernel(fos_design* h, int mode, int reps, float* ms_avg) {
    FOS_REQUIRE(h && ms_avg && reps >= 1, "bad argument");
    FOS_REQUIRE(mode > 0 && mode < 16, "mode must be a GM_* bit set");
    FOS_REQUIRE(!(mode & GM_PROBE) || h->kern_kind == 1, "the streaming probe needs the streaming kernel (d > 512)");
    FOS_CUDA(cudaSetDevice(h->device));
    FOS_TRY(fos_launch_grad(h, mode));  // warm-up
    FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < reps; ++i) FOS_TRY(fos_launch_grad(h, mode));
    FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    FOS_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *ms_avg = ms / reps;
    return FOS_OK;
}

// Debug: per-CTA start/end %globaltimer stamps of ONE gradient-kernel launch in `mode` (ns, relative
// to the earliest start).  out has 2*n_parts entries; returns n_parts through *n_parts.
extern "C" int fos_debug_cta_times(fos_design* h, int mode, long long* out, int cap, int* n_parts) {
    FOS_REQUIRE(h && out && n_parts, "null pointer argument");
    FOS_REQUIRE(cap >= 2 * h->n_parts, "output buffer too small");
    FOS_CUDA(cudaSetDevice(h->device));
    unsigned long long* buf = nullptr;
    FOS_CUDA(cudaMalloc(&buf, 2 * static_cast<size_t>(h->n_parts) * sizeof(unsigned long long)));
    FOS_CUDA(cudaMemset(buf, 0, 2 * static_cast<size_t>(h->n_parts) * sizeof(unsigned long long)));
    FOS_TRY(fos_launch_grad(h, mode));  // warm-up without stamps
    h->cta_times = buf;
    int st = fos_launch_grad(h, mode);
    h->cta_times = nullptr;
    std::vector<unsigned long long> host(2 * h->n_parts);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(host.data(), buf, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (st != FOS_OK) return st;
    if (e != cudaSuccess) {
        fos_set_error("debug launch failed: %s", cudaGetErrorString(e));
        return FOS_ERR_CUDA;
    }
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < h->n_parts; ++i) t0 = std::min(t0, host[2 * i]);
    for (int i = 0; i < h->n_parts; ++i) {
        const unsigned long long smid = host[2 * i + 1] >> 48;
        const unsigned long long end = host[2 * i + 1] & 0xFFFFFFFFFFFFull;
        // out[2i] = SM id * 2^40 + start offset, out[2i+1] = end offset  (ns)
        out[2 * i] = static_cast<long long>((smid << 40) | (host[2 * i] - t0));
        out[2 * i + 1] = static_cast<long long>(end - t0);
    }
    *n_parts = h->n_parts;
    return FOS_OK;
}

extern "C" int fos_debug_solve_profile(fos_design* h, unsigned long long* out9, int reset) {
    FOS_REQUIRE(h && out9, "null pointer argument");
    for (int i = 0; i < 9; ++i) out9[i] = 0;
    if (h->gsync == nullptr) return FOS_OK;
    FOS_CUDA(cudaSetDevice(h->device));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    FOS_CUDA(cudaMemcpy(out9, h->gsync->prof, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) FOS_CUDA(cudaMemset(h->gsync->prof, 0, sizeof(h->gsync->prof)));
    return FOS_OK;
}

extern "C" int fos_design_lambda_max(fos_design* h, double* out) {
    FOS_REQUIRE(h && out, "null pointer argument");
    std::vector<double> zero(h->d, 0.0), g(h->d);
    double loss;
    FOS_TRY(fos_grad(h, zero.data(), 0.0, g.data(), &loss));  // g = -A^T b
    double m = 0.0;
    for (double v : g) m = std::max(m, fabs(v));
    *out = m;
    return FOS_OK;
}

// Estimated duration of one pass (used to size launch batches between polls).
static int batch_size(const fos_design* h) {
    const double bytes = static_cast<double>(h->n) * h->lda * elem_size(h->dtype);
    const double t_pass = bytes / 5.0e12 + 12e-6;
    int b = static_cast<int>(1.0e-3 / t_pass);
    return std::max(1, std::min(b, 32));
}

// Launch (gradient, epilogue) pairs until the device reports completion.  `done` inspects a
// pinned snapshot of the control block.  At most two batches are in flight.
template <typename DoneFn>
static int drive_passes(fos_design* h, int eop, const FosHist& hist, long long max_pairs, DoneFn done,
                        long long* pairs_out, bool exact = false) {
    if (exact) {
        // The number of passes is known in advance: no polling, no snapshots.  Throttle the
        // launch queue with one event every 64 pairs.
        cudaEvent_t ev[2];
        FOS_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        FOS_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        int status = FOS_OK;
        long long launched = 0;
        int chunk = 0;
        while (launched < max_pairs && status == FOS_OK) {
            if (chunk >= 2) cudaEventSynchronize(ev[chunk & 1]);
            for (int i = 0; i < 64 && launched < max_pairs && status == FOS_OK; ++i, ++launched) {
                const bool prof = h->profile && h->prof_used + 2 <= h->prof_ev.size();
                if (prof) cudaEventRecord(h->prof_ev[h->prof_used], h->stream);
                status = fos_launch_grad(h, -1);
                if (prof) {
                    cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream);
                    h->prof_used += 2;
                }
                if (status == FOS_OK) status = fos_launch_epilogue(h, eop, 0, hist, 0.0, 0.0, 0);
            }
            cudaEventRecord(ev[chunk & 1], h->stream);
            ++chunk;
        }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
        if (status == FOS_OK && e != cudaSuccess) {
            fos_set_error("solver loop failed: %s", cudaGetErrorString(e));
            status = FOS_ERR_CUDA;
        }
        if (pairs_out) *pairs_out = launched;
        return status;
    }
    const int B = batch_size(h);
    cudaEvent_t ev[2];
    FOS_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    FOS_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    long long launched = 0;
    int status = FOS_OK;
    bool finished = false;
    int batch = 0;
    while (!finished && status == FOS_OK) {
        if (batch >= 2) {
            cudaError_t e = cudaEventSynchronize(ev[batch & 1]);
            if (e != cudaSuccess) {
                fos_set_error("solver loop failed: %s", cudaGetErrorString(e));
                status = FOS_ERR_CUDA;
                break;
            }
            if (done(h->ctrl_host[1 + (batch & 1)])) {
                finished = true;
                break;
            }
        }
        if (launched >= max_pairs) {
            // everything that can be needed is in flight: wait for the newest snapshot
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) {
                fos_set_error("solver loop failed: %s", cudaGetErrorString(e));
                status = FOS_ERR_CUDA;
                break;
            }
            const FosCtrl& last = h->ctrl_host[1 + ((batch + 1) & 1)];
            if (batch == 0 || done(last)) {
                finished = true;
            } else {
                fos_set_error("solver did not finish within %lld passes (non-finite data?)", max_pairs);
                status = FOS_ERR_INVALID;
            }
            break;
        }
        for (int i = 0; i < B && launched < max_pairs && status == FOS_OK; ++i, ++launched) {
            const bool prof = h->profile && h->prof_used + 2 <= h->prof_ev.size();
            if (prof) cudaEventRecord(h->prof_ev[h->prof_used], h->stream);
            status = fos_launch_grad(h, -1);
            if (prof) {
                cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream);
                h->prof_used += 2;
            }
            if (status == FOS_OK) status = fos_launch_epilogue(h, eop, 0, hist, 0.0, 0.0, 0);
        }
        if (status != FOS_OK) break;
        cudaError_t e = cudaMemcpyAsync(&h->ctrl_host[1 + (batch & 1)], h->ctrl, sizeof(FosCtrl),
                                        cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[batch & 1], h->stream);
        if (e != cudaSuccess) {
            fos_set_error("solver loop failed: %s", cudaGetErrorString(e));
            status = FOS_ERR_CUDA;
            break;
        }
        ++batch;
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (status == FOS_OK && e != cudaSuccess) {
        fos_set_error("solver loop failed: %s", cudaGetErrorString(e));
        status = FOS_ERR_CUDA;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (pairs_out) *pairs_out = launched;
    return status;
}

extern "C" int fos_power_iter(fos_design* h, const double* v0, int n_iter, double tol, double* L_out,
                              int* iters_out, float* gpu_ms_out) {
    FOS_REQUIRE(h && v0 && L_out, "null pointer argument");
    FOS_REQUIRE(n_iter >= 1, "n_iter must be >= 1");
    FOS_CUDA(cudaSetDevice(h->device));
    // a Gram matrix accumulated under the upload: iterate on it instead of streaming A twice a
    // hundred times
    // (row-sharded ranks: state 2 = every rank confirmed it holds the matrix of its own rows)
    if (h->G_up && ((h->world == 1 && h->G_state == 1) || (h->world > 1 && h->G_state == 2)))
        return fos_gram_power_iter(h, v0, n_iter, tol, L_out, iters_out, gpu_ms_out);
    FosCtrl* c = h->ctrl_host;
    memset(c, 0, sizeof(FosCtrl));
    c->g_mode = GM_GRAD | GM_NOB;
    c->phase = PH_DONE;
    c->ptol = tol;
    c->pit_max = n_iter;
    c->L_prev = 0.0;
    FOS_CUDA(cudaMemcpyAsync(h->ctrl, c, sizeof(FosCtrl), cudaMemcpyHostToDevice, h->stream));
    FOS_TRY(upload_vec(h, h->y, v0, 0));
    FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
    FosHist none{};
    long long pairs = 0;
    h->grad_only_hint = true;
    const int pst = drive_passes(h, EOP_POWER, none, n_iter, [](const FosCtrl& s) { return s.g_mode == GM_SKIP; }, &pairs);
    h->grad_only_hint = false;
    FOS_TRY(pst);
    FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
    FOS_CUDA(cudaMemcpyAsync(c, h->ctrl, sizeof(FosCtrl), cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    if (c->stop_reason < 0) {
        fos_set_error("multi-GPU exchange timed out: a peer rank never arrived");
        h->fused_ok = false;  // the grid counters of an abandoned launch are not reusable
        return FOS_ERR_COMM;
    }
    *L_out = c->L;
    if (iters_out) *iters_out = c->pit;
    if (gpu_ms_out) FOS_CUDA(cudaEventElapsedTime(gpu_ms_out, h->ev0, h->ev1));
    return FOS_OK;
}

int fos_arena_reserve(fos_design* h, size_t bytes, void** base) {
    if (bytes > h->arena_bytes) {
        fos_pool_free(h->arena);
        h->arena = nullptr;
        h->arena_bytes = 0;
        const size_t want = (bytes + (1u << 20) - 1) & ~static_cast<size_t>((1u << 20) - 1);
        if (fos_pool_malloc(&h->arena, want) != cudaSuccess) {
            cudaGetLastError();
            fos_set_error("cannot allocate %zu bytes of solver workspace on the device", want);
            return FOS_ERR_NOMEM;
        }
        h->arena_bytes = want;
    }
    *base = h->arena;
    return FOS_OK;
}

// ------------------------------------------------------------------------------------------
// proximal-gradient engine
// ------------------------------------------------------------------------------------------
extern "C" int fos_prox_grad(fos_design* h, const fos_pg_params* p, fos_pg_result* r) {
    FOS_REQUIRE(h && p && r, "null pointer argument");
    FOS_REQUIRE(p->scheme >= 0 && p->scheme <= 2, "unknown scheme %d", p->scheme);
    FOS_REQUIRE(p->max_iter >= 0, "max_iter must be >= 0");
    FOS_REQUIRE(p->step0 > 0.0 || p->max_iter == 0, "initial step must be positive");
    FOS_REQUIRE(!p->backtracking || (p->eta > 0.0 && p->eta < 1.0), "eta must lie in (0, 1) for backtracking");
    FOS_CUDA(cudaSetDevice(h->device));
    const int K = p->max_iter;
    const int d = h->d;
    const bool want_hist = p->want_history != 0;
    // ISTA records no objective in the reference (iterative_solvers.py:83); the engine gets it
    // for free (second dot of the same pass) and the drop-in exposes it as an extra.
    const bool want_obj = want_hist;
    const long long launches0 = h->launches;

    // ---- per-solve device arrays: carved out of one grow-only arena owned by the design
    // (cudaMalloc / cudaFree per solve cost milliseconds each once peer mappings exist)
    const auto t_host0 = std::chrono::steady_clock::now();
    struct Carve {
        size_t off = 0;
        size_t take(size_t bytes) {
            const size_t o = off;
            off += (bytes + 255) & ~static_cast<size_t>(255);
            return o;
        }
    } cv;
    const size_t K1 = static_cast<size_t>(std::max(K, 1));
    const size_t o_xh = cv.take(want_hist ? static_cast<size_t>(K + 1) * d * sizeof(double) : 0);
    const size_t o_oh = cv.take(K1 * sizeof(double));
    const size_t o_th = cv.take((K1 + 1) * sizeof(double));
    const size_t o_sh = cv.take(K1 * sizeof(double));
    const size_t o_li = cv.take(K1 * sizeof(int));
    const size_t o_gm = cv.take((K1 + 1) * sizeof(float));
    const size_t o_lm = cv.take(K1 * sizeof(float));
    void* arena_base = nullptr;
    FOS_TRY(fos_arena_reserve(h, cv.off, &arena_base));
    char* base = static_cast<char*>(arena_base);
    struct { double* p; } xh{reinterpret_cast<double*>(base + o_xh)}, oh{reinterpret_cast<double*>(base + o_oh)},
        th{reinterpret_cast<double*>(base + o_th)}, sh{reinterpret_cast<double*>(base + o_sh)};
    struct { int* p; } li{reinterpret_cast<int*>(base + o_li)};
    struct { float* p; } gm{reinterpret_cast<float*>(base + o_gm)}, lm{reinterpret_cast<float*>(base + o_lm)};
    // everything after the iterate history is zeroed in one go (objectives, steps, counts, timings)
    FOS_CUDA(cudaMemsetAsync(base + o_oh, 0, cv.off - o_oh, h->stream));
    FosHist hist{};
    hist.x_hist = want_hist ? xh.p : nullptr;
    hist.obj_hist = oh.p;
    hist.t_hist = th.p;
    hist.step_hist = sh.p;
    hist.ls_iters = li.p;
    hist.grad_ms = gm.p;
    hist.ls_ms = lm.p;

    // ---- control block and start point
    FosCtrl* c = h->ctrl_host;
    memset(c, 0, sizeof(FosCtrl));
    c->scheme = p->scheme;
    c->backtracking = p->backtracking;
    c->adaptive_restart = p->adaptive_restart;
    c->want_hist = want_hist;
    c->want_obj = want_obj;
    c->max_iter = K;
    c->obj_terms = p->obj_terms;
    c->alpha1 = p->alpha1;
    c->alpha2 = p->alpha2;
    c->eta = p->eta;
    c->tol = p->tol;
    c->tol_ratio = p->tol_ratio;
    c->restart_thr = p->restart_threshold;
    c->delta = p->delta;
    c->armijo_c = p->armijo_c;
    // Fixed-step solves with history on streaming designs: the recorded objective's residual norm comes from
    // the row-wise residual recurrence (GM_QREC) instead of a second dot product in every pass (FOS_QREC=0: off)
    // It trades 16 FP64 FMAs per 16 elements for ~40 cycles of latency per stage on one warp.  Which one costs
    // more depends on the regime (measured, alternating A/B on the same box): a solve long enough to sit at the
    // power cap gains (1M x 4096: 4.62 vs 5.11 ms per pass; 8 x 125k rows sustained: 1400 vs 1366 it/s), a short
    // burst at full clocks loses (125k x 4096, 20 iterations = 13 ms: 0.645 vs 0.628 ms per pass).  Hence: on
    // when the solve is expected to stream for >= 60 ms (max_iter passes at 6.5 TB/s); FOS_QREC=1 / 0 forces it.
    bool qrec = want_obj && !p->backtracking && h->kern_kind == 1 && K > 0;
    {
        const double est_ms = static_cast<double>(K) * static_cast<double>(h->n) * h->lda * elem_size(h->dtype) / 6.5e9;
        const char* e = getenv("FOS_QREC");
        if (e && e[0] == '0') qrec = false;
        else if (!(e && e[0] == '1')) qrec = qrec && est_ms >= 60.0;
    }
    if (qrec && h->qres == nullptr &&
        fos_pool_malloc(reinterpret_cast<void**>(&h->qres), 2 * static_cast<size_t>(h->n) * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        h->qres = nullptr;
        qrec = false;
    }
    c->use_qrec = qrec ? 1 : 0;
    c->beta_y = 0.0;   // y_0 = x_0
    c->phase = (K > 0) ? PH_GRAD : PH_DONE;
    c->g_mode = (K > 0) ? (GM_GRAD | (qrec ? GM_QREC : 0)) : GM_SKIP;
    c->tau = p->step0;
    c->trial_t = p->step0;
    c->t_mom = 1.0;
    c->prev_step = 0.0;
    FOS_CUDA(cudaMemcpyAsync(h->ctrl, c, sizeof(FosCtrl), cudaMemcpyHostToDevice, h->stream));
    FOS_TRY(upload_vec(h, h->y, p->x0, 0));
    const size_t vb = static_cast<size_t>(h->ldv) * sizeof(double);
    FOS_CUDA(cudaMemcpyAsync(h->xc, h->y, vb, cudaMemcpyDeviceToDevice, h->stream));
    FOS_CUDA(cudaMemcpyAsync(h->xk, h->y, vb, cudaMemcpyDeviceToDevice, h->stream));
    if (want_hist)
        FOS_CUDA(cudaMemcpyAsync(xh.p, h->y, static_cast<size_t>(d) * sizeof(double), cudaMemcpyDeviceToDevice,
                                 h->stream));
    FOS_CUDA(cudaMemcpyAsync(th.p, &h->ctrl->tau, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));

    // ---- the loop
    long long max_pairs;
    const bool exact = !p->backtracking && p->tol <= 0.0 && p->tol_ratio <= 0.0;
    if (exact)
        max_pairs = static_cast<long long>(K) + (want_obj ? 1 : 0);  // known exactly
    else
        max_pairs = static_cast<long long>(K) * 1200 + 8;  // tau underflows to 0 well before
    h->prof_used = 0;
    if (h->profile) {
        const size_t want = std::min<size_t>(2 * static_cast<size_t>(max_pairs), 8192);
        while (h->prof_ev.size() < want) {
            cudaEvent_t e;
            FOS_CUDA(cudaEventCreate(&e));
            h->prof_ev.push_back(e);
        }
    }
    FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
    const auto t_host1 = std::chrono::steady_clock::now();
    long long pairs = 0;
    bool fused = K > 0 && fos_solve_stages(h) > 0;
    if (fused) {
        // ONE launch of the persistent kernel runs every pass, the in-kernel epilogues and (row-sharded
        // designs) the peer exchange; it leaves when the state machine reaches PH_DONE.  A refused
        // launch (the cooperative launch cannot place every CTA: SMs reserved by another context)
        // switches the design to the two-launch path for good.
        // The persistent kernel holds every SM until the solve is over, and on a row-sharded design it waits
        // for its peers: a collective of ANOTHER library enqueued earlier on another stream of this device
        // (say an NCCL all-reduce the peers are already blocked in) must not end up queued behind it.
        if (h->world > 1) cudaDeviceSynchronize();
        if (fos_launch_solve(h, hist, max_pairs) == FOS_OK) {
            pairs = 1;
        } else {
            cudaGetLastError();
            h->fused_ok = false;
            fused = false;
        }
    }
    if (!fused && K > 0) {
        FOS_TRY(drive_passes(h, EOP_PG, hist, max_pairs, [](const FosCtrl& s) { return s.phase == PH_DONE; }, &pairs,
                             exact));
    }
    FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
    const auto t_host2 = std::chrono::steady_clock::now();

    // ---- results
    FOS_CUDA(cudaMemcpyAsync(c, h->ctrl, sizeof(FosCtrl), cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaMemcpyAsync(h->vec_host, h->xk, vb, cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    if (c->stop_reason < 0) {
        fos_set_error("multi-GPU exchange timed out: a peer rank never arrived");
        h->fused_ok = false;  // the grid counters of an abandoned launch are not reusable
        return FOS_ERR_COMM;
    }
    if (c->phase != PH_DONE && K > 0) {
        fos_set_error("solver stopped in phase %d after %lld passes", c->phase, pairs);
        return FOS_ERR_INVALID;
    }
    const int it = c->k;
    if (r->x) memcpy(r->x, h->vec_host, static_cast<size_t>(d) * sizeof(double));
    if (r->x_hist && want_hist)
        FOS_CUDA(cudaMemcpy(r->x_hist, xh.p, static_cast<size_t>(it + 1) * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (r->obj_hist && it > 0) FOS_CUDA(cudaMemcpy(r->obj_hist, oh.p, it * sizeof(double), cudaMemcpyDeviceToHost));
    if (r->t_hist) FOS_CUDA(cudaMemcpy(r->t_hist, th.p, (it + 1) * sizeof(double), cudaMemcpyDeviceToHost));
    if (r->step_hist && it > 0) FOS_CUDA(cudaMemcpy(r->step_hist, sh.p, it * sizeof(double), cudaMemcpyDeviceToHost));
    if (r->ls_iters && it > 0) FOS_CUDA(cudaMemcpy(r->ls_iters, li.p, it * sizeof(int), cudaMemcpyDeviceToHost));
    if (r->grad_ms && c->n_grad_calls > 0)
        FOS_CUDA(cudaMemcpy(r->grad_ms, gm.p, c->n_grad_calls * sizeof(float), cudaMemcpyDeviceToHost));
    if (r->ls_ms && it > 0) FOS_CUDA(cudaMemcpy(r->ls_ms, lm.p, it * sizeof(float), cudaMemcpyDeviceToHost));
    r->n_iters = it;
    r->n_grad_calls = c->n_grad_calls;
    r->n_passes = c->n_passes;
    r->stop_reason = c->stop_reason;
    FOS_CUDA(cudaEventElapsedTime(&r->loop_ms, h->ev0, h->ev1));
    r->kernel_launches = h->launches - launches0;
    r->epilogue_ms = static_cast<float>(c->epi_ns * 1e-6);
    r->exchange_ms = static_cast<float>(c->xchg_ns * 1e-6);
    r->grad_kernel_ms = 0.f;
    r->grad_kernel_launches = 0;
    if (fused) {
        // device-stamped (%globaltimer) duration of the gradient phases: pass start -> every CTA's partials
        // published; the rest of loop_ms is the in-kernel tails
        r->grad_kernel_ms = static_cast<float>(c->grad_ns * 1e-6);
        r->grad_kernel_launches = c->n_passes;
    }
    // passes launched after the device reported completion exit immediately: count only the
    // ones that did work (the first n_passes)
    for (size_t i = 0; i + 1 < h->prof_used && static_cast<int>(i / 2) < c->n_passes; i += 2) {
        float ms = 0.f;
        FOS_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
        r->grad_kernel_ms += ms;
        r->grad_kernel_launches += 1;
    }
    const auto t_host3 = std::chrono::steady_clock::now();
    r->host_setup_ms = std::chrono::duration<float, std::milli>(t_host1 - t_host0).count();
    r->host_loop_ms = std::chrono::duration<float, std::milli>(t_host2 - t_host1).count();
    r->host_finish_ms = std::chrono::duration<float, std::milli>(t_host3 - t_host2).count();
    return FOS_OK;
}

// ------------------------------------------------------------------------------------------
// standalone prox operators on host buffers
// ------------------------------------------------------------------------------------------
static int prox_host(const double* v, int64_t len, double thresh, double scale, double* out, int device) {
    FOS_REQUIRE(len >= 0 && (len == 0 || (v && out)), "null pointer argument");
    if (len == 0) return FOS_OK;
    if (fos_device_count() <= 0) {
        fos_set_error("no CUDA device visible: libfos_b200 has no CPU fallback");
        return FOS_ERR_CUDA;
    }
    FOS_CUDA(cudaSetDevice(device));
    double* dv = nullptr;
    FOS_CUDA(cudaMalloc(&dv, 2 * static_cast<size_t>(len) * sizeof(double)));
    int st = FOS_OK;
    cudaError_t e = cudaMemcpy(dv, v, static_cast<size_t>(len) * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = fos_launch_prox(dv, dv + len, len, thresh, scale, nullptr);
        if (st == FOS_OK) e = cudaMemcpy(out, dv + len, static_cast<size_t>(len) * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(dv);
    if (st != FOS_OK) return st;
    if (e != cudaSuccess) {
        fos_set_error("prox failed: %s", cudaGetErrorString(e));
        return FOS_ERR_CUDA;
    }
    return FOS_OK;
}

extern "C" int fos_prox_l1(const double* v, int64_t len, double thresh, double* out, int device) {
    return prox_host(v, len, thresh, 1.0, out, device);
}

extern "C" int fos_prox_elastic_net(const double* v, int64_t len, double tau, double alpha1, double alpha2,
                                    double* out, int device) {
    return prox_host(v, len, tau * alpha1, 1.0 + tau * alpha2, out, device);
}

// ------------------------------------------------------------------------------------------
// multi-GPU plumbing: peer-memory exchange windows (CUDA IPC), used by the fused all-reduce
// inside the epilogue kernel
// ------------------------------------------------------------------------------------------
// Layout: [pull region: 2 slots x (ldv + PAD) doubles][FOS_MAX_WORLD + 8 u64: arrival flags, exchange counter]
//         [push region: 2 slots x FOS_MAX_WORLD x (ldv + PAD) doubles][FOS_MAX_WORLD x FOS_MAX_PARTS u64 flags]
static size_t window_doubles(const fos_design* h) { return 2 * static_cast<size_t>(h->ldv + FOS_WIN_PAD); }
static size_t window_push_doubles(const fos_design* h) {
    return 2 * static_cast<size_t>(FOS_MAX_WORLD) * static_cast<size_t>(h->ldv + FOS_WIN_PAD);
}
size_t fos_window_bytes(const fos_design* h) {
    return window_doubles(h) * sizeof(double) + (FOS_MAX_WORLD + 8) * sizeof(unsigned long long) +
           window_push_doubles(h) * sizeof(double) +
           static_cast<size_t>(FOS_MAX_WORLD) * FOS_MAX_PARTS * sizeof(unsigned long long);
}
void fos_window_bind(fos_design* h, int r, void* base) {
    h->peer.win[r] = static_cast<double*>(base);
    h->peer.flag[r] = reinterpret_cast<unsigned long long*>(h->peer.win[r] + window_doubles(h));
    h->peer.fwin[r] = reinterpret_cast<double*>(h->peer.flag[r] + FOS_MAX_WORLD + 8);
    h->peer.fflag[r] = reinterpret_cast<unsigned long long*>(h->peer.fwin[r] + window_push_doubles(h));
    if (r == h->rank) h->peer.epoch = h->peer.flag[r] + FOS_MAX_WORLD;
}

extern "C" int fos_comm_window_alloc(fos_design* h, int rank, int world, void* ipc_handle_out64) {
    FOS_REQUIRE(h && ipc_handle_out64, "null pointer argument");
    FOS_REQUIRE(world >= 1 && world <= FOS_MAX_WORLD, "world size must be in 1..%d", FOS_MAX_WORLD);
    FOS_REQUIRE(rank >= 0 && rank < world, "rank %d out of range", rank);
    FOS_REQUIRE(h->window == nullptr, "exchange window already allocated");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    FOS_CUDA(cudaSetDevice(h->device));
    h->window_bytes = fos_window_bytes(h);
    FOS_CUDA(cudaMalloc(&h->window, h->window_bytes));
    FOS_CUDA(cudaMemset(h->window, 0, h->window_bytes));
    cudaIpcMemHandle_t hd;
    FOS_CUDA(cudaIpcGetMemHandle(&hd, h->window));
    memcpy(ipc_handle_out64, &hd, sizeof(hd));
    h->rank = rank;
    h->world = 1;  // becomes `world` once the peers are attached
    h->peer_base[rank] = nullptr;
    fos_window_bind(h, rank, h->window);
    // the exchange counter starts at 1 so that a zeroed flag never satisfies a wait
    unsigned long long one = 1;
    FOS_CUDA(cudaMemcpy(h->peer.epoch, &one, sizeof(one), cudaMemcpyHostToDevice));
    return FOS_OK;
}

extern "C" int fos_comm_attach(fos_design* h, const void* ipc_handles, int world) {
    FOS_REQUIRE(h && ipc_handles, "null pointer argument");
    FOS_REQUIRE(h->window != nullptr, "call fos_comm_window_alloc first");
    FOS_REQUIRE(world >= 1 && world <= FOS_MAX_WORLD && h->rank < world, "bad world size %d", world);
    FOS_CUDA(cudaSetDevice(h->device));
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(ipc_handles);
    for (int r = 0; r < world; ++r) {
        if (r == h->rank) continue;
        void* base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, hs[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            fos_set_error("cannot map the exchange window of rank %d: %s", r, cudaGetErrorString(e));
            return FOS_ERR_COMM;
        }
        h->peer_base[r] = base;
        fos_window_bind(h, r, base);
    }
    h->world = world;
    return fos_comm_after_attach(h);
}

int fos_comm_after_attach(fos_design* h) {
    const char* e = getenv("FOS_BALANCE");
    if (h->world >= 4 && !(e && e[0] == '0') && !h->balanced) {
        {
            std::lock_guard<std::mutex> lock(*h->life_mu);
            if (h->uploading) {  // calibrate once the copy is done (fos_design_upload's last step)
                h->balance_pending = true;
                return FOS_OK;
            }
        }
        FOS_TRY(fos_balance_rows(h));
    }
    return FOS_OK;
}

extern "C" int fos_comm_info(const fos_design* h, int* rank, int* world) {
    FOS_REQUIRE(h, "null design");
    if (rank) *rank = h->rank;
    if (world) *world = h->world;
    return FOS_OK;
}
