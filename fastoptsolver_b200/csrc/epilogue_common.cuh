// epilogue_common.cuh -- device helpers shared by the epilogue kernels (epilogue_kernels.cu,
// lbfgs_kernels.cu): ordered cluster reductions through distributed shared memory, ordered
// sums of the per-CTA partials, and the fused multi-GPU exchange over peer memory.
#pragma once

#include <cooperative_groups.h>

#include "fos_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NS = FOS_NSCAL;
constexpr int EW = FOS_EPI_THREADS / 32;

struct Shared {
    double wred[EW][NS];
    double cl[FOS_EPI_CLUSTER][NS];  // written by every CTA of the cluster (DSMEM)
    double sc[2];
};

// Sum NS per-thread values over the whole cluster in a fixed order; result identical in
// every thread of every CTA.
__device__ __forceinline__ void cluster_sum(double (&v)[NS], Shared& sh) {
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < NS; ++k) v[k] = fos_warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NS; ++k) sh.wred[warp][k] = v[k];
    }
    __syncthreads();
    if (tid < NS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < EW; ++w) t += sh.wred[w][tid];
        const unsigned me = cluster.block_rank();
        for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
            double* dst = cluster.map_shared_rank(&sh.cl[me][tid], r);
            *dst = t;
        }
    }
    cluster.sync();
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double t = 0.0;
        for (unsigned r = 0; r < cluster.num_blocks(); ++r) t += sh.cl[r][k];
        v[k] = t;
    }
}

// ---------------------------------------------------------------- multi-GPU exchange
// One-shot all-reduce of the (d + 2)-vector [A^T r partial, s1, s2] across the row shards,
// fused into this kernel: every rank writes its local sum into its own window, publishes an
// arrival counter into every peer (st.release.sys over NVLink), waits for all peers, and
// then reads all windows in rank order -- so every rank forms bit-identical sums and the
// replicated solver state never diverges.  Windows are double-buffered by exchange parity.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_sys_v2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ size_t win_slot_offset(const EpiArgs& e, unsigned long long epoch) {
    return static_cast<size_t>(epoch & 1ull) * static_cast<size_t>(e.ldv + FOS_WIN_PAD);
}

// ordered sum of the per-CTA residual-norm partials; same value in every thread
__device__ __forceinline__ void load_pass_scalars(const EpiArgs& e, Shared& sh, double& s1, double& s2) {
    if (threadIdx.x < 32) {
        double a = 0.0, b = 0.0;
        for (int p = threadIdx.x; p < e.n_parts; p += 32) {
            a += e.partial_s[2 * p + 0];
            b += e.partial_s[2 * p + 1];
        }
        a = fos_warp_sum(a);
        b = fos_warp_sum(b);
        if (threadIdx.x == 0) {
            sh.sc[0] = a;
            sh.sc[1] = b;
        }
    }
    __syncthreads();
    s1 = sh.sc[0];
    s2 = sh.sc[1];
}

// ordered column sum of the partial gradients for the column pair starting at c
__device__ __forceinline__ double2 column_sum(const EpiArgs& e, int c) {
    double2 s = make_double2(0.0, 0.0);
    const double* p = e.partial_g + c;
    int i = 0;
    for (; i + 8 <= e.n_parts; i += 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            v[u] = *reinterpret_cast<const double2*>(p + static_cast<size_t>(i + u) * e.ldv);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s.x += v[u].x;
            s.y += v[u].y;
        }
    }
    for (; i < e.n_parts; ++i) {
        const double2 v = *reinterpret_cast<const double2*>(p + static_cast<size_t>(i) * e.ldv);
        s.x += v.x;
        s.y += v.y;
    }
    return s;
}

// Publish the local sums and wait for every rank's.  Returns false on timeout (a peer died).
__device__ __forceinline__ bool peer_exchange(const EpiArgs& e, Shared& sh, bool has_grad, double s1, double s2,
                                              unsigned long long epoch) {
    cg::cluster_group cluster = cg::this_cluster();
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    double* mine = e.peer.win[e.rank] + win_slot_offset(e, epoch);
    if (has_grad) {
        for (int c = 2 * (static_cast<int>(cluster.block_rank()) * FOS_EPI_THREADS + static_cast<int>(threadIdx.x));
             c < e.ldv; c += 2 * FOS_EPI_THREADS * FOS_EPI_CLUSTER)
            *reinterpret_cast<double2*>(mine + c) = column_sum(e, c);
    }
    if (leader) {
        mine[e.ldv + 0] = s1;
        mine[e.ldv + 1] = s2;
    }
    __threadfence_system();
    cluster.sync();
    // one thread per peer publishes the arrival counter (the stores travel over NVLink in parallel)
    if (cluster.block_rank() == 0 && threadIdx.x < e.world)
        st_release_sys(e.peer.flag[threadIdx.x] + e.rank, epoch);
    if (threadIdx.x == 0) sh.sc[0] = 1.0;
    __syncthreads();
    if (threadIdx.x < e.world) {
        const unsigned long long* f = e.peer.flag[e.rank] + threadIdx.x;
        unsigned long long spins = 0;
        while (ld_acquire_sys(f) < epoch) {
            if (++spins > (1ull << 28)) {
                sh.sc[0] = 0.0;
                break;
            }
        }
    }
    __syncthreads();
    const bool ok = sh.sc[0] != 0.0;
    __syncthreads();
    return ok;
}

// the two pass scalars summed over ranks in rank order (after peer_exchange)
__device__ __forceinline__ void reduced_scalars(const EpiArgs& e, Shared& sh, unsigned long long epoch, double& s1,
                                                double& s2) {
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < e.world; ++r) {
            const double2 v = ld_sys_v2(e.peer.win[r] + win_slot_offset(e, epoch) + e.ldv);
            a += v.x;
            b += v.y;
        }
        sh.sc[0] = a;
        sh.sc[1] = b;
    }
    __syncthreads();
    s1 = sh.sc[0];
    s2 = sh.sc[1];
    __syncthreads();
}

// column pair of the gradient summed over CTAs (one GPU) or over ranks (after peer_exchange)
__device__ __forceinline__ double2 reduced_column(const EpiArgs& e, int c, unsigned long long epoch) {
    if (e.world <= 1) return column_sum(e, c);
    double2 s = make_double2(0.0, 0.0);
    const size_t off = win_slot_offset(e, epoch) + c;
    double2 v[FOS_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < FOS_MAX_WORLD; ++r)
        if (r < e.world) v[r] = ld_sys_v2(e.peer.win[r] + off);
#pragma unroll
    for (int r = 0; r < FOS_MAX_WORLD; ++r)
        if (r < e.world) {
            s.x += v[r].x;
            s.y += v[r].y;
        }
    return s;
}

#define FOR_MY_COLUMN_PAIRS(c)                                                              \
    for (int c = 2 * (static_cast<int>(cg::this_cluster().block_rank()) * FOS_EPI_THREADS +  \
                      static_cast<int>(threadIdx.x));                                        \
         c < e.ldv; c += 2 * FOS_EPI_THREADS * FOS_EPI_CLUSTER)


}  // namespace
