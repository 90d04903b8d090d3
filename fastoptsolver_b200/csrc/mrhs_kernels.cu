// mrhs_kernels.cu -- streaming multi-RHS mode (north_star item 4: "A^T A and A X really are dense
// contractions"): fixed-step FISTA for a whole batch of L1 penalties WITHOUT the Gram matrix, for
// designs where d^2 does not fit or n is not >> d.  Per iteration and per batch of 8 penalties A is
// streamed from HBM ONCE and both contractions run on the fp64 tensor pipe (mma.sync.m8n8k4.f64;
// tcgen05 has no f64 kind):
//
//     U = A Y - b 1^T        (rows x 8)     D[8 rows x 8 lambda] += A_tile[8 x 4] . Y[4 x 8]
//     Gpart += A^T U         (d x 8)        D[8 cols x 8 lambda] += A_tile^T[8 x 4] . U[4 x 8]
//
// i.e. the batched form of iterative_solvers.py:173-175 (r = A y - b; grad = A^T r), each column
// with its own penalty.  Arithmetic intensity 32 flop / 8 B: the kernel sits between the HBM and
// the DMMA roofline (76 % tensor-pipe duty at the HBM rate for d = 4096).  Measured (500 000 x 4096,
// 8 penalties): 5.7 ms per pass = 2.9 TB/s effective / 11.5 TFLOP/s, 3.3 x faster than eight
// single-penalty passes; at d = 4096 only three 64 KB stages fit beside a tile that is still needed
// for the second contraction, so the ring depth, not the tensor pipe, sets the pace.
//
// Decomposition.  Y (d x 8) and the accumulators (d x 8) do not fit one CTA's registers at d = 4096,
// so a thread-block CLUSTER of 4 CTAs shares a row block: CTA q owns the column quarter q (Y and
// accumulator slices live in registers: 32 + 32 doubles per lane), streams the quarter-rows of its
// 8-row tiles through its own cp.async.bulk ring (one bulk copy per row, padded pitch: conflict-free
// DMMA fragment loads), and the partial products U_q (8 x 8) are summed across the cluster through
// distributed shared memory in rank order -- one split-phase cluster barrier per tile, overlapped with
// the first contraction of the NEXT tile (software pipeline: arrive(t+1) ... wait(t+1) one tile later).
// Every sum has a fixed order: results are bit-reproducible.
//
//   mrhs_stream_kernel<CW>   one pass: per-cluster partial gradients + per-column residual norms
//   mrhs_update_kernel       sum of the cluster partials (fixed order), (+a2 y), soft threshold with
//                            the column's penalty, Nesterov step (same roundings as pg_logic.cuh)
//   mrhs_obj_kernel          objective of every column from a norms-only pass
#include <cooperative_groups.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "fos_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int MR_CLUSTER = 4;    // CTAs per row block (column quarters); d = 4096 may use 8 (column eighths), see fos_mrhs_fista
constexpr int MR_CLMAX = 8;
constexpr int MR_TILE = 8;       // rows per tile (the M of the first contraction, the K of the second)
constexpr int MR_NB = 8;         // penalties per pass (the N of both contractions)
constexpr int MR_PAD = 32;       // bytes added to the shared row pitch (8 words: conflict-free fragments)
constexpr int MR_MAXST = 8;

__device__ __forceinline__ uint32_t mr_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mr_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mr_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mr_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MR_DONE;\n\t"
        "bra MR_WAIT;\n\t"
        "MR_DONE:\n\t"
        "}" ::"r"(mr_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mr_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            mr_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(mr_smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void mr_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

struct MrhsArgs {
    const double* A;        // n x lda row-major fp64
    const double* b;        // n
    const double* Y;        // [8][ldv] the batch's points (row l = penalty l)
    double* part;           // [n_clusters][8][ldv] per-cluster partial gradients
    double* norms;          // [n_clusters][8] per-cluster sum_i r_il^2
    const long long* row_lo;  // [n_clusters + 1], multiples of 8 except the last entry (= n)
    int d, lda, ldv;
    int grad;               // 1: gradient + norms, 0: norms only (objective pass)
    int nstage, stage_bytes, pitch;  // ring geometry (pitch in bytes)
    unsigned long long* prof;        // debug (nullable): clock64 cycles of block 0 / thread 0 per phase
};

constexpr int MR_UBUF = 4;       // buffers of the cluster exchange (a CTA may run up to 3 tiles ahead of a slow peer warp)
constexpr int MR_THREADS = 288;  // 8 consumer warps + 1 producer warp

struct MrhsSmem {
    uint64_t full[MR_MAXST];          // tile landed (bulk-copy bytes)
    uint64_t empty[MR_MAXST];         // all 8 consumer warps are done with the slot
    uint64_t redbar[2];               // all 8 warps stored their partial U of a tile
    uint64_t ubar[MR_UBUF];           // the CTA partials of all 4 cluster CTAs arrived (st.async bytes)
    double red[2][8][64];             // [parity][warp][row*8 + lambda] warp partials of U
    double clu[MR_UBUF][MR_CLMAX][64];  // [buffer][source CTA][row*8 + lambda], written by st.async from the peers
};

__device__ __forceinline__ void mr_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mr_mapa(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
// 8-byte store into a peer CTA's shared memory that also counts its bytes on the peer's mbarrier
__device__ __forceinline__ void mr_st_async(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v),
                 "r"(remote_bar)
                 : "memory");
}

// CW: columns per warp (d / 32): 128 for d = 4096, 64 for d = 2048, 32 for d = 1024.
// Control flow: no block-wide barrier inside the loop.  Warp 8 only moves data (waits for a free slot,
// requests the next tile); warps 0-7 run   contract1(t+1) -> [warps 0,1: publish(t+1)] -> contract2(t)
// and meet only through mbarriers: redbar (warp partials complete), ubar (the 4 CTA partials arrived
// from the cluster, counted in bytes by st.async), empty (slot reusable).
template <int CW, int CL = MR_CLUSTER>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(MR_THREADS, 1) mrhs_stream_kernel(const MrhsArgs a) {
    constexpr int KS = CW / 4;   // k-steps of the first contraction per warp
    constexpr int MB = CW / 8;   // column blocks of the second contraction per warp
    constexpr int NCH = (KS >= 16) ? 16 : 8;  // independent accumulation chains of the first contraction
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(16) MrhsSmem sm;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fk = lane & 3, fc = lane >> 2;
    const int q = static_cast<int>(cluster.block_rank());      // column quarter
    const int cl = blockIdx.x / CL;                            // row block
    const long long lo = a.row_lo[cl], hi = a.row_lo[cl + 1];
    const int ntile = static_cast<int>((hi - lo + MR_TILE - 1) / MR_TILE);
    const int dq = a.d / CL;
    const uint32_t row_bytes = static_cast<uint32_t>(dq) * 8u;

    if (tid == 0) {
        for (int s = 0; s < a.nstage; ++s) {
            mr_mbar_init(&sm.full[s], 1);
            mr_mbar_init(&sm.empty[s], 8);
        }
        mr_mbar_init(&sm.redbar[0], 8);
        mr_mbar_init(&sm.redbar[1], 8);
        for (int u = 0; u < MR_UBUF; ++u) mr_mbar_init(&sm.ubar[u], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();   // every CTA's barriers exist before anyone signals a peer

    if (warp == 8) {
        // ===== producer warp: one lane requests tiles; a slot is refilled as soon as the 8 consumer warps released it
        if (lane == 0) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            int slot = 0;
            uint32_t round = 0;
            for (int t = 0; t < ntile; ++t) {
                if (round > 0) mr_mbar_wait(&sm.empty[slot], (round - 1u) & 1u);
                const long long r0 = lo + static_cast<long long>(t) * MR_TILE;
                const int rows = static_cast<int>(min(static_cast<long long>(MR_TILE), hi - r0));
                mr_mbar_expect_tx(&sm.full[slot], static_cast<uint32_t>(rows) * row_bytes);
                unsigned char* dst = ring + static_cast<size_t>(slot) * a.stage_bytes;
                const double* src = a.A + r0 * a.lda + q * dq;
                for (int r = 0; r < rows; ++r)
                    mr_bulk_g2s(dst + static_cast<size_t>(r) * a.pitch, src + static_cast<size_t>(r) * a.lda, row_bytes,
                                &sm.full[slot], pol);
                if (++slot == a.nstage) {
                    slot = 0;
                    ++round;
                }
            }
        }
    } else {
        // ===== consumer warps
        const int col0 = q * dq + warp * CW;                       // first column of this warp
        // Y fragments of this warp's columns: B[k = col][n = lambda] -> lane holds Y[lambda fc][col k0 + fk]
        double yb[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) yb[k] = a.Y[static_cast<size_t>(fc) * a.ldv + col0 + 4 * k + fk];
        double acc[MB][2];
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[m][0] = acc[m][1] = 0.0;
        double nrm0 = 0.0, nrm1 = 0.0;   // sum of squares of R[fk][fc] and R[fk+4][fc]
        const uint32_t clu_base = mr_smem_u32(&sm.clu[0][0][0]);
        const uint32_t ubar_base = mr_smem_u32(&sm.ubar[0]);

        double b_pref = 0.0; // (CTA 0, warps 0-1) b of the row this thread publishes next
        if (q == 0 && warp < 2) {
            const long long row = lo + (tid >> 3);
            b_pref = (row < hi) ? __ldg(a.b + row) : 0.0;
        }
        int slot1 = 0;       // ring slot / round of the tile contract1 works on
        uint32_t round1 = 0;
        int slot2 = 0;       // ... and of the tile contract2 works on
        // first contraction of tile t: partial U of this warp's columns -> sm.red[t & 1][warp]
        auto contract1 = [&](int t) {
            mr_mbar_wait(&sm.full[slot1], round1 & 1u);
            const long long r0 = lo + static_cast<long long>(t) * MR_TILE;
            const int rows = static_cast<int>(min(static_cast<long long>(MR_TILE), hi - r0));
            const unsigned char* base = ring + static_cast<size_t>(slot1) * a.stage_bytes;
            const double* arow = reinterpret_cast<const double*>(base + static_cast<size_t>(fc) * a.pitch) + warp * CW + fk;
            double u[NCH][2];
#pragma unroll
            for (int c = 0; c < NCH; ++c) u[c][0] = u[c][1] = 0.0;
            const bool live = fc < rows;   // rows past the end of the block: the slot holds stale bytes there
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const double av = live ? arow[4 * k] : 0.0;
                mr_dmma(u[k % NCH][0], u[k % NCH][1], av, yb[k]);
            }
            // fixed pairwise tree over the chains
#pragma unroll
            for (int span = 1; span < NCH; span *= 2)
#pragma unroll
                for (int c = 0; c + span < NCH; c += 2 * span) {
                    u[c][0] += u[c + span][0];
                    u[c][1] += u[c + span][1];
                }
            const double u0 = u[0][0], u1 = u[0][1];
            *reinterpret_cast<double2*>(&sm.red[t & 1][warp][fc * 8 + 2 * fk]) = make_double2(u0, u1);
            __syncwarp();
            if (lane == 0) mr_mbar_arrive(&sm.redbar[t & 1]);
            if (++slot1 == a.nstage) {
                slot1 = 0;
                ++round1;
            }
        };
        // warps 0 and 1 (64 threads, one per entry of the 8 x 8 tile): ordered sum over the warps, CTA 0 folds in
        // -b, then one st.async per cluster CTA (data + byte count on that CTA's barrier)
        auto publish = [&](int t) {
            mr_mbar_wait(&sm.redbar[t & 1], static_cast<uint32_t>(t >> 1) & 1u);
            const int buf = t % MR_UBUF;
            if (tid == 0) mr_mbar_expect_tx(&sm.ubar[buf], CL * 64 * 8);
            double v = sm.red[t & 1][0][tid];
#pragma unroll
            for (int w = 1; w < 8; ++w) v += sm.red[t & 1][w][tid];
            if (q == 0) {   // b of this tile was requested one tile ago; request the next one now
                v -= b_pref;
                const long long row = lo + static_cast<long long>(t + 1) * MR_TILE + (tid >> 3);
                b_pref = (row < hi) ? __ldg(a.b + row) : 0.0;
            }
            const uint32_t off = static_cast<uint32_t>(((buf * MR_CLMAX + q) * 64 + tid) * 8);
#pragma unroll
            for (int r = 0; r < CL; ++r)
                mr_st_async(mr_mapa(clu_base + off, r), v, mr_mapa(ubar_base + buf * 8, r));
        };
        // second contraction of tile t: R = sum_q (U_q [- b]) in rank order, accumulators += A_tile^T R
        auto contract2 = [&](int t) {
            const int buf = t % MR_UBUF;
            mr_mbar_wait(&sm.ubar[buf], static_cast<uint32_t>(t / MR_UBUF) & 1u);
            const long long r0 = lo + static_cast<long long>(t) * MR_TILE;
            const int rows = static_cast<int>(min(static_cast<long long>(MR_TILE), hi - r0));
            double rlo, rhi;
            {
                const int i0 = fk * 8 + fc, i1 = (fk + 4) * 8 + fc;
                double s0 = sm.clu[buf][0][i0], s1 = sm.clu[buf][0][i1];
#pragma unroll
                for (int r = 1; r < CL; ++r) {
                    s0 += sm.clu[buf][r][i0];
                    s1 += sm.clu[buf][r][i1];
                }
                rlo = (fk < rows) ? s0 : 0.0;
                rhi = (fk + 4 < rows) ? s1 : 0.0;
            }
            nrm0 = fma(rlo, rlo, nrm0);
            nrm1 = fma(rhi, rhi, nrm1);
            if (a.grad) {
                const unsigned char* base = ring + static_cast<size_t>(slot2) * a.stage_bytes;
                const double* alo = reinterpret_cast<const double*>(base + static_cast<size_t>(fk) * a.pitch) + warp * CW + fc;
                const double* ahi = reinterpret_cast<const double*>(base + static_cast<size_t>(fk + 4) * a.pitch) + warp * CW + fc;
                const bool live_lo = fk < rows, live_hi = fk + 4 < rows;
                // rows 0-3 for every column block, then rows 4-7: the two MMAs of one accumulator are MB
                // instructions apart (issued back to back the second one waits out the full DMMA latency)
#pragma unroll
                for (int m = 0; m < MB; ++m) mr_dmma(acc[m][0], acc[m][1], live_lo ? alo[8 * m] : 0.0, rlo);
#pragma unroll
                for (int m = 0; m < MB; ++m) mr_dmma(acc[m][0], acc[m][1], live_hi ? ahi[8 * m] : 0.0, rhi);
            }
            __syncwarp();
            if (lane == 0) mr_mbar_arrive(&sm.empty[slot2]);   // this warp is done with the tile's slot
            if (++slot2 == a.nstage) slot2 = 0;
        };

        if (ntile > 0) {
            contract1(0);
            if (warp < 2) publish(0);
        }
        for (int t = 0; t < ntile; ++t) {
            if (t + 1 < ntile) {
                contract1(t + 1);
                if (warp < 2) publish(t + 1);
            }
            contract2(t);
        }

        // ---- results: C fragment -> G[col m0 + fc][lambda 2 fk + {0,1}]
        if (a.grad) {
            double* out = a.part + static_cast<size_t>(cl) * MR_NB * a.ldv;
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                const int col = col0 + 8 * m + fc;
                out[static_cast<size_t>(2 * fk) * a.ldv + col] = acc[m][0];
                out[static_cast<size_t>(2 * fk + 1) * a.ldv + col] = acc[m][1];
            }
        }
        if (q == 0 && warp == 0) {
            // lanes with the same fc hold rows fk and fk+4 of column lambda = fc: sum over fk (lane bits 0,1)
            double v = nrm0 + nrm1;
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (fk == 0) a.norms[static_cast<size_t>(cl) * MR_NB + fc] = v;
        }
    }
    cluster.sync();   // nobody exits while a peer may still write into its shared memory
}

struct MrhsUpdateArgs {
    const double* part;    // [ncl][8][ldv]
    const double* Yin;     // [8][ldv]
    double* Yout;          // [8][ldv]
    double* X;             // [8][ldv] in: x_k, out: x_{k+1}
    const double* alpha1;  // [8]
    double alpha2, tau, beta;
    int d, ldv, ncl;
    double* step_part;     // [gridDim.x][8] partial sums of (x+ - x)^2 (nullable)
};

// one thread per (column, penalty); the cluster partials are summed in cluster order
__global__ void __launch_bounds__(256) mrhs_update_kernel(const MrhsUpdateArgs u) {
    const int l = threadIdx.x >> 5;          // penalty = warp
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;
    double dx2 = 0.0;
    if (col < u.d) {
        const size_t idx = static_cast<size_t>(l) * u.ldv + col;
        double g = 0.0;
        for (int c = 0; c < u.ncl; ++c) g += u.part[(static_cast<size_t>(c) * MR_NB + l) * u.ldv + col];
        const double y = u.Yin[idx], xk = u.X[idx];
        if (u.alpha2 > 0.0) g = __dadd_rn(g, __dmul_rn(u.alpha2, y));
        double v = __dsub_rn(y, __dmul_rn(u.tau, g));
        const double a1 = u.alpha1[l];
        if (a1 > 0.0) v = fos_soft_threshold(v, __dmul_rn(u.tau, a1));
        u.X[idx] = v;
        u.Yout[idx] = __dadd_rn(v, __dmul_rn(u.beta, __dsub_rn(v, xk)));
        const double dx = v - xk;
        dx2 = dx * dx;
    }
    if (u.step_part != nullptr) {
        dx2 = fos_warp_sum(dx2);
        if (lane == 0) u.step_part[static_cast<size_t>(blockIdx.x) * MR_NB + l] = dx2;
    }
}

// objective of the 8 columns of X: 0.5 sum_c norms[c][l] (+0.5 a2 |x|^2) (+a1[l] |x|_1); one warp per penalty
__global__ void __launch_bounds__(256) mrhs_obj_kernel(const double* __restrict__ norms, int ncl, const double* __restrict__ X,
                                                       int d, int ldv, const double* __restrict__ alpha1, double alpha2,
                                                       double* __restrict__ obj) {
    const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double l1 = 0.0, l2 = 0.0;
    for (int c = lane; c < d; c += 32) {
        const double x = X[static_cast<size_t>(l) * ldv + c];
        l1 += fabs(x);
        l2 = fma(x, x, l2);
    }
    l1 = fos_warp_sum(l1);
    l2 = fos_warp_sum(l2);
    if (lane == 0) {
        double s = 0.0;
        for (int c = 0; c < ncl; ++c) s += norms[static_cast<size_t>(c) * MR_NB + l];
        double v = 0.5 * s;
        if (alpha2 > 0.0) v += 0.5 * alpha2 * l2;
        if (alpha1[l] > 0.0) v += alpha1[l] * l1;
        obj[l] = v;
    }
}

__global__ void mrhs_step_finish_kernel(const double* __restrict__ part, int nblk, int n_live, double* __restrict__ out) {
    // max over the live penalties of ||x+ - x||_2 (blocks summed in order); 8 threads
    __shared__ double red[8];
    const int l = threadIdx.x;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[static_cast<size_t>(b) * MR_NB + l];
    red[l] = (l < n_live) ? sqrt(s) : 0.0;
    __syncthreads();
    if (l == 0) {
        double m = 0.0;
        for (int i = 0; i < 8; ++i) m = fmax(m, red[i]);
        out[0] = fmax(out[0], m);
    }
}

const void* mrhs_kernel_for(int cw, int cl) {
    if (cl == 8) return (cw == 64) ? reinterpret_cast<const void*>(&mrhs_stream_kernel<64, 8>) : nullptr;
    switch (cw) {
        case 128: return reinterpret_cast<const void*>(&mrhs_stream_kernel<128>);
        case 64: return reinterpret_cast<const void*>(&mrhs_stream_kernel<64>);
        case 32: return reinterpret_cast<const void*>(&mrhs_stream_kernel<32>);
        default: return nullptr;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side: fos_mrhs_fista (include/fos.h)
// ------------------------------------------------------------------------------------------
extern "C" int fos_mrhs_fista(fos_design* h, const fos_path_params* pp, fos_path_result* pr) {
    FOS_REQUIRE(h && pp && pr, "null pointer argument");
    FOS_REQUIRE(pp->n_lambda >= 1 && pp->alphas1 != nullptr, "at least one penalty is required");
    FOS_REQUIRE(pp->step > 0.0 && pp->max_iter >= 0, "step must be positive, max_iter >= 0");
    FOS_REQUIRE(h->world == 1, "the streaming multi-RHS mode runs on one GPU (row-sharded designs: use the Gram mode)");
    if (h->dtype != FOS_F64 || h->lda != h->d || h->d % 1024 != 0 || h->d > 4096 || (h->d / 32 != 32 && h->d / 32 != 64 && h->d / 32 != 128)) {
        fos_set_error("streaming multi-RHS mode needs a dense float64 design with d in {1024, 2048, 4096} (got d = %d)", h->d);
        return FOS_ERR_UNSUPPORTED;
    }
    FOS_CUDA(cudaSetDevice(h->device));
    const int d = h->d, ldv = h->ldv;
    // CTAs per row block.  At d = 4096 a cluster of 4 leaves room for three 64 KB tiles only, and since a tile's slot is
    // free only after its SECOND contraction, one tile is in flight per CTA: the pass runs at the load latency
    // (5.7 ms at 500k rows).  A cluster of 8 halves the tile (six 32 KB slots, four in flight) at the price of SM
    // coverage (as many clusters of 8 as are co-resident, 16 on this part = 128 of 148 SMs): 5.0 ms, measured, so
    // d = 4096 uses it; FOS_MRHS_CLUSTER=4 forces quarters.
    const int n_lambda = pp->n_lambda;
    const int nbatch = (n_lambda + MR_NB - 1) / MR_NB;
    const int Lpad = nbatch * MR_NB;
    int CL = MR_CLUSTER, CW = 0, dq = 0, pitch = 0, stage_bytes = 0, nstage = 0, ncl = 0;
    const void* fn = nullptr;
    auto plan = [&](int cl_want) -> int {
        CL = cl_want;
        CW = d / CL / 8;
        fn = mrhs_kernel_for(CW, CL);
        FOS_REQUIRE(fn != nullptr, "no multi-RHS kernel for d = %d with clusters of %d", d, CL);
        dq = d / CL;
        pitch = dq * 8 + MR_PAD;
        stage_bytes = (MR_TILE * pitch + 127) & ~127;
        nstage = std::min(MR_MAXST, (200 * 1024) / stage_bytes);
        FOS_REQUIRE(nstage >= 2, "row too wide for the multi-RHS ring");
        FOS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, nstage * stage_bytes));
        ncl = std::max(1, h->sm_count / CL);
        if (CL > MR_CLUSTER) {
            // every cluster must be resident at once: a second wave would double the pass
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(ncl * CL);
            cfg.blockDim = dim3(MR_THREADS);
            cfg.dynamicSmemBytes = static_cast<size_t>(nstage) * stage_bytes;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CL;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg) != cudaSuccess) {
                cudaGetLastError();
                max_clusters = 0;
            }
            if (max_clusters < 8) return FOS_ERR_UNSUPPORTED;   // not placeable / not worth it: the caller falls back
            ncl = std::min(ncl, max_clusters);
        }
        return FOS_OK;
    };
    {
        int want = (d == 4096) ? 8 : MR_CLUSTER;
        if (const char* e = getenv("FOS_MRHS_CLUSTER")) {
            if (atoi(e) == 4) want = MR_CLUSTER;
        }
        int st = plan(want);
        if (st == FOS_ERR_UNSUPPORTED && want != MR_CLUSTER) st = plan(MR_CLUSTER);
        FOS_TRY(st);
    }

    // row blocks: multiples of MR_TILE rows, the last cluster takes the tail
    std::vector<long long> row_lo(ncl + 1);
    for (int c = 0; c <= ncl; ++c) row_lo[c] = std::min<long long>(h->n, ((h->n * c / ncl) + MR_TILE - 1) / MR_TILE * MR_TILE);
    row_lo[0] = 0;
    row_lo[ncl] = h->n;

    const size_t vec = static_cast<size_t>(Lpad) * ldv * sizeof(double);
    double *Y0 = nullptr, *Y1 = nullptr, *X = nullptr, *a1 = nullptr, *part = nullptr, *norms = nullptr, *obj = nullptr,
           *spart = nullptr, *smax = nullptr;
    long long* rl = nullptr;
    double* smax_host = nullptr;
    const int ublk = (d + 31) / 32;
    auto cleanup = [&]() {
        for (void* p : {static_cast<void*>(Y0), static_cast<void*>(Y1), static_cast<void*>(X), static_cast<void*>(a1),
                        static_cast<void*>(part), static_cast<void*>(norms), static_cast<void*>(obj),
                        static_cast<void*>(spart), static_cast<void*>(smax), static_cast<void*>(rl), static_cast<void*>(smax_host)})
            fos_pool_free(p);
    };
    cudaStream_t s = h->stream;
    auto body = [&]() -> int {
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&Y0), vec));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&Y1), vec));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&X), vec));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&a1), Lpad * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&part), static_cast<size_t>(ncl) * MR_NB * ldv * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&norms), static_cast<size_t>(ncl) * MR_NB * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&obj), Lpad * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&spart), static_cast<size_t>(ublk) * MR_NB * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&smax), sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&rl), (ncl + 1) * sizeof(long long)));
        FOS_CUDA(fos_pool_malloc_host(reinterpret_cast<void**>(&smax_host), sizeof(double)));
        FOS_CUDA(cudaMemsetAsync(Y0, 0, vec, s));
        FOS_CUDA(cudaMemsetAsync(Y1, 0, vec, s));
        FOS_CUDA(cudaMemsetAsync(X, 0, vec, s));
        FOS_CUDA(cudaMemsetAsync(a1, 0, Lpad * sizeof(double), s));
        FOS_CUDA(cudaMemcpyAsync(a1, pp->alphas1, n_lambda * sizeof(double), cudaMemcpyHostToDevice, s));
        FOS_CUDA(cudaMemcpyAsync(rl, row_lo.data(), (ncl + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
        if (pp->X0) {
            FOS_CUDA(cudaMemcpy2DAsync(X, ldv * sizeof(double), pp->X0, d * sizeof(double), d * sizeof(double), n_lambda,
                                       cudaMemcpyHostToDevice, s));
            FOS_CUDA(cudaMemcpyAsync(Y0, X, vec, cudaMemcpyDeviceToDevice, s));
        }
        MrhsArgs ma{};
        ma.A = static_cast<const double*>(h->A);
        ma.b = h->b;
        ma.row_lo = rl;
        ma.d = d;
        ma.lda = h->lda;
        ma.ldv = ldv;
        ma.nstage = nstage;
        ma.stage_bytes = stage_bytes;
        ma.pitch = pitch;
        ma.part = part;
        ma.norms = norms;
        ma.prof = nullptr;
        auto launch_pass = [&](const double* Yb, int grad) -> int {
            ma.Y = Yb;
            ma.grad = grad;
            void* params[1] = {&ma};
            FOS_CUDA(fos_launch_ex(fn, dim3(ncl * CL), dim3(MR_THREADS), static_cast<size_t>(nstage) * stage_bytes, s, params,
                                   false, 0));
            h->launches++;
            return FOS_OK;
        };
        const long long launches0 = h->launches;
        FOS_CUDA(cudaEventRecord(h->ev0, s));
        double t_prev = 1.0;
        double* Yin = Y0;
        double* Yout = Y1;
        int it = 0;
        double last_max = 0.0;
        const int check = std::max(1, pp->check_every);
        for (; it < pp->max_iter; ++it) {
            const double t_cur = 0.5 * (1.0 + sqrt(1.0 + 4.0 * (t_prev * t_prev)));
            const double beta = (t_prev - 1.0) / t_cur;
            const bool chk = pp->tol > 0.0 && ((it + 1) % check == 0);
            if (chk) FOS_CUDA(cudaMemsetAsync(smax, 0, sizeof(double), s));
            for (int bt = 0; bt < nbatch; ++bt) {
                const size_t off = static_cast<size_t>(bt) * MR_NB * ldv;
                FOS_TRY(launch_pass(Yin + off, 1));
                MrhsUpdateArgs ua{};
                ua.part = part;
                ua.Yin = Yin + off;
                ua.Yout = Yout + off;
                ua.X = X + off;
                ua.alpha1 = a1 + bt * MR_NB;
                ua.alpha2 = pp->alpha2;
                ua.tau = pp->step;
                ua.beta = beta;
                ua.d = d;
                ua.ldv = ldv;
                ua.ncl = ncl;
                ua.step_part = chk ? spart : nullptr;
                mrhs_update_kernel<<<dim3(ublk), dim3(256), 0, s>>>(ua);
                h->launches++;
                if (chk) {
                    mrhs_step_finish_kernel<<<1, 8, 0, s>>>(spart, ublk, std::min(MR_NB, n_lambda - bt * MR_NB), smax);
                    h->launches++;
                }
            }
            FOS_CUDA(cudaGetLastError());
            std::swap(Yin, Yout);
            t_prev = t_cur;
            if (chk) {
                FOS_CUDA(cudaMemcpyAsync(smax_host, smax, sizeof(double), cudaMemcpyDeviceToHost, s));
                FOS_CUDA(cudaStreamSynchronize(s));
                last_max = *smax_host;
                if (last_max < pp->tol) {
                    ++it;
                    break;
                }
            }
        }
        // objectives of the final iterates: one norms-only pass per batch
        for (int bt = 0; bt < nbatch; ++bt) {
            const size_t off = static_cast<size_t>(bt) * MR_NB * ldv;
            FOS_TRY(launch_pass(X + off, 0));
            mrhs_obj_kernel<<<1, 256, 0, s>>>(norms, ncl, X + off, d, ldv, a1 + bt * MR_NB, pp->alpha2, obj + bt * MR_NB);
            h->launches++;
        }
        FOS_CUDA(cudaGetLastError());
        FOS_CUDA(cudaEventRecord(h->ev1, s));
        FOS_CUDA(cudaStreamSynchronize(s));
        if (pr->X)
            FOS_CUDA(cudaMemcpy2D(pr->X, d * sizeof(double), X, ldv * sizeof(double), d * sizeof(double), n_lambda,
                                  cudaMemcpyDeviceToHost));
        if (pr->obj) FOS_CUDA(cudaMemcpy(pr->obj, obj, n_lambda * sizeof(double), cudaMemcpyDeviceToHost));
        FOS_CUDA(cudaEventElapsedTime(&pr->loop_ms, h->ev0, h->ev1));
        pr->kernel_launches = h->launches - launches0;
        pr->n_iters = it;
        pr->last_max_step = last_max;
        pr->tile_rows = MR_TILE;
        return FOS_OK;
    };
    const int st = body();
    if (st != FOS_OK) cudaStreamSynchronize(s);
    cleanup();
    return st;
}
