// gram_kernels.cu -- Gram-matrix mode for n >> d and the batched multi-lambda path
// (north_star item 4; BASELINE.json configs[4]).  These are the only dense contractions on the
// path and the only place tensor cores are used: fp64 DMMA (mma.sync.m8n8k4.f64 -- tcgen05 has
// no f64 kind).
//
//   gram_syrk_kernel   G_partial[s] = A[k-range s]^T A[k-range s]   (upper-triangular 128x128 tiles)
//   gram_reduce_kernel G = sum_s G_partial[s], mirrored to the lower triangle (fixed order)
//   path_step_kernel   one FISTA iteration for Lambda penalties at once:
//                      Grad = G Y - c 1^T (+a2 Y);  X+ = prox(Y - tau Grad, tau a1[l]);
//                      Y+ = X+ + beta (X+ - X)          (iterative_solvers.py:173-221, batched)
//   path_obj_kernel    per-column  0.5 x^T G x - c^T x + 0.5 b^T b (+0.5 a2 |x|^2) (+a1 |x|_1)
//
// Operands are staged by the TMA unit (2-D tensor maps, 128-byte swizzle, mbarrier hand-over; the default) or with
// cp.async (16 B, L1 bypass; FOS_GRAM_TMA=0 / FOS_PATH_TMA=0) into a 4-stage shared-memory ring whose
// row pitch is padded by 4 doubles, which makes every DMMA fragment load conflict free.
// Bound: the fp64 tensor pipe (n d^2 FMA for the build, d^2 Lambda per path iteration).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include <cuda.h>  // CUtensorMap (the encoder is resolved at run time, no link against libcuda)

#include "fos_common.cuh"

namespace {

constexpr int GT = 128;       // Gram tile edge
constexpr int GKB = 16;       // rows of A per pipeline stage
constexpr int GLD = GT + 4;   // padded shared row pitch (doubles)
constexpr int GSTAGES = 4;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
// 16-byte async copy global -> shared; src_bytes == 0 zero-fills (used for rows past the end)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct SyrkArgs {
    const double* A;
    long long n;
    int lda;
    int nb;                  // d / 128
    int ntiles;              // nb (nb+1) / 2
    long long rows_per_split;  // multiple of GKB
    double* W;               // [nsplit][dpad][dpad]
    long long dpad;
    int accumulate;          // 1: W[split] += tile (row chunks arriving one after the other)
};

__global__ void __launch_bounds__(256, 1) gram_syrk_kernel(const SyrkArgs g) {
    extern __shared__ __align__(16) double smem[];  // [GSTAGES][2][GKB][GLD]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x / g.ntiles;
    int rem = blockIdx.x % g.ntiles, bi = 0;
    while (rem >= g.nb - bi) {
        rem -= g.nb - bi;
        ++bi;
    }
    const int bj = bi + rem;
    const long long k_lo = split * g.rows_per_split;
    const long long k_hi = min(g.n, k_lo + g.rows_per_split);
    const int nst = static_cast<int>((k_hi - k_lo + GKB - 1) / GKB);

    // cp.async assignment: 2 operands x 16 rows x 64 chunks of 16 B = 2048 chunks, 8 per thread
    auto issue = [&](int st) {
        if (st < nst) {
            double* base = smem + static_cast<size_t>(st % GSTAGES) * (2 * GKB * GLD);
            const long long r0 = k_lo + static_cast<long long>(st) * GKB;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int chunk = tid + 256 * q;       // 0..2047
                const int op = chunk >> 10;            // 0: column block bi, 1: column block bj
                const int row = (chunk >> 6) & 15;
                const int c16 = chunk & 63;            // 16-byte chunk within the 1 KB row segment
                const long long r = r0 + row;
                const bool ok = r < k_hi;
                const double* src = g.A + (ok ? r : k_lo) * g.lda + (op ? bj : bi) * GT + c16 * 2;
                cp_async16(base + (op * GKB + row) * GLD + c16 * 2, src, ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int wi = warp >> 1, wj = warp & 1;  // 4 x 2 warps, warp tile 32 (i) x 64 (j)
    const int fk = lane & 3, fc = lane >> 2;

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) issue(s);
    for (int s = 0; s < nst; ++s) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        issue(s + GSTAGES - 1);
        const double* tI = smem + static_cast<size_t>(s % GSTAGES) * (2 * GKB * GLD);
        const double* tJ = tI + GKB * GLD;
#pragma unroll
        for (int k4 = 0; k4 < GKB; k4 += 4) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tI[(k4 + fk) * GLD + wi * 32 + i * 8 + fc];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = tJ[(k4 + fk) * GLD + wj * 64 + j * 8 + fc];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    double* out = g.W + static_cast<size_t>(split) * g.dpad * g.dpad;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long row = static_cast<long long>(bi) * GT + wi * 32 + i * 8 + fc;
            const long long col = static_cast<long long>(bj) * GT + wj * 64 + j * 8 + fk * 2;
            double2* dst = reinterpret_cast<double2*>(out + row * g.dpad + col);
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            if (g.accumulate) {  // this (tile, split) slot belongs to this CTA alone; chunks are stream ordered
                const double2 old = *dst;
                v.x += old.x;
                v.y += old.y;
            }
            *dst = v;
        }
}

// ------------------------------------------------------------------------------------------
// The same tile product with the operands staged by the TMA unit (SASS: UTMALDG) instead of 8 LDGSTS per
// thread and stage.  One elected lane issues 16 two-dimensional tensor copies per stage -- boxes of 16 columns
// x 16 rows, SWIZZLE_128B; the eight warps take that turn round robin (a ninth, dedicated producer warp would
// cap the kernel at 168 registers: three warps on one scheduler) -- and the warps never meet at a CTA barrier:
// a stage is handed over by its `full` mbarrier (byte count) and returned by its `empty` mbarrier (one
// arrival per warp).
//   shared layout of one operand and stage: [8 column groups][16 rows][128 B], the 16-byte chunk c of row r stored
//   at chunk c ^ (r % 8).  A DMMA step takes the four rows {0,1,4,5} + 2 (t & 1) + 8 (t >> 1) of the stage (the
//   order of the k index inside a product is free as long as both operands use the same one): rows whose index
//   differs in bit 2 land in the two halves of the 128-byte line, so every fragment load is the minimum of two
//   wavefronts (tests/test_syrk_tma_layout_cpu.py replays the addressing).
// Rows past the end of the split are zero-filled by the copy unit (tensor extent = rows of this call).
// ------------------------------------------------------------------------------------------
constexpr int TS_STAGES = 6;
constexpr int TS_AHEAD = TS_STAGES - 2;  // stages requested ahead of the one being multiplied
constexpr int TS_OP_BYTES = GKB * GT * 8;        // 16 KB: one operand of one stage
constexpr int TS_STAGE_BYTES = 2 * TS_OP_BYTES;  // 32 KB
constexpr int TS_BOX_BYTES = GKB * 16 * 8;       // 2 KB: one box
constexpr int TS_SMEM_BYTES = TS_STAGES * TS_STAGE_BYTES + 2 * TS_STAGES * 8 + 1024;

__device__ __forceinline__ void ts_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ts_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TS_DONE;\n\t"
        "bra TS_WAIT;\n\t"
        "TS_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// one box of the 2-D tensor (column, row) -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void ts_tma_2d(uint32_t dst, const CUtensorMap* map, int col, int row, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(bar)
        : "memory");
}

__global__ void __launch_bounds__(256, 1) gram_syrk_tma_kernel(const SyrkArgs g, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ unsigned char ts_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = static_cast<uint32_t>(__cvta_generic_to_shared(ts_raw));
    const uint32_t ring = (raw + 1023u) & ~1023u;  // SWIZZLE_128B repeats every 1024 bytes of shared address
    const unsigned char* ring_p = ts_raw + (ring - raw);
    const uint32_t full0 = ring + TS_STAGES * TS_STAGE_BYTES, empty0 = full0 + 8 * TS_STAGES;

    const int split = blockIdx.x / g.ntiles;
    int rem = blockIdx.x % g.ntiles, bi = 0;
    while (rem >= g.nb - bi) {
        rem -= g.nb - bi;
        ++bi;
    }
    const int bj = bi + rem;
    const long long k_lo = split * g.rows_per_split;
    const long long k_hi = min(g.n, k_lo + g.rows_per_split);
    const int nst = (k_hi > k_lo) ? static_cast<int>((k_hi - k_lo + GKB - 1) / GKB) : 0;

    // request stage sp (one lane): its slot was last used by stage sp - TS_STAGES
    auto produce = [&](int sp) {
        if (sp < nst) {
            const int slot = sp % TS_STAGES, turn = sp / TS_STAGES;
            if (turn > 0) ts_mbar_wait(empty0 + 8 * slot, (turn - 1) & 1);
            const uint32_t bar = full0 + 8 * slot;
            ts_mbar_expect_tx(bar, TS_STAGE_BYTES);
            const uint32_t dst = ring + slot * TS_STAGE_BYTES;
            const int row = static_cast<int>(k_lo + static_cast<long long>(sp) * GKB);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ts_tma_2d(dst + q * TS_BOX_BYTES, &tmap, bi * GT + q * 16, row, bar);
                ts_tma_2d(dst + TS_OP_BYTES + q * TS_BOX_BYTES, &tmap, bj * GT + q * 16, row, bar);
            }
        }
    };

    if (tid == 0) {
        for (int s = 0; s < TS_STAGES; ++s) {
            ts_mbar_init(full0 + 8 * s, 1);
            ts_mbar_init(empty0 + 8 * s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < TS_AHEAD; ++s) produce(s);
    }
    __syncthreads();

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int wi = warp >> 1, wj = warp & 1;  // 4 x 2 warps, warp tile 32 (i) x 64 (j)
    const int fk = lane & 3, fc = lane >> 2;
    // byte offset of this lane's element inside a box, for the two row quartets (t & 1) and the two 8-column
    // halves of a box
    int loff[2][2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int x = p * 2 + (fk & 1) + (fk >> 1) * 4;  // row % 8
        const int y = (fc >> 1) ^ x;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) loff[p][hh] = x * 128 + ((y ^ (hh * 4)) * 16) + (fc & 1) * 8;
    }
    const int a_base = wi * 2 * TS_BOX_BYTES;                // column groups 2 wi .. 2 wi + 1 of operand I
    const int b_base = TS_OP_BYTES + wj * 4 * TS_BOX_BYTES;  // column groups 4 wj .. 4 wj + 3 of operand J

    for (int s = 0; s < nst; ++s) {
        const int slot = s % TS_STAGES;
        // the warps take turns as producer: stage s + TS_AHEAD goes into the slot of stage s - 2, which every
        // warp has normally left long ago (the wait on its empty barrier then falls through)
        if (warp == (s & 7) && lane == 0) produce(s + TS_AHEAD);
        __syncwarp();
        ts_mbar_wait(full0 + 8 * slot, (s / TS_STAGES) & 1);
        const unsigned char* st = ring_p + static_cast<size_t>(slot) * TS_STAGE_BYTES;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a[i] = *reinterpret_cast<const double*>(st + a_base + (i >> 1) * TS_BOX_BYTES + (t >> 1) * 1024 +
                                                        loff[t & 1][i & 1]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                b[j] = *reinterpret_cast<const double*>(st + b_base + (j >> 1) * TS_BOX_BYTES + (t >> 1) * 1024 +
                                                        loff[t & 1][j & 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) ts_mbar_arrive(empty0 + 8 * slot);
    }

    double* out = g.W + static_cast<size_t>(split) * g.dpad * g.dpad;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long row = static_cast<long long>(bi) * GT + wi * 32 + i * 8 + fc;
            const long long col = static_cast<long long>(bj) * GT + wj * 64 + j * 8 + fk * 2;
            double2* dst = reinterpret_cast<double2*>(out + row * g.dpad + col);
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            if (g.accumulate) {
                const double2 old = *dst;
                v.x += old.x;
                v.y += old.y;
            }
            *dst = v;
        }
}

// G[i][j] = sum_s W[s][min-tile order] for the upper tile triangle, mirrored below it
__global__ void gram_reduce_kernel(const double* __restrict__ W, double* __restrict__ G, long long dpad,
                                   int nsplit) {
    const long long i = blockIdx.y;
    const long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= dpad) return;
    const long long ti = i / GT, tj = j / GT;
    long long si = i, sj = j;
    if (ti > tj) {  // below the tile diagonal: read the transposed element
        si = j;
        sj = i;
    }
    double s = 0.0;
    for (int p = 0; p < nsplit; ++p) s += W[static_cast<size_t>(p) * dpad * dpad + si * dpad + sj];
    G[i * dpad + j] = s;
}

// ------------------------------------------------------------------------------------------
// batched path iteration: out tile PM (i) x 64 (l), K chunks of 16.  PM = 128 when there are
// enough penalties to fill the SMs, 64 / 32 for short penalty lists (more, smaller tiles).
// ------------------------------------------------------------------------------------------
#ifndef FOS_PATH_PK
#define FOS_PATH_PK 16
#endif
#ifndef FOS_PATH_STAGES
#define FOS_PATH_STAGES 4
#endif
// k-chunk 16 x 4 stages; 32 x 3 stages (half as many block barriers) measured the same on the tile schedule and
// slower on the stream-K schedule (coarser ranges, longer pipeline prologue per segment): -DFOS_PATH_PK=32 -DFOS_PATH_STAGES=3
constexpr int PN = 64, PK = FOS_PATH_PK, PLD = PK + 4, PSTAGES = FOS_PATH_STAGES;
constexpr int PCPR = PK / 2;   // 16-byte chunks per tile row and stage

struct PathArgs {
    const double* G;      // [d][d] row-major, symmetric
    const double* c;      // A^T b
    const double* Yin;    // [Lpad][d]
    double* Yout;         // [Lpad][d]
    double* X;            // [Lpad][d]  (in: x_k, out: x_{k+1})
    const double* alpha1; // [Lpad]
    double alpha2, tau, beta;
    int d, Lpad;
    int mode;             // 0: FISTA step, 1: objective partials of X (Yin = X)
    int ymap;             // which tensor map describes Yin (PathMaps::y[ymap]; TMA-staged kernels only)
    double* obj_part;     // [d/PM][Lpad][4]: x^T G x, c^T x, |x|_1, |x|^2 partials per i-block
    double* step_part;    // [d/PM][Lpad]: sum_i (x+ - x)^2 partials (nullable: not a check iteration)
};

// Shared by the tile-per-CTA kernel and the stream-K kernel: the pipelined contraction of one
// (PM x PN) tile over the k-steps [ks_lo, ks_lo + nst) and the fused epilogue on a finished tile.
template <int PM, int TN = PN>
struct PathTile {
    static constexpr int WI = PM / 32;          // warps along i
    static constexpr int WL = 8 / WI;           // warps along l
    static constexpr int LW = TN / WL;          // l columns per warp
    static constexpr int NJ = LW / 8;           // mma blocks along l per warp
    static constexpr int CH = (PM + TN) * PCPR / 256;  // 16-byte chunks per thread and stage
};

template <int PM, int TN = PN>
__device__ __forceinline__ void path_mainloop(const PathArgs& p, double* smem, int i0, int l0, int ks_lo, int nst,
                                              double (&acc)[4][PathTile<PM, TN>::NJ][2]) {
    using PT = PathTile<PM, TN>;
    constexpr int WL = PT::WL, LW = PT::LW, NJ = PT::NJ, CH = PT::CH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto issue = [&](int st) {
        if (st < nst) {
            double* base = smem + static_cast<size_t>(st % PSTAGES) * ((PM + TN) * PLD);
            const int k0 = (ks_lo + st) * PK;
#pragma unroll
            for (int q = 0; q < CH; ++q) {
                const int chunk = tid + 256 * q;
                const int row = chunk / PCPR, c16 = chunk % PCPR;
                const double* src = (row < PM) ? p.G + static_cast<size_t>(i0 + row) * p.d + k0 + c16 * 2
                                               : p.Yin + static_cast<size_t>(l0 + row - PM) * p.d + k0 + c16 * 2;
                cp_async16(base + row * PLD + c16 * 2, src, 16);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int wi = warp / WL, wl = warp % WL;  // warp tile 32 (i) x LW (l)
    const int fk = lane & 3, fc = lane >> 2;

#pragma unroll
    for (int s = 0; s < PSTAGES - 1; ++s) issue(s);
    for (int s = 0; s < nst; ++s) {
        cp_async_wait<PSTAGES - 2>();
        __syncthreads();
        issue(s + PSTAGES - 1);
        const double* tG = smem + static_cast<size_t>(s % PSTAGES) * ((PM + TN) * PLD);
        const double* tY = tG + PM * PLD;
#pragma unroll
        for (int k4 = 0; k4 < PK; k4 += 4) {
            double a[4], b[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tG[(wi * 32 + i * 8 + fc) * PLD + k4 + fk];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = tY[(wl * LW + j * 8 + fc) * PLD + k4 + fk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // the ring may be refilled by the next tile of the same CTA
}

// ---- the same contraction with both operands staged by the TMA unit (compare gram_syrk_tma_kernel) ----------
// One box of PM rows x 16 k-values of G and one of TN rows x 16 k-values of Yin per stage, 128-byte swizzle: the
// 16-byte chunk c of tile row r sits at chunk c ^ (r % 8) of the row's 128-byte line.  A fragment load reads the
// eight rows r0 + fc at k = k4 + fk: chunk ((k4 >> 1) + (fk >> 1)) ^ fc, eight different chunks for the eight
// rows, so the load is the minimum of two wavefronts without any padding.  The k order inside a product is the
// cp.async kernel's: both stagings give the same bits.
// The mbarrier phases run on across the calls of one CTA (stream-K: several segments), tracked in PathRing::gs.
struct PathMaps {
    CUtensorMap g;      // G as (k, i), box 16 x PM
    CUtensorMap y[3];   // Y0, Y1, X as (k, l), box 16 x TN
};
struct PathRing {
    uint32_t ring, full0, empty0;
    const unsigned char* ring_p;
    int gs;             // stages consumed so far by this CTA
};
template <int PM, int TN>
struct PathTma {
    static constexpr int STAGE_BYTES = (PM + TN) * 128;
    static constexpr int STAGES = (PM + TN >= 256) ? 5 : 6;
    static constexpr int AHEAD = STAGES - 2;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;
};
template <int PM, int TN>
__device__ __forceinline__ void path_ring_init(PathRing& r, unsigned char* raw_p) {
    using TM = PathTma<PM, TN>;
    const uint32_t raw = static_cast<uint32_t>(__cvta_generic_to_shared(raw_p));
    r.ring = (raw + 1023u) & ~1023u;
    r.ring_p = raw_p + (r.ring - raw);
    r.full0 = r.ring + TM::STAGES * TM::STAGE_BYTES;
    r.empty0 = r.full0 + 8 * TM::STAGES;
    r.gs = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TM::STAGES; ++s) {
            ts_mbar_init(r.full0 + 8 * s, 1);
            ts_mbar_init(r.empty0 + 8 * s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
template <int PM, int TN = PN>
__device__ __forceinline__ void path_mainloop_tma(const PathMaps& m, int ymap, PathRing& rg, int i0, int l0, int ks_lo,
                                                  int nst, double (&acc)[4][PathTile<PM, TN>::NJ][2]) {
    using PT = PathTile<PM, TN>;
    using TM = PathTma<PM, TN>;
    constexpr int WL = PT::WL, LW = PT::LW, NJ = PT::NJ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gs0 = rg.gs;
    auto produce = [&](int sl) {  // one lane: request stage sl of this call
        if (sl < nst) {
            const int g2 = gs0 + sl, slot = g2 % TM::STAGES, turn = g2 / TM::STAGES;
            if (turn > 0) ts_mbar_wait(rg.empty0 + 8 * slot, (turn - 1) & 1);
            const uint32_t bar = rg.full0 + 8 * slot;
            ts_mbar_expect_tx(bar, TM::STAGE_BYTES);
            const uint32_t dst = rg.ring + slot * TM::STAGE_BYTES;
            const int k0 = (ks_lo + sl) * 16;
            ts_tma_2d(dst, &m.g, k0, i0, bar);
            ts_tma_2d(dst + PM * 128, &m.y[ymap], k0, l0, bar);
        }
    };
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int wi = warp / WL, wl = warp % WL;  // warp tile 32 (i) x LW (l)
    const int fk = lane & 3, fc = lane >> 2;
    int koff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) koff[q] = fc * 128 + ((((fk >> 1) + 2 * q) ^ fc) * 16) + (fk & 1) * 8;
    const int a_base = wi * 32 * 128, b_base = PM * 128 + wl * LW * 128;

    if (tid == 0)
        for (int s = 0; s < TM::AHEAD; ++s) produce(s);
    for (int s = 0; s < nst; ++s) {
        const int g2 = gs0 + s, slot = g2 % TM::STAGES;
        if (warp == (s & 7) && lane == 0) produce(s + TM::AHEAD);
        __syncwarp();
        ts_mbar_wait(rg.full0 + 8 * slot, (g2 / TM::STAGES) & 1);
        const unsigned char* st = rg.ring_p + static_cast<size_t>(slot) * TM::STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double a[4], b[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double*>(st + a_base + i * 1024 + koff[q]);
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double*>(st + b_base + j * 1024 + koff[q]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) ts_mbar_arrive(rg.empty0 + 8 * slot);
    }
    rg.gs = gs0 + nst;
}

// iblk: index of the tile's i-block (row of the partial-sum arrays)
template <int PM, int TN = PN>
__device__ __forceinline__ void path_epilogue(const PathArgs& p, int i0, int l0, int iblk,
                                              double (&acc)[4][PathTile<PM, TN>::NJ][2],
                                              double (*part)[PathTile<PM, TN>::LW][4]) {
    using PT = PathTile<PM, TN>;
    constexpr int WI = PT::WI, WL = PT::WL, LW = PT::LW, NJ = PT::NJ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wi = warp / WL, wl = warp % WL;
    const int fk = lane & 3, fc = lane >> 2;
    // per-(l) sums of this thread: [0..3] objective pieces (mode 1) or [0] squared step (mode 0)
    double sums[NJ][2][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int q = 0; q < 4; ++q) sums[j][e][q] = 0.0;

    if (p.mode == 0) {
        // fused FISTA update on the tile (same roundings as the single-lambda epilogue).  The loads of a batch of
        // JB x 2 elements are issued together, ahead of the batch's stores: Yout / X may alias Yin / X as far as the
        // compiler can tell, so an element-by-element loop is one exposed load latency per element (64 of them per
        // thread on a 128 x 128 tile: ~50 us per tile, measured, on the critical path of the iteration)
        constexpr int JB = (NJ < 4) ? NJ : 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jb = 0; jb < NJ; jb += JB) {
                double yv[JB][2], xv[JB][2], a1v[JB][2];
                const double cv = __ldg(p.c + i0 + wi * 32 + i * 8 + fc);
#pragma unroll
                for (int jj = 0; jj < JB; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = i0 + wi * 32 + i * 8 + fc;
                        const int l = l0 + wl * LW + (jb + jj) * 8 + fk * 2 + e;
                        const size_t idx = static_cast<size_t>(l) * p.d + col;
                        yv[jj][e] = __ldcg(p.Yin + idx);
                        xv[jj][e] = __ldcg(p.X + idx);
                        a1v[jj][e] = __ldg(p.alpha1 + l);
                    }
#pragma unroll
                for (int jj = 0; jj < JB; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = jb + jj;
                        const int col = i0 + wi * 32 + i * 8 + fc;
                        const int l = l0 + wl * LW + j * 8 + fk * 2 + e;
                        const size_t idx = static_cast<size_t>(l) * p.d + col;
                        const double y = yv[jj][e], xk = xv[jj][e];
                        double g = __dsub_rn(acc[i][j][e], cv);
                        if (p.alpha2 > 0.0) g = __dadd_rn(g, __dmul_rn(p.alpha2, y));
                        double v = __dsub_rn(y, __dmul_rn(p.tau, g));
                        const double a1 = a1v[jj][e];
                        if (a1 > 0.0) v = fos_soft_threshold(v, __dmul_rn(p.tau, a1));
                        p.X[idx] = v;
                        p.Yout[idx] = __dadd_rn(v, __dmul_rn(p.beta, __dsub_rn(v, xk)));
                        const double dx = v - xk;
                        sums[j][e][0] = fma(dx, dx, sums[j][e][0]);
                    }
            }
        if (p.step_part == nullptr) return;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = i0 + wi * 32 + i * 8 + fc;
                    const int l = l0 + wl * LW + j * 8 + fk * 2 + e;
                    const double x = p.Yin[static_cast<size_t>(l) * p.d + col];
                    sums[j][e][0] = fma(x, acc[i][j][e], sums[j][e][0]);
                    sums[j][e][1] = fma(x, p.c[col], sums[j][e][1]);
                    sums[j][e][2] += fabs(x);
                    sums[j][e][3] = fma(x, x, sums[j][e][3]);
                }
    }
    // reduce over the 8 lanes that share fk (lane>>2 varies): xor 4, 8, 16
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double v = sums[j][e][q];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (fc == 0) part[warp][j * 8 + fk * 2 + e][q] = v;
            }
    __syncthreads();
    // combine the WI i-warps of each l, fixed order; TN l x 4 sums over the 256 threads
    for (int item = tid; item < TN * 4; item += 256) {
        const int l_loc = item >> 2, q = item & 3;
        const int wl2 = l_loc / LW, lw = l_loc % LW;
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < WI; ++w) t += part[w * WL + wl2][lw][q];
        if (p.mode == 0) {
            if (q == 0) p.step_part[static_cast<size_t>(iblk) * p.Lpad + l0 + l_loc] = t;
        } else {
            p.obj_part[(static_cast<size_t>(iblk) * p.Lpad + l0 + l_loc) * 4 + q] = t;
        }
    }
}

template <int PM, bool TMA>
__global__ void __launch_bounds__(256, 1) path_step_kernel(const PathArgs p, const __grid_constant__ PathMaps maps) {
    extern __shared__ __align__(16) double smem[];  // cp.async: [PSTAGES][(PM+PN)][PLD]; TMA: swizzled ring + mbarriers
    __shared__ double part[8][PathTile<PM>::LW][4];
    const int i0 = blockIdx.x * PM, l0 = blockIdx.y * PN;
    double acc[4][PathTile<PM>::NJ][2];
    if constexpr (TMA) {
        PathRing rg;
        path_ring_init<PM, PN>(rg, reinterpret_cast<unsigned char*>(smem));
        path_mainloop_tma<PM>(maps, p.ymap, rg, i0, l0, 0, p.d / 16, acc);
    } else {
        path_mainloop<PM>(p, smem, i0, l0, 0, p.d / PK, acc);
    }
    path_epilogue<PM>(p, i0, l0, blockIdx.x, acc, part);
}

// ------------------------------------------------------------------------------------------
// Stream-K schedule of the same iteration (128 x 64 tiles): the T tiles x d/16 k-steps are cut into P
// equal contiguous ranges, one per CTA (P = number of SMs), so every SM works the whole time whatever
// the tile count -- 128 tiles on 148 SMs at 256 penalties, 32 tiles at the 32 penalties a rank holds
// on 8 GPUs.  A CTA whose range covers a whole tile finishes it in registers.  A partial range is
// written to the workspace and counted on the tile; the LAST contributor to arrive (atomic ticket, no
// waiting, no co-residency assumption) adds the contributors' partials in ascending CTA order -- the
// same order whoever arrives last, so results stay bit-reproducible -- and runs the fused epilogue.
// ------------------------------------------------------------------------------------------
struct PathSkArgs {
    PathArgs p;
    double* W;            // [P][2][128 x TN] double: partial accumulators (fragment layout), 2 slots per CTA
    unsigned* ticket;     // [T] arrivals per tile (reset by the finishing CTA)
    int T, KT, P;         // tiles, k-steps per tile, CTAs
};

__device__ __forceinline__ long long sk_begin(const PathSkArgs& a, int c) {
    return (static_cast<long long>(a.T) * a.KT * c) / a.P;
}

// TN = 64 or 128 penalties per tile: 128 x 128 tiles (warp tile 32 x 64, 12 fragment loads per 32 MMAs, as in
// the SYRK kernel) are the efficient shape, and with equal k-step ranges their small number (64 tiles at 256
// penalties) no longer starves SMs
template <int TN, bool TMA>
__global__ void __launch_bounds__(256, 1) path_step_sk_kernel(const PathSkArgs a, const __grid_constant__ PathMaps maps) {
    constexpr int PM = 128;
    constexpr int NJ = PathTile<PM, TN>::NJ;
    extern __shared__ __align__(16) double smem[];
    __shared__ double part[8][PathTile<PM, TN>::LW][4];
    __shared__ unsigned s_ticket;
    const PathArgs& p = a.p;
    const int tid = threadIdx.x;
    const int c = blockIdx.x;
    PathRing rg;
    if constexpr (TMA) path_ring_init<PM, TN>(rg, reinterpret_cast<unsigned char*>(smem));
    const int nlb = p.Lpad / TN;                 // l-blocks; tile t = (iblk = t / nlb, lblk = t % nlb)
    const long long r0 = sk_begin(a, c), r1 = sk_begin(a, c + 1);
    double acc[4][NJ][2];
    for (long long r = r0; r < r1;) {
        const int t = static_cast<int>(r / a.KT);
        const int k_lo = static_cast<int>(r - static_cast<long long>(t) * a.KT);
        const int k_hi = static_cast<int>(min(static_cast<long long>(a.KT), r1 - static_cast<long long>(t) * a.KT));
        const int iblk = t / nlb, i0 = iblk * PM, l0 = (t % nlb) * TN;
        if constexpr (TMA) path_mainloop_tma<PM, TN>(maps, p.ymap, rg, i0, l0, k_lo, k_hi - k_lo, acc);
        else path_mainloop<PM, TN>(p, smem, i0, l0, k_lo, k_hi - k_lo, acc);
        if (k_lo == 0 && k_hi == a.KT) {
            path_epilogue<PM, TN>(p, i0, l0, iblk, acc, part);
        } else {
            // publish the partial (slot = 0 for a range that starts inside a tile, 1 for the one that ends inside)
            const int slot = (k_lo != 0) ? 0 : 1;
            double2* w = reinterpret_cast<double2*>(a.W + (static_cast<size_t>(c) * 2 + slot) * (PM * TN));
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) w[(i * NJ + j) * 256 + tid] = make_double2(acc[i][j][0], acc[i][j][1]);
            __threadfence();
            __syncthreads();
            // contributors of tile t: the CTAs whose ranges intersect [t KT, (t+1) KT)
            const long long tb = static_cast<long long>(t) * a.KT, te = tb + a.KT;
            int c_first = static_cast<int>((tb * a.P) / (static_cast<long long>(a.T) * a.KT));
            while (sk_begin(a, c_first + 1) <= tb) ++c_first;
            while (c_first > 0 && sk_begin(a, c_first) > tb) --c_first;
            int c_last = c_first;
            while (c_last + 1 < a.P && sk_begin(a, c_last + 1) < te) ++c_last;
            const unsigned ncontrib = static_cast<unsigned>(c_last - c_first + 1);
            if (tid == 0) s_ticket = atomicAdd(&a.ticket[t], 1u);
            __syncthreads();
            if (s_ticket == ncontrib - 1) {   // last to arrive: every partial of this tile is in memory
                __threadfence();
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
                for (int cc = c_first; cc <= c_last; ++cc) {
                    // slot 0: the contributor's range STARTS inside this tile; slot 1: it starts at or before the
                    // tile's first k-step (and ends inside it) -- the rule the contributors used above
                    const int slot_cc = (sk_begin(a, cc) > tb) ? 0 : 1;
                    const double2* wv = reinterpret_cast<const double2*>(a.W + (static_cast<size_t>(cc) * 2 + slot_cc) * (PM * TN));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const double2 v = __ldcg(&wv[(i * NJ + j) * 256 + tid]);
                            acc[i][j][0] += v.x;
                            acc[i][j][1] += v.y;
                        }
                }
                if (tid == 0) a.ticket[t] = 0u;   // ready for the next launch
                path_epilogue<PM, TN>(p, i0, l0, iblk, acc, part);
            }
            __syncthreads();
        }
        r = static_cast<long long>(t) * a.KT + k_hi;
    }
}

__global__ void path_obj_finish_kernel(const double* __restrict__ part, int nblk, int Lpad, const double* alpha1,
                                       double alpha2, double half_bb, double* __restrict__ obj) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= Lpad) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = 0; b < nblk; ++b)
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] += part[(static_cast<size_t>(b) * Lpad + l) * 4 + q];
    double v = 0.5 * s[0] - s[1] + half_bb;
    if (alpha2 > 0.0) v += 0.5 * alpha2 * s[3];
    if (alpha1[l] > 0.0) v += alpha1[l] * s[2];
    obj[l] = v;
}

// max over the real columns of ||x+ - x||_2, from the per-i-block partials (fixed order)
__global__ void path_step_finish_kernel(const double* __restrict__ part, int nblk, int Lpad, int n_lambda,
                                        double* __restrict__ out) {
    __shared__ double red[256];
    double m = 0.0;
    for (int l = threadIdx.x; l < n_lambda; l += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < nblk; ++b) s += part[static_cast<size_t>(b) * Lpad + l];
        m = fmax(m, sqrt(s));
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ------------------------------------------------------------------------------------------
// power iteration on the Gram matrix: w = G v; L = ||w||; v = w / L  (estimate_lipschitz,
// iterative_solvers.py:45-60, with A^T(A v) replaced by (A^T A) v).  Two launches per step,
// all of them queued up front; after the stop test fires the remaining ones return at once.
// ------------------------------------------------------------------------------------------
struct GramPowerState {
    double L, L_prev, tol;
    int pit, pit_max, done, pad;
};

// `st` (one GPU: finish kernel below) or `ctrl` (row-sharded ranks: the epilogue kernel's EOP_POWER
// step follows, which adds the ranks' products through the exchange windows) says when to stop.
__global__ void __launch_bounds__(256) gram_matvec_kernel(const double* __restrict__ G, const double* __restrict__ v,
                                                          double* __restrict__ w, int d, const GramPowerState* st,
                                                          const FosCtrl* ctrl) {
    if (st ? st->done : (ctrl->g_mode == GM_SKIP)) return;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= d) return;
    const double2* g2 = reinterpret_cast<const double2*>(G + static_cast<size_t>(row) * d);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const int n2 = d / 2;  // d is a multiple of 128: n2 is a multiple of 64
    for (int c = lane; c < n2; c += 64) {
        const double2 a0 = g2[c], a1 = g2[c + 32];
        const double2 x0 = v2[c], x1 = v2[c + 32];
        s[0] = fma(a0.x, x0.x, s[0]);
        s[1] = fma(a0.y, x0.y, s[1]);
        s[2] = fma(a1.x, x1.x, s[2]);
        s[3] = fma(a1.y, x1.y, s[3]);
    }
    const double t = fos_warp_sum((s[0] + s[1]) + (s[2] + s[3]));
    if (lane == 0) w[row] = t;
}

__global__ void __launch_bounds__(1024) gram_power_finish_kernel(const double* __restrict__ w, double* __restrict__ v,
                                                                 int d, GramPowerState* st) {
    if (st->done) return;
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double s = 0.0;
    for (int c = tid; c < d; c += 1024) s = fma(w[c], w[c], s);
    s = fos_warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        const double t = fos_warp_sum(red[lane]);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    const double L = sqrt(red[0]);
    for (int c = tid; c < d; c += 1024) v[c] = __ddiv_rn(w[c], L);
    if (tid == 0) {
        const int pit = st->pit + 1;
        st->done = (fabs(L - st->L_prev) < st->tol) || (pit >= st->pit_max);
        st->L = L;
        st->L_prev = L;
        st->pit = pit;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Strong-rule screening (gram.py: fista_path_screened): the Gram system restricted to a feature
// subset, G_S[i][j] = G[idx[i]][idx[j]], zero padded to the tile width (a padded feature has a zero
// row, a zero c entry and therefore stays at 0).
__global__ void gram_gather_kernel(const double* __restrict__ G, const double* __restrict__ c, int d,
                                   const int* __restrict__ idx, int n_idx, int dsub, double* __restrict__ Gs,
                                   double* __restrict__ cs) {
    const int i = blockIdx.y;
    const int si = (i < n_idx) ? idx[i] : -1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < dsub; j += gridDim.x * blockDim.x) {
        const int sj = (j < n_idx) ? idx[j] : -1;
        Gs[static_cast<size_t>(i) * dsub + j] = (si >= 0 && sj >= 0) ? G[static_cast<size_t>(si) * d + sj] : 0.0;
        if (i == 0) cs[j] = (sj >= 0) ? c[sj] : 0.0;
    }
}

// out[l][i] = (G x_l)[i] - c[i]: gradient of the smooth part (without the alpha2 term) for n_cols
// columns; one warp per (row, column).  Used for the strong rule and its KKT re-check.
__global__ void __launch_bounds__(256) gram_apply_kernel(const double* __restrict__ G, const double* __restrict__ c,
                                                         const double* __restrict__ X, double* __restrict__ out, int d) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= d) return;
    const double2* g2 = reinterpret_cast<const double2*>(G + static_cast<size_t>(row) * d);
    const double2* v2 = reinterpret_cast<const double2*>(X + static_cast<size_t>(blockIdx.y) * d);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const int n2 = d / 2;
    for (int q = lane; q < n2; q += 64) {
        const double2 a0 = g2[q], a1 = g2[q + 32];
        const double2 x0 = v2[q], x1 = v2[q + 32];
        s[0] = fma(a0.x, x0.x, s[0]);
        s[1] = fma(a0.y, x0.y, s[1]);
        s[2] = fma(a1.x, x1.x, s[2]);
        s[3] = fma(a1.y, x1.y, s[3]);
    }
    const double t = fos_warp_sum((s[0] + s[1]) + (s[2] + s[3]));
    if (lane == 0) out[static_cast<size_t>(blockIdx.y) * d + row] = t - c[row];
}

struct fos_gram {
    int device = 0;
    cudaStream_t stream = nullptr;
    int d = 0;
    double* G = nullptr;    // [d][d]
    double* c = nullptr;    // [d]
    double bb = 0.0;        // b^T b
    float build_ms = 0.f;
    int nsplit = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

static void gram_free(fos_gram* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    // recycled blocks are handed to the next request without the implicit device synchronisation of cudaFree:
    // nothing of this object may still be running
    if (g->stream) cudaStreamSynchronize(g->stream);
    fos_pool_free(g->G);
    fos_pool_free(g->c);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->stream) cudaStreamDestroy(g->stream);
    cudaGetLastError();
    delete g;
}


// ------------------------------------------------------------------------------------------
// Gram matrix accumulated under the host->device upload (fos_design_create), and the power
// iteration on it.  While the PCIe copy of a dense float64 design is in flight the SMs are
// idle; the rows that have already arrived are pushed through the SYRK kernel chunk by chunk
// (d^2 FMA per row against 8 d bytes over PCIe: hidden for d <= 4096), so that
// estimate_lipschitz (iterative_solvers.py:45-60: <= 100 x two passes over A in the reference)
// costs <= 100 products with the d x d matrix once the upload ends.
// ------------------------------------------------------------------------------------------
static size_t syrk_smem_bytes() { return static_cast<size_t>(GSTAGES) * 2 * GKB * GLD * sizeof(double); }

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup
typedef CUresult (*TsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TsEncodeFn ts_encode_fn() {
    static TsEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<TsEncodeFn>(p);
    }();
    return fn;
}
// A[rows][lda] as a 2-D tensor (columns innermost), boxes of 16 columns x GKB rows, 128-byte swizzle, zero fill
static bool syrk_tensor_map(const SyrkArgs& a, CUtensorMap* m) {
    TsEncodeFn fn = ts_encode_fn();
    if (!fn || (a.lda & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) || a.n <= 0 || a.n > 0x7fffffffLL) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(a.lda), static_cast<cuuint64_t>(a.n)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.lda) * sizeof(double)};
    const cuuint32_t box[2] = {16, GKB};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(a.A), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// tile products of one call: TMA-staged kernel, or the cp.async one (FOS_GRAM_TMA=0, odd pitch, no encoder)
static std::atomic<long long> g_syrk_tma{0}, g_syrk_cp{0}, g_path_tma{0}, g_path_cp{0};
static cudaError_t launch_syrk(const SyrkArgs& a, int nsplit, cudaStream_t s) {
    const char* e = getenv("FOS_GRAM_TMA");
    CUtensorMap m;
    if (!(e && e[0] == '0') && syrk_tensor_map(a, &m)) {
        cudaError_t st = cudaFuncSetAttribute(gram_syrk_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES);
        if (st != cudaSuccess) return st;
        gram_syrk_tma_kernel<<<dim3(a.ntiles * nsplit), dim3(256), TS_SMEM_BYTES, s>>>(a, m);
        g_syrk_tma += 1;
        return cudaGetLastError();
    }
    g_syrk_cp += 1;
    cudaError_t st = cudaFuncSetAttribute(gram_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(syrk_smem_bytes()));
    if (st != cudaSuccess) return st;
    gram_syrk_kernel<<<dim3(a.ntiles * nsplit), dim3(256), syrk_smem_bytes(), s>>>(a);
    return cudaGetLastError();
}

static int syrk_pick_split(const fos_design* h, int ntiles, long long rows_avail, long long d) {
    int best = 1;
    double best_eff = 0.0;
    long long max_split = std::min<long long>(h->sm_count, rows_avail / (4 * GKB));
    max_split = std::min<long long>(max_split, (1536LL << 20) / (d * d * 8));  // workspace <= 1.5 GB
    max_split = std::max<long long>(1, ntiles >= 64 ? std::min<long long>(max_split, 16) : max_split);
    for (int s = 1; s <= max_split; ++s) {
        const long long units = static_cast<long long>(ntiles) * s;
        const long long waves = (units + h->sm_count - 1) / h->sm_count;
        const double eff = static_cast<double>(units) / (waves * h->sm_count);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = s;
        }
    }
    return best;
}

// Split workspace of the upload-time SYRK (nsplit x d^2 doubles, 0.94 GB at d = 4096): one buffer per
// device is kept between uploads instead of a cudaMalloc / cudaFree pair per design (the free alone
// costs 5-35 ms when several processes share the box).  fos_trim() releases it.
#include <map>
#include <mutex>
static std::mutex g_ws_mu;
static std::map<int, std::pair<void*, size_t>> g_ws_cache;  // device -> (buffer, bytes), not in use

static void* ws_take(int device, size_t bytes) {
    {
        std::lock_guard<std::mutex> lock(g_ws_mu);
        auto it = g_ws_cache.find(device);
        if (it != g_ws_cache.end()) {
            void* p = it->second.first;
            const size_t have = it->second.second;
            g_ws_cache.erase(it);
            if (have >= bytes) return p;
            cudaFree(p);
        }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

static void ws_give(int device, void* p, size_t bytes) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_ws_mu);
        if (g_ws_cache.find(device) == g_ws_cache.end()) {
            g_ws_cache[device] = {p, bytes};
            return;
        }
    }
    cudaFree(p);
}

extern "C" int fos_trim(void) {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    for (auto& kv : g_ws_cache) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.first);
    }
    g_ws_cache.clear();
    cudaGetLastError();
    fos_block_cache_trim();
    return FOS_OK;
}

bool fos_upload_gram_eligible(const fos_design* h) {
    if (h->dtype != FOS_F64 || h->d % GT != 0 || h->d > 4096 || h->lda != h->d) return false;
    const char* e = getenv("FOS_UPLOAD_GRAM");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return h->n >= 16LL * GKB;
    // automatic: tall designs whose upload is long enough to hide the contraction
    const double bytes = static_cast<double>(h->n) * h->d * 8.0;
    return h->n >= 16LL * h->d && bytes >= 1.0e9;
}

long long fos_upload_gram_chunk_rows(const fos_design* h) {
    long long rows = (512LL << 20) / (static_cast<long long>(h->d) * 8);
    rows = std::max<long long>(rows, 64LL * GKB);
    if (const char* e = getenv("FOS_UPLOAD_GRAM_CHUNK_ROWS")) {  // tests: several chunks on a small design
        const long long v = atoll(e);
        if (v >= 1) rows = v;
    }
    return std::min<long long>(rows, h->n);
}

// allocate the split workspace and the result; a failed allocation only disables the feature
int fos_upload_gram_begin(fos_design* h, cudaStream_t s) {
    const long long d = h->d;
    const int nb = h->d / GT, ntiles = nb * (nb + 1) / 2;
    h->up_nsplit = syrk_pick_split(h, ntiles, fos_upload_gram_chunk_rows(h), d);
    const size_t wbytes = static_cast<size_t>(h->up_nsplit) * d * d * sizeof(double);
    const auto b0 = std::chrono::steady_clock::now();
    h->up_W = static_cast<double*>(ws_take(h->device, wbytes));
    h->up_W_bytes = wbytes;
    if (h->up_W == nullptr || fos_pool_malloc(reinterpret_cast<void**>(&h->G_up), d * d * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        fos_upload_gram_drop(h);
        return FOS_OK;
    }
    const auto b1 = std::chrono::steady_clock::now();
    FOS_CUDA(cudaFuncSetAttribute(gram_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(syrk_smem_bytes())));
    FOS_CUDA(cudaFuncSetAttribute(gram_syrk_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
    if (getenv("FOS_UPLOAD_DEBUG"))
        fprintf(stderr, "[fos] gram_begin: cudaMalloc(W %.0f MB + G) %.1f ms, func attribute %.1f ms\n", wbytes / 1e6,
                std::chrono::duration<double, std::milli>(b1 - b0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b1).count());
    FOS_CUDA(cudaMemsetAsync(h->up_W, 0, wbytes, s));
    return FOS_OK;
}

int fos_upload_gram_chunk(fos_design* h, long long row0, long long rows, cudaStream_t s) {
    if (!h->up_W || rows <= 0) return FOS_OK;
    SyrkArgs a{};
    a.A = static_cast<const double*>(h->A) + static_cast<size_t>(row0) * h->lda;
    a.n = rows;
    a.lda = h->lda;
    a.nb = h->d / GT;
    a.ntiles = a.nb * (a.nb + 1) / 2;
    a.dpad = h->d;
    a.W = h->up_W;
    a.accumulate = 1;
    long long rps = (rows + h->up_nsplit - 1) / h->up_nsplit;
    a.rows_per_split = (rps + GKB - 1) / GKB * GKB;
    FOS_CUDA(launch_syrk(a, h->up_nsplit, s));
    h->launches += 1;
    return FOS_OK;
}

int fos_upload_gram_finish(fos_design* h, cudaStream_t s) {
    if (!h->up_W) return FOS_OK;
    const long long d = h->d;
    gram_reduce_kernel<<<dim3(static_cast<unsigned>((d + 255) / 256), static_cast<unsigned>(d)), dim3(256), 0, s>>>(
        h->up_W, h->G_up, d, h->up_nsplit);
    FOS_CUDA(cudaGetLastError());
    const auto f0 = std::chrono::steady_clock::now();
    FOS_CUDA(cudaStreamSynchronize(s));
    const auto f1 = std::chrono::steady_clock::now();
    ws_give(h->device, h->up_W, h->up_W_bytes);
    if (getenv("FOS_UPLOAD_DEBUG"))
        fprintf(stderr, "[fos] gram_finish: sync %.1f ms, workspace release %.1f ms\n",
                std::chrono::duration<double, std::milli>(f1 - f0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - f1).count());
    h->up_W = nullptr;
    h->G_state = 1;
    h->launches += 1;
    return FOS_OK;
}

void fos_upload_gram_drop(fos_design* h) {
    if (h->G_up || h->up_W) cudaDeviceSynchronize();  // recycled blocks: no kernel may still read them (rare path)
    if (h->up_W) ws_give(h->device, h->up_W, h->up_W_bytes);
    fos_pool_free(h->G_up);
    h->up_W = nullptr;
    h->G_up = nullptr;
    h->G_state = 0;
}

// Row-sharded ranks: every rank holds the Gram matrix of ITS rows, so A^T A v = sum_r G_r v.  Each
// step is a local product followed by the epilogue kernel's power-iteration step (EOP_POWER) with
// the product standing in as the single "partial" row: that step sums the ranks' vectors in rank
// order through the exchange windows (the same fused all-reduce as every streaming pass: 32 KB,
// no library collective, bit-identical L on all ranks), normalises and applies the stop test.
static int gram_power_iter_sharded(fos_design* h, const double* v0, int n_iter, double tol, double* L_out,
                                   int* iters_out, float* gpu_ms_out) {
    const int d = h->d;
    void* base = nullptr;
    FOS_TRY(fos_arena_reserve(h, static_cast<size_t>(h->ldv) * sizeof(double), &base));
    double* w = static_cast<double*>(base);
    FosCtrl* c = h->ctrl_host;
    memset(c, 0, sizeof(FosCtrl));
    c->g_mode = GM_GRAD | GM_NOB;
    c->phase = PH_DONE;
    c->ptol = tol;
    c->pit_max = n_iter;
    FOS_CUDA(cudaMemcpyAsync(h->ctrl, c, sizeof(FosCtrl), cudaMemcpyHostToDevice, h->stream));
    memcpy(h->vec_host, v0, static_cast<size_t>(d) * sizeof(double));
    for (int q = d; q < h->ldv; ++q) h->vec_host[q] = 0.0;
    FOS_CUDA(cudaMemcpyAsync(h->y, h->vec_host, static_cast<size_t>(h->ldv) * sizeof(double), cudaMemcpyHostToDevice,
                             h->stream));
    FOS_CUDA(cudaMemsetAsync(w, 0, static_cast<size_t>(h->ldv) * sizeof(double), h->stream));
    FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
    // the epilogue reads its partial rows from the design: point it at the product for this loop
    double* saved_pg = h->partial_g;
    const int saved_np = h->n_parts;
    h->partial_g = w;
    h->n_parts = 1;
    FosHist none{};
    int status = FOS_OK;
    for (int k = 0; k < n_iter && status == FOS_OK; ++k) {
        gram_matvec_kernel<<<dim3((d + 7) / 8), dim3(256), 0, h->stream>>>(h->G_up, h->y, w, d, nullptr, h->ctrl);
        h->launches += 1;
        status = fos_launch_epilogue(h, EOP_POWER, 0, none, 0.0, 0.0, 0);
    }
    h->partial_g = saved_pg;
    h->n_parts = saved_np;
    FOS_TRY(status);
    FOS_CUDA(cudaGetLastError());
    FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
    FOS_CUDA(cudaMemcpyAsync(c, h->ctrl, sizeof(FosCtrl), cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    if (c->stop_reason < 0) {
        fos_set_error("multi-GPU exchange timed out: a peer rank never arrived");
        return FOS_ERR_COMM;
    }
    *L_out = c->L;
    if (iters_out) *iters_out = c->pit;
    if (gpu_ms_out) FOS_CUDA(cudaEventElapsedTime(gpu_ms_out, h->ev0, h->ev1));
    return FOS_OK;
}

int fos_gram_power_iter(fos_design* h, const double* v0, int n_iter, double tol, double* L_out, int* iters_out,
                        float* gpu_ms_out) {
    if (h->world > 1) return gram_power_iter_sharded(h, v0, n_iter, tol, L_out, iters_out, gpu_ms_out);
    const int d = h->d;
    void* base = nullptr;
    FOS_TRY(fos_arena_reserve(h, 256 + static_cast<size_t>(d) * sizeof(double), &base));
    GramPowerState* st = static_cast<GramPowerState*>(base);
    double* w = reinterpret_cast<double*>(static_cast<char*>(base) + 256);
    GramPowerState init{};
    init.tol = tol;
    init.pit_max = n_iter;
    int status = FOS_OK;
    auto body = [&]() -> int {
        FOS_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
        // y <- v0 through the pinned staging buffer (the padding of y beyond d stays zero)
        memcpy(h->vec_host, v0, static_cast<size_t>(d) * sizeof(double));
        FOS_CUDA(cudaMemcpyAsync(h->y, h->vec_host, static_cast<size_t>(d) * sizeof(double), cudaMemcpyHostToDevice,
                                 h->stream));
        FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
        for (int k = 0; k < n_iter; ++k) {
            gram_matvec_kernel<<<dim3((d + 7) / 8), dim3(256), 0, h->stream>>>(h->G_up, h->y, w, d, st, nullptr);
            gram_power_finish_kernel<<<dim3(1), dim3(1024), 0, h->stream>>>(w, h->y, d, st);
        }
        FOS_CUDA(cudaGetLastError());
        h->launches += 2LL * n_iter;
        FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
        FOS_CUDA(cudaMemcpyAsync(&init, st, sizeof(init), cudaMemcpyDeviceToHost, h->stream));
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        *L_out = init.L;
        if (iters_out) *iters_out = init.pit;
        if (gpu_ms_out) FOS_CUDA(cudaEventElapsedTime(gpu_ms_out, h->ev0, h->ev1));
        return FOS_OK;
    };
    status = body();
    return status;
}

extern "C" int fos_gram_create(fos_design* h, fos_gram** out) {
    FOS_REQUIRE(h && out, "null pointer argument");
    FOS_REQUIRE(h->dtype == FOS_F64, "Gram mode needs float64 storage");
    FOS_REQUIRE(h->d % GT == 0, "Gram mode needs d to be a multiple of %d (got %d)", GT, h->d);
    FOS_CUDA(cudaSetDevice(h->device));
    fos_gram* g = new fos_gram();
    g->device = h->device;
    g->d = h->d;
    auto body = [&]() -> int {
        FOS_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
        FOS_CUDA(cudaEventCreate(&g->ev0));
        FOS_CUDA(cudaEventCreate(&g->ev1));
        const long long d = h->d;
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&g->G), static_cast<size_t>(d) * d * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&g->c), static_cast<size_t>(d) * sizeof(double)));
        // c = A^T b and b^T b from one pass of the streaming kernel with x = 0: g = -A^T b
        std::vector<double> zero(h->d, 0.0), gneg(h->d);
        double half_bb = 0.0;
        FOS_TRY(fos_grad(h, zero.data(), 0.0, gneg.data(), &half_bb));
        for (double& v : gneg) v = -v;
        g->bb = 2.0 * half_bb;
        FOS_CUDA(cudaMemcpy(g->c, gneg.data(), static_cast<size_t>(d) * sizeof(double), cudaMemcpyHostToDevice));

        if (h->G_up && h->G_state >= 1) {
            // already accumulated under the upload of this design: copy instead of rebuilding
            cudaEventRecord(g->ev0, g->stream);
            FOS_CUDA(cudaMemcpyAsync(g->G, h->G_up, static_cast<size_t>(d) * d * sizeof(double), cudaMemcpyDeviceToDevice,
                                     g->stream));
            cudaEventRecord(g->ev1, g->stream);
            FOS_CUDA(cudaStreamSynchronize(g->stream));
            FOS_CUDA(cudaEventElapsedTime(&g->build_ms, g->ev0, g->ev1));
            g->nsplit = h->up_nsplit;
            return FOS_OK;
        }
        SyrkArgs a{};
        a.A = static_cast<const double*>(h->A);
        a.n = h->n;
        a.lda = h->lda;
        a.nb = h->d / GT;
        a.ntiles = a.nb * (a.nb + 1) / 2;
        a.dpad = d;
        // split the rows so that tiles x splits fills whole waves of SMs
        int best = 1;
        double best_eff = 0.0;
        const long long max_split = std::max<long long>(1, std::min<long long>(16, h->n / (4 * GKB)));
        for (int s = 1; s <= max_split; ++s) {
            const long long units = static_cast<long long>(a.ntiles) * s;
            const long long waves = (units + h->sm_count - 1) / h->sm_count;
            const double eff = static_cast<double>(units) / (waves * h->sm_count);
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                best = s;
            }
        }
        g->nsplit = best;
        long long rps = (h->n + best - 1) / best;
        rps = (rps + GKB - 1) / GKB * GKB;
        a.rows_per_split = rps;
        double* W = nullptr;
        FOS_CUDA(cudaMalloc(&W, static_cast<size_t>(best) * d * d * sizeof(double)));
        a.W = W;
        cudaEventRecord(g->ev0, g->stream);
        cudaError_t e = launch_syrk(a, best, g->stream);
        if (e == cudaSuccess) {
            gram_reduce_kernel<<<dim3(static_cast<unsigned>((d + 255) / 256), static_cast<unsigned>(d)), dim3(256), 0,
                                 g->stream>>>(W, g->G, d, best);
            cudaEventRecord(g->ev1, g->stream);
            e = cudaStreamSynchronize(g->stream);
        }
        cudaFree(W);
        if (e != cudaSuccess) {
            fos_set_error("Gram build failed: %s", cudaGetErrorString(e));
            return FOS_ERR_CUDA;
        }
        FOS_CUDA(cudaEventElapsedTime(&g->build_ms, g->ev0, g->ev1));
        return FOS_OK;
    };
    int st = body();
    if (st != FOS_OK) {
        gram_free(g);
        return st;
    }
    *out = g;
    return FOS_OK;
}

extern "C" int fos_gram_destroy(fos_gram* g) {
    gram_free(g);
    return FOS_OK;
}

extern "C" int fos_gram_info(const fos_gram* g, int* d, double* btb, float* build_ms, int* nsplit) {
    FOS_REQUIRE(g, "null gram handle");
    if (d) *d = g->d;
    if (btb) *btb = g->bb;
    if (build_ms) *build_ms = g->build_ms;
    if (nsplit) *nsplit = g->nsplit;
    return FOS_OK;
}

extern "C" int fos_gram_pointers(fos_gram* g, double** G_dev, double** c_dev) {
    FOS_REQUIRE(g, "null gram handle");
    if (G_dev) *G_dev = g->G;
    if (c_dev) *c_dev = g->c;
    return FOS_OK;
}

extern "C" int fos_gram_download(fos_gram* g, double* G_out, double* c_out) {
    FOS_REQUIRE(g, "null gram handle");
    FOS_CUDA(cudaSetDevice(g->device));
    const size_t d = g->d;
    if (G_out) FOS_CUDA(cudaMemcpy(G_out, g->G, d * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (c_out) FOS_CUDA(cudaMemcpy(c_out, g->c, d * sizeof(double), cudaMemcpyDeviceToHost));
    return FOS_OK;
}

// replace G and c (after an all-reduce over row-sharded ranks done by the caller)
extern "C" int fos_gram_subset(fos_gram* g, const int* idx, int n_idx, fos_gram** out) {
    FOS_REQUIRE(g && idx && out, "null pointer argument");
    FOS_REQUIRE(n_idx >= 1 && n_idx <= g->d, "subset size %d out of range (1..%d)", n_idx, g->d);
    for (int i = 0; i < n_idx; ++i)
        FOS_REQUIRE(idx[i] >= 0 && idx[i] < g->d && (i == 0 || idx[i] > idx[i - 1]),
                    "subset indices must be strictly increasing and lie in [0, %d)", g->d);
    FOS_CUDA(cudaSetDevice(g->device));
    fos_gram* s = new fos_gram();
    s->device = g->device;
    s->d = (n_idx + GT - 1) / GT * GT;
    s->bb = g->bb;
    s->nsplit = 0;
    int* idx_dev = nullptr;
    auto body = [&]() -> int {
        FOS_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        FOS_CUDA(cudaEventCreate(&s->ev0));
        FOS_CUDA(cudaEventCreate(&s->ev1));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&s->G), static_cast<size_t>(s->d) * s->d * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&s->c), static_cast<size_t>(s->d) * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&idx_dev), static_cast<size_t>(n_idx) * sizeof(int)));
        FOS_CUDA(cudaMemcpyAsync(idx_dev, idx, static_cast<size_t>(n_idx) * sizeof(int), cudaMemcpyHostToDevice, s->stream));
        FOS_CUDA(cudaStreamSynchronize(g->stream));  // the parent matrix is complete
        FOS_CUDA(cudaEventRecord(s->ev0, s->stream));
        gram_gather_kernel<<<dim3(static_cast<unsigned>((s->d + 255) / 256), static_cast<unsigned>(s->d)), dim3(256), 0,
                             s->stream>>>(g->G, g->c, g->d, idx_dev, n_idx, s->d, s->G, s->c);
        FOS_CUDA(cudaGetLastError());
        FOS_CUDA(cudaEventRecord(s->ev1, s->stream));
        FOS_CUDA(cudaStreamSynchronize(s->stream));
        FOS_CUDA(cudaEventElapsedTime(&s->build_ms, s->ev0, s->ev1));
        return FOS_OK;
    };
    const int st = body();
    fos_pool_free(idx_dev);
    if (st != FOS_OK) {
        gram_free(s);
        return st;
    }
    *out = s;
    return FOS_OK;
}

extern "C" int fos_gram_apply(fos_gram* g, const double* X, int n_cols, double* out) {
    FOS_REQUIRE(g && X && out, "null pointer argument");
    FOS_REQUIRE(n_cols >= 1 && n_cols <= 65535, "n_cols out of range");
    FOS_CUDA(cudaSetDevice(g->device));
    const size_t bytes = static_cast<size_t>(n_cols) * g->d * sizeof(double);
    double* buf = nullptr;
    FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&buf), 2 * bytes));
    auto body = [&]() -> int {
        FOS_CUDA(cudaMemcpyAsync(buf, X, bytes, cudaMemcpyHostToDevice, g->stream));
        gram_apply_kernel<<<dim3(static_cast<unsigned>((g->d + 7) / 8), static_cast<unsigned>(n_cols)), dim3(256), 0,
                            g->stream>>>(g->G, g->c, buf, buf + static_cast<size_t>(n_cols) * g->d, g->d);
        FOS_CUDA(cudaGetLastError());
        FOS_CUDA(cudaMemcpyAsync(out, buf + static_cast<size_t>(n_cols) * g->d, bytes, cudaMemcpyDeviceToHost, g->stream));
        FOS_CUDA(cudaStreamSynchronize(g->stream));
        return FOS_OK;
    };
    const int st = body();
    if (st != FOS_OK) cudaStreamSynchronize(g->stream);
    fos_pool_free(buf);
    return st;
}

extern "C" int fos_gram_set_btb(fos_gram* g, double btb) {
    FOS_REQUIRE(g, "null gram handle");
    g->bb = btb;
    return FOS_OK;
}

extern "C" int fos_debug_gram_staging(long long* tma_launches, long long* cp_async_launches) {
    if (tma_launches) *tma_launches = g_syrk_tma.load();
    if (cp_async_launches) *cp_async_launches = g_syrk_cp.load();
    return FOS_OK;
}
extern "C" int fos_debug_path_staging(long long* tma_solves, long long* cp_async_solves) {
    if (tma_solves) *tma_solves = g_path_tma.load();
    if (cp_async_solves) *cp_async_solves = g_path_cp.load();
    return FOS_OK;
}

// G and the three l x d matrices of a path solve as 2-D tensors (k innermost), boxes of 16 k-values x pm / tn rows
static bool path_tensor_maps(PathMaps* m, const double* G, const double* Y0, const double* Y1, const double* X, int d,
                             int Lpad, int pm, int tn) {
    TsEncodeFn fn = ts_encode_fn();
    if (!fn || PK != 16 || (d & 1)) return false;
    auto enc = [&](CUtensorMap* t, const double* base, int rows, int box_rows) {
        if (reinterpret_cast<uintptr_t>(base) & 15) return false;
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows)};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d) * sizeof(double)};
        const cuuint32_t box[2] = {16, static_cast<cuuint32_t>(box_rows)};
        const cuuint32_t estr[2] = {1, 1};
        return fn(t, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    return enc(&m->g, G, d, pm) && enc(&m->y[0], Y0, Lpad, tn) && enc(&m->y[1], Y1, Lpad, tn) && enc(&m->y[2], X, Lpad, tn);
}

// cudaFuncSetAttribute once per kernel AND device: the attribute lives in the device's context, and one process may
// hold designs on several GPUs
static cudaError_t set_smem_once(const void* fn, int bytes, std::atomic<unsigned long long>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// maps != nullptr: operands staged by the TMA unit; nullptr: the cp.async ring
template <int PM>
static cudaError_t launch_path(const PathArgs& p, const PathMaps* maps, cudaStream_t st) {
    if (maps) {
        constexpr int smem = PathTma<PM, PN>::SMEM_BYTES;
        static std::atomic<unsigned long long> configured{0};
        cudaError_t e = set_smem_once(reinterpret_cast<const void*>(&path_step_kernel<PM, true>), smem, configured);
        if (e != cudaSuccess) return e;
        path_step_kernel<PM, true><<<dim3(p.d / PM, p.Lpad / PN), dim3(256), smem, st>>>(p, *maps);
        return cudaGetLastError();
    }
    const size_t smem = static_cast<size_t>(PSTAGES) * (PM + PN) * PLD * sizeof(double);
    static std::atomic<unsigned long long> configured{0};
    cudaError_t e = set_smem_once(reinterpret_cast<const void*>(&path_step_kernel<PM, false>), static_cast<int>(smem), configured);
    if (e != cudaSuccess) return e;
    path_step_kernel<PM, false><<<dim3(p.d / PM, p.Lpad / PN), dim3(256), smem, st>>>(p, PathMaps{});
    return cudaGetLastError();
}

// stream-K launch (128-row tiles only); W and ticket are owned by the caller
template <int TN>
static cudaError_t launch_path_sk(const PathArgs& p, const PathMaps* maps, double* W, unsigned* ticket, int P,
                                  cudaStream_t st) {
    PathSkArgs a;
    a.p = p;
    a.W = W;
    a.ticket = ticket;
    a.T = (p.d / 128) * (p.Lpad / TN);
    a.KT = p.d / PK;
    a.P = P;
    if (maps) {
        constexpr int smem = PathTma<128, TN>::SMEM_BYTES;
        static std::atomic<unsigned long long> attr_done{0};
        cudaError_t e = set_smem_once(reinterpret_cast<const void*>(&path_step_sk_kernel<TN, true>), smem, attr_done);
        if (e != cudaSuccess) return e;
        path_step_sk_kernel<TN, true><<<dim3(P), dim3(256), smem, st>>>(a, *maps);
        return cudaGetLastError();
    }
    const size_t smem = static_cast<size_t>(PSTAGES) * (128 + TN) * PLD * sizeof(double);
    static std::atomic<unsigned long long> attr_done{0};
    cudaError_t e = set_smem_once(reinterpret_cast<const void*>(&path_step_sk_kernel<TN, false>), static_cast<int>(smem), attr_done);
    if (e != cudaSuccess) return e;
    path_step_sk_kernel<TN, false><<<dim3(P), dim3(256), smem, st>>>(a, PathMaps{});
    return cudaGetLastError();
}

static cudaError_t launch_path_pm(int pm, const PathArgs& p, const PathMaps* maps, cudaStream_t st) {
    if (pm == 128) return launch_path<128>(p, maps, st);
    if (pm == 64) return launch_path<64>(p, maps, st);
    return launch_path<32>(p, maps, st);
}

// Schedule and tile shape of the batched path iteration for d features and n_lambda penalties on a part with
// sm_count SMs (pure host logic: fos_debug_path_plan exposes it to the CPU tests)
struct PathPlan {
    int Lpad;            // padded penalty count (rows of Y / X)
    int pm, tn;          // tile: pm features x tn penalties (tile schedule: tn = PN)
    bool use_sk;         // stream-K schedule (128-row tiles) instead of one tile per CTA
    long long n_tiles;   // 128-row tiles of the stream-K schedule (ticket words)
};
static PathPlan path_plan(int d, int n_lambda, int sm_count) {
    int Lpad = (n_lambda + PN - 1) / PN * PN;
    // i-tile: the largest that still gives every SM a tile
    int pm = 128;
    while (pm > 32 && static_cast<long long>(d / pm) * (Lpad / PN) < 120) pm /= 2;
    // Stream-K over 128-row tiles whenever the tile count does not fill the SMs in whole waves (FOS_PATH_SK=0/1
    // forces the choice): every SM gets the same number of k-steps whatever the number of penalties.
    auto sk_fits = [&](long long tiles) { return d % 128 == 0 && tiles * (d / PK) >= 4LL * sm_count && tiles <= 4LL * sm_count; };
    bool use_sk = sk_fits(static_cast<long long>(d / 128) * (Lpad / PN));
    if (const char* e = getenv("FOS_PATH_SK")) use_sk = use_sk && e[0] != '0';
    // penalties per tile: 128 when that costs no extra padding; 32 when the last block of 64 would be at most half
    // full and the list is short (the 32 penalties a rank holds of a 256-penalty path on 8 GPUs: half the
    // contraction of a 64-wide tile); FOS_PATH_TN=32|64|128 forces a width that divides the padded count
    int tn = (use_sk && Lpad % 128 == 0) ? 128 : PN;
    const int Lpad32 = (n_lambda + 31) / 32 * 32;
    bool tn32 = use_sk && Lpad32 < Lpad && Lpad32 <= 160 && sk_fits(static_cast<long long>(d / 128) * (Lpad32 / 32));
    if (const char* e = getenv("FOS_PATH_TN")) {
        const int want = atoi(e);
        tn32 = tn32 && want == 32;
        if (want != 32) tn = (want == 128 && use_sk && Lpad % 128 == 0) ? 128 : PN;
    }
    if (tn32) {
        tn = 32;
        Lpad = Lpad32;
    }
    if (use_sk) pm = 128;
    PathPlan out;
    out.Lpad = Lpad;
    out.pm = pm;
    out.tn = tn;
    out.use_sk = use_sk;
    out.n_tiles = static_cast<long long>(d / 128) * (Lpad / tn);
    return out;
}

extern "C" int fos_debug_path_plan(int d, int n_lambda, int sm_count, int* padded_lambdas, int* tile_rows, int* tile_cols,
                                   int* stream_k, long long* n_tiles) {
    FOS_REQUIRE(d >= 128 && d % 128 == 0 && n_lambda >= 1 && sm_count >= 1, "bad argument");
    const PathPlan pl = path_plan(d, n_lambda, sm_count);
    if (padded_lambdas) *padded_lambdas = pl.Lpad;
    if (tile_rows) *tile_rows = pl.pm;
    if (tile_cols) *tile_cols = pl.tn;
    if (stream_k) *stream_k = pl.use_sk ? 1 : 0;
    if (n_tiles) *n_tiles = pl.n_tiles;
    return FOS_OK;
}

extern "C" int fos_gram_path_fista(fos_gram* g, const fos_path_params* pp, fos_path_result* pr) {
    FOS_REQUIRE(g && pp && pr && pp->alphas1, "null pointer argument");
    FOS_REQUIRE(pp->n_lambda >= 1 && pp->max_iter >= 0 && pp->step > 0.0, "bad argument");
    FOS_REQUIRE(pp->check_every >= 1 || pp->tol <= 0.0, "check_every must be >= 1 when tol > 0");
    FOS_CUDA(cudaSetDevice(g->device));
    const int d = g->d, n_lambda = pp->n_lambda;
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, g->device);
    const PathPlan plan = path_plan(d, n_lambda, sm_count);
    const int Lpad = plan.Lpad, pm = plan.pm, tn = plan.tn;
    const bool use_sk = plan.use_sk;
    const long long n_tiles = plan.n_tiles;
    const int nblk = d / pm;
    const size_t mat = static_cast<size_t>(Lpad) * d;
    double *Y0 = nullptr, *Y1 = nullptr, *X = nullptr, *a1 = nullptr, *part = nullptr, *obj = nullptr,
           *spart = nullptr, *smax = nullptr;
    double* smax_host = nullptr;
    double* skW = nullptr;
    unsigned* sk_ticket = nullptr;
    auto cleanup = [&]() {
        for (double* q : {Y0, Y1, X, a1, part, obj, spart, smax, skW})
            fos_pool_free(q);
        fos_pool_free(smax_host);
        fos_pool_free(sk_ticket);
    };
    auto body = [&]() -> int {
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&Y0), mat * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&Y1), mat * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&X), mat * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&a1), Lpad * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&part), static_cast<size_t>(nblk) * Lpad * 4 * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&spart), static_cast<size_t>(nblk) * Lpad * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&obj), Lpad * sizeof(double)));
        FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&smax), sizeof(double)));
        FOS_CUDA(fos_pool_malloc_host(reinterpret_cast<void**>(&smax_host), sizeof(double)));
        if (use_sk) {
            FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&skW), static_cast<size_t>(sm_count) * 2 * 128 * tn * sizeof(double)));
            FOS_CUDA(fos_pool_malloc(reinterpret_cast<void**>(&sk_ticket), static_cast<size_t>(n_tiles) * sizeof(unsigned)));
            FOS_CUDA(cudaMemsetAsync(sk_ticket, 0, static_cast<size_t>(n_tiles) * sizeof(unsigned), g->stream));
        }
        FOS_CUDA(cudaMemsetAsync(Y0, 0, mat * sizeof(double), g->stream));
        FOS_CUDA(cudaMemsetAsync(Y1, 0, mat * sizeof(double), g->stream));
        FOS_CUDA(cudaMemsetAsync(X, 0, mat * sizeof(double), g->stream));
        if (pp->X0) {  // warm start: x_0 = y_0 = X0 (rows beyond n_lambda stay zero)
            FOS_CUDA(cudaMemcpyAsync(X, pp->X0, static_cast<size_t>(n_lambda) * d * sizeof(double),
                                     cudaMemcpyHostToDevice, g->stream));
            FOS_CUDA(cudaMemcpyAsync(Y0, X, static_cast<size_t>(n_lambda) * d * sizeof(double),
                                     cudaMemcpyDeviceToDevice, g->stream));
        }
        std::vector<double> al(Lpad, 0.0);
        for (int l = 0; l < n_lambda; ++l) al[l] = pp->alphas1[l];
        FOS_CUDA(cudaMemcpyAsync(a1, al.data(), Lpad * sizeof(double), cudaMemcpyHostToDevice, g->stream));
        // operand staging: TMA tensor maps (FOS_PATH_TMA=0: the cp.async ring); same bits either way
        PathMaps maps_store;
        const PathMaps* maps = nullptr;
        {
            const char* e = getenv("FOS_PATH_TMA");
            if (!(e && e[0] == '0') && path_tensor_maps(&maps_store, g->G, Y0, Y1, X, d, Lpad, pm, use_sk ? tn : PN))
                maps = &maps_store;
        }
        (maps ? g_path_tma : g_path_cp) += 1;
        auto launch_step = [&](const PathArgs& q) -> cudaError_t {
            if (!use_sk) return launch_path_pm(pm, q, maps, g->stream);
            if (tn == 128) return launch_path_sk<128>(q, maps, skW, sk_ticket, sm_count, g->stream);
            if (tn == 32) return launch_path_sk<32>(q, maps, skW, sk_ticket, sm_count, g->stream);
            return launch_path_sk<PN>(q, maps, skW, sk_ticket, sm_count, g->stream);
        };
        PathArgs p{};
        p.G = g->G;
        p.c = g->c;
        p.X = X;
        p.alpha1 = a1;
        p.alpha2 = pp->alpha2;
        p.tau = pp->step;
        p.d = d;
        p.Lpad = Lpad;
        p.obj_part = part;
        double t_prev = 1.0;
        int64_t n_launch = 0;
        int iters = 0;
        double last_step = -1.0;
        FOS_CUDA(cudaEventRecord(g->ev0, g->stream));
        for (int k = 0; k < pp->max_iter; ++k) {
            // Nesterov momentum of fista (iterative_solvers.py:219-221), same for every column
            const double t_cur = 0.5 * (1.0 + sqrt(1.0 + 4.0 * (t_prev * t_prev)));
            p.beta = (t_prev - 1.0) / t_cur;
            t_prev = t_cur;
            p.mode = 0;
            p.Yin = (k & 1) ? Y1 : Y0;
            p.Yout = (k & 1) ? Y0 : Y1;
            p.ymap = k & 1;
            const bool check = pp->tol > 0.0 && ((k + 1) % pp->check_every == 0);
            p.step_part = check ? spart : nullptr;
            FOS_CUDA(launch_step(p));
            ++n_launch;
            ++iters;
            if (check) {
                // stop when every column's step ||x_{k+1}-x_k||_2 is below tol (iterative_solvers.py:238)
                path_step_finish_kernel<<<1, 256, 0, g->stream>>>(spart, nblk, Lpad, n_lambda, smax);
                FOS_CUDA(cudaMemcpyAsync(smax_host, smax, sizeof(double), cudaMemcpyDeviceToHost, g->stream));
                FOS_CUDA(cudaStreamSynchronize(g->stream));
                ++n_launch;
                last_step = *smax_host;
                if (last_step < pp->tol) break;
            }
        }
        FOS_CUDA(cudaEventRecord(g->ev1, g->stream));
        p.mode = 1;
        p.Yin = X;
        p.ymap = 2;
        p.Yout = nullptr;
        p.step_part = nullptr;
        FOS_CUDA(launch_step(p));
        path_obj_finish_kernel<<<dim3((Lpad + 127) / 128), dim3(128), 0, g->stream>>>(part, nblk, Lpad, a1, pp->alpha2,
                                                                                   0.5 * g->bb, obj);
        n_launch += 2;
        FOS_CUDA(cudaGetLastError());
        FOS_CUDA(cudaStreamSynchronize(g->stream));
        if (pr->X)
            FOS_CUDA(cudaMemcpy(pr->X, X, static_cast<size_t>(n_lambda) * d * sizeof(double), cudaMemcpyDeviceToHost));
        if (pr->obj) FOS_CUDA(cudaMemcpy(pr->obj, obj, n_lambda * sizeof(double), cudaMemcpyDeviceToHost));
        FOS_CUDA(cudaEventElapsedTime(&pr->loop_ms, g->ev0, g->ev1));
        pr->kernel_launches = n_launch;
        pr->n_iters = iters;
        pr->last_max_step = last_step;
        pr->tile_rows = pm;
        return FOS_OK;
    };
    const int st = body();
    if (st != FOS_OK) cudaStreamSynchronize(g->stream);
    cleanup();
    return st;
}
