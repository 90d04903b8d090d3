// grad_kernels.cu -- the fused single-pass gradient  g = A^T (A v1 - b)  (+ second residual
// norm ||A v2 - b||^2 in the same pass).  Replaces the two dgemv calls of the reference
// (iterative_solvers.py:173, :292; lbfgs.py:46-48) and the extra dgemv of the objective
// (iterative_solvers.py:225, objective_functions.py:13): A is streamed from HBM exactly once.
//
// Streaming kernel (d > 512): one persistent CTA per SM.  A producer thread moves contiguous
// row groups global -> shared with cp.async.bulk (TMA bulk engine, SASS UBLKCP) into an
// mbarrier ring with an L2 evict-first policy; NT consumer threads own CPT columns each, keep
// the row group in registers between the dot product (r_i = a_i . v - b_i) and the rank-1
// update (acc += r_i * a_i), and exchange the per-warp dot partials through a double-buffered
// shared array with ONE named barrier per stage.  Every sum has a fixed order (static row
// partition, ordered cross-warp and cross-CTA sums), so results are bit-reproducible.
//
// The pass mode (gradient / second dot / both) is read from the device control block and
// dispatched ONCE per CTA to a compile-time specialised loop, so the steady-state loop has
// no mode predicates, no column guards and no integer division.
//
// HBM bound: algorithmic bytes per pass = n*lda*sizeof(T) + 8n; 4 flop / 8 B in fp64.
#include "fos_common.cuh"
#include "pg_logic.cuh"

namespace {

// ---------------------------------------------------------------- PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
template <typename T>
struct Vec;
template <>
struct Vec<double> {
    static constexpr int N = 2;
    using type = double2;
    __device__ static __forceinline__ void load_shared(uint32_t addr, double (&o)[2]) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr));
    }
};
template <>
struct Vec<float> {
    static constexpr int N = 4;
    using type = float4;
    __device__ static __forceinline__ void load_shared(uint32_t addr, double (&o)[4]) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
        o[0] = x;
        o[1] = y;
        o[2] = z;
        o[3] = w;
    }
};


constexpr int MAX_STAGES = 8;

// ------------------------------------------------------------------------------------------
// streaming kernel
//   NT  consumer threads (plus one producer warp)
//   CPT columns per consumer thread (NT*CPT >= lda)
//   R   rows per pipeline stage
// ------------------------------------------------------------------------------------------
template <int NW, int R>
struct StreamSmem {
    uint64_t full_bar[MAX_STAGES];
    double red[2][R][NW][2];  // [parity][row][warp][dot1, dot2]
};

// Consumer loop, specialised on the pass mode.  GRAD: first dot + rank-1 update;
// DOT2: second dot.  Column guards are folded into per-thread shared-memory offsets
// (an out-of-range column reads offset 0 and multiplies it by a zero vector entry).
// WRAP (persistent solve kernel): the ring keeps running across passes -- once the last stages of this
// pass have been requested, freed slots are refilled with the FIRST stages of the next pass over the
// same rows (A never changes), so HBM keeps streaming while the pass's tail runs; slot_io / parity_io
// carry the ring position from pass to pass, and the vectors are read through L2 (other CTAs of the
// same launch wrote them).  Requires nst >= nstage.
// QREC (with GRAD, without DOT2): the second residual norm s2 = |A x_k - b|^2 comes from the recurrence
//   q_k = (r_y + beta q_{k-1}) / (1 + beta),  r_y = A y_k - b,  y_k = x_k + beta (x_k - x_{k-1})
// kept row by row in a.qres (16 bytes of traffic per row instead of 16 FMAs per 16 elements: the pass is 10 %
// faster on a power-capped part).  Warp 0 carries it: q_{k-1} of the stage's rows is requested at the top of
// the stage and consumed after the ordered sum, one FMA + one multiply per row.
template <typename T, int NT, int CPT, int R, bool GRAD, bool DOT2, bool WRAP = false, bool QREC = false>
__device__ __forceinline__ void stream_consume(const GradArgs& a, StreamSmem<NT / 32, R>& sm,
                                               unsigned char* ring, int stage_bytes, int nstage,
                                               long long lo, long long hi, bool use_b, uint64_t pol, int cta,
                                               int& slot_io, uint32_t& parity_io, double beta_y = 0.0, int qsel = 0) {
    constexpr int VEC = Vec<T>::N;
    constexpr int NV = CPT / VEC;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t row_bytes = static_cast<uint32_t>(a.lda) * sizeof(T);
    const int nst = static_cast<int>((hi - lo + R - 1) / R);

    double v1[NV][VEC], v2[NV][VEC], acc[NV][VEC];
    uint32_t off[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c0 = VEC * (tid + NT * j);
        off[j] = (c0 < a.lda) ? static_cast<uint32_t>(c0) * sizeof(T) : 0u;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int c = c0 + e;
            v1[j][e] = (GRAD && c < a.d) ? (WRAP ? __ldcg(a.v1 + c) : a.v1[c]) : 0.0;
            v2[j][e] = (DOT2 && c < a.d) ? (WRAP ? __ldcg(a.v2 + c) : a.v2[c]) : 0.0;
            acc[j][e] = 0.0;
        }
    }
    double s1 = 0.0, s2 = 0.0;
    const double q_inv = QREC ? 1.0 / (1.0 + beta_y) : 1.0;
    const bool q_here = QREC && (warp == NW - 1);
    // two copies of q, written and read in alternation from pass to pass: storing q_k(row i) and loading
    // q_{k-1}(row i+2) of the SAME 32-byte sector in one stage made the load wait for the store (measured at
    // 125k x 4096: CTA 0's streaming phase 575 -> 540 us per pass; the pass itself, set by the slowest CTA, did
    // not move: 0.644 -> 0.648 ms)
    double* const q_wr = QREC ? a.qres + static_cast<size_t>(qsel & 1) * static_cast<size_t>(a.n) : nullptr;
    const double* const q_rd = QREC ? a.qres + static_cast<size_t>((qsel & 1) ^ 1) * static_cast<size_t>(a.n) : nullptr;
    double q_nxt[R], q_p[R], r_p[R];
    int rows_p = 0;   // rows of the previous stage whose recurrence update is still pending
#pragma unroll
    for (int r = 0; r < R; ++r) {
        q_nxt[r] = (QREC && q_here && beta_y != 0.0 && lo + r < hi) ? __ldcg(q_rd + lo + r) : 0.0;
        q_p[r] = r_p[r] = 0.0;
    }

    // b values of the next two stages, prefetched by lanes < R of both half-warps of warp 0 (each
    // half subtracts b from the dot it ends up holding, see the paired butterfly below)
    const int blane = lane & 15;
    const bool b_lane = use_b && warp == 0 && blane < R;
    const double* bp = a.b + lo + blane;
    long long b_left = hi - lo - blane;  // rows remaining for this lane's prefetch stream
    double b_cur = 0.0, b_nxt = 0.0;
    if (b_lane && b_left > 0) b_cur = __ldg(bp);
    if (b_lane && b_left > R) b_nxt = __ldg(bp + R);

    const uint32_t ring_u32 = smem_u32(ring);
    int slot = slot_io;
    uint32_t parity = parity_io;
    for (int s = 0; s < nst; ++s) {
        double b_far = 0.0;
        if (b_lane && b_left > 2 * R) b_far = __ldg(bp + 2 * R);
        bp += R;
        b_left -= R;
        const int rows = static_cast<int>(min(static_cast<long long>(R), hi - lo - static_cast<long long>(s) * R));
        const uint32_t st = ring_u32 + static_cast<uint32_t>(slot) * static_cast<uint32_t>(stage_bytes);
        // (the LAST warp carries the recurrence: warp 0 already subtracts b and forms the residual sums, and the
        // other warps wait for the slowest one at the next barrier; q_{k-1} is requested one stage ahead)
        double q_old[R];
        if (QREC && q_here) {
            // the update of the PREVIOUS stage's rows runs here, under this stage's barrier wait and loads: at the
            // end of a stage it would sit on the path every warp waits for (40 cycles per stage, measured)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < rows_p) {
                    const double qn = fma(beta_y, q_p[r], r_p[r]) * q_inv;
                    s2 = fma(qn, qn, s2);
                    if (lane == 0) q_wr[lo + static_cast<long long>(s - 1) * R + r] = qn;
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                q_old[r] = q_nxt[r];
                const long long rn = lo + static_cast<long long>(s + 1) * R + r;
                q_nxt[r] = (beta_y != 0.0 && rn < hi) ? __ldcg(q_rd + rn) : 0.0;
            }
        }

        mbar_wait(&sm.full_bar[slot], parity);

        double av[R][NV][VEC];
        if (rows == R) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < NV; ++j) Vec<T>::load_shared(st + r * row_bytes + off[j], av[r][j]);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    if (r < rows) {
                        Vec<T>::load_shared(st + r * row_bytes + off[j], av[r][j]);
                    } else {
#pragma unroll
                        for (int e = 0; e < VEC; ++e) av[r][j][e] = 0.0;
                    }
                }
        }

        // per-thread partial dots: NCH independent FMA chains per dot (fixed assignment of
        // elements to chains and a fixed combine order: deterministic)
        constexpr int NCH = (CPT >= 16) ? 4 : 2;
        double d1[R], d2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double p1[NCH], p2[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) p1[c] = p2[c] = 0.0;
#pragma unroll
            for (int j = 0; j < NV; ++j)
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int c = (j * VEC + e) % NCH;
                    if (GRAD) p1[c] = fma(av[r][j][e], v1[j][e], p1[c]);
                    if (DOT2) p2[c] = fma(av[r][j][e], v2[j][e], p2[c]);
                }
            if (NCH == 4) {
                d1[r] = GRAD ? (p1[0] + p1[1]) + (p1[2] + p1[3]) : 0.0;
                d2[r] = DOT2 ? (p2[0] + p2[1]) + (p2[2] + p2[3]) : 0.0;
            } else {
                d1[r] = GRAD ? p1[0] + p1[1] : 0.0;
                d2[r] = DOT2 ? p2[0] + p2[1] : 0.0;
            }
        }
        const int par = s & 1;
        if (GRAD && DOT2) {
            // one butterfly for both dots (10 instead of 20 SHFL per row): after the first exchange the lower half-warp carries the
            // pair sums of dot 1 and the upper half those of dot 2; the remaining levels stay inside
            // a half.  Operands and their order are those of two separate butterflies: same bits.
            const bool upper = (lane & 16) != 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double keep = upper ? d2[r] : d1[r];
                const double send = upper ? d1[r] : d2[r];
                double v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (blane == r) {
                    // warp 0 folds -b_i into its partial so that the ordered sum below yields r_i
                    const double bi = (warp == 0) ? b_cur : 0.0;
                    sm.red[par][r][warp][upper ? 1 : 0] = v - bi;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (GRAD) d1[r] = fos_warp_sum(d1[r]);
                if (DOT2) d2[r] = fos_warp_sum(d2[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (lane == r) {
                    const double bi = (warp == 0) ? b_cur : 0.0;
                    *reinterpret_cast<double2*>(&sm.red[par][r][warp][0]) =
                        make_double2(GRAD ? d1[r] - bi : 0.0, DOT2 ? d2[r] - bi : 0.0);
                }
            }
        }
        __syncthreads();
        // Every consumer has copied stage s into registers before arriving at the barrier, so
        // its slot is free: thread 0 refills it with stage s + nstage (no empty barriers needed).
        if (tid == 0 && (WRAP || s + nstage < nst)) {
            const int ns = (s + nstage < nst) ? s + nstage : s + nstage - nst;  // >= nst: stage of the next pass
            const long long rs = lo + static_cast<long long>(ns) * R;
            const int nr = static_cast<int>(min(static_cast<long long>(R), hi - rs));
            const uint32_t bytes = static_cast<uint32_t>(nr) * row_bytes;
            mbar_expect_tx(&sm.full_bar[slot], bytes);
            bulk_g2s(ring + static_cast<size_t>(slot) * stage_bytes,
                     static_cast<const unsigned char*>(a.A) + static_cast<size_t>(rs) * row_bytes, bytes,
                     &sm.full_bar[slot], pol);
        }

        // ordered (pairwise tree) sum over the NW warp partials; all loads issue first.  The second
        // dot's total and the squared-residual sums are only ever stored by thread 0, so only warp 0
        // forms them.
        const bool sums_here = (warp == 0);
        double r1[R], r2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double2 pr[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) pr[w] = *reinterpret_cast<const double2*>(&sm.red[par][r][w][0]);
#pragma unroll
            for (int span = 1; span < NW; span *= 2)
#pragma unroll
                for (int w = 0; w + span < NW; w += 2 * span) {
                    if (GRAD) pr[w].x += pr[w + span].x;
                    if (DOT2 && sums_here) pr[w].y += pr[w + span].y;
                }
            r1[r] = pr[0].x;
            r2[r] = pr[0].y;
        }
        if (GRAD) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < NV; ++j)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[j][e] = fma(r1[r], av[r][j][e], acc[j][e]);
        }
        if (sums_here) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < rows) {
                    if (GRAD) s1 = fma(r1[r], r1[r], s1);
                    if (DOT2) s2 = fma(r2[r], r2[r], s2);
                }
            }
        }
        if (QREC && q_here) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                r_p[r] = r1[r];
                q_p[r] = q_old[r];
            }
            rows_p = rows;
        }
        b_cur = b_nxt;
        b_nxt = b_far;
        if (++slot == nstage) {
            slot = 0;
            parity ^= 1u;
        }
    }

    if (QREC && q_here) {   // the last stage's rows
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < rows_p) {
                const double qn = fma(beta_y, q_p[r], r_p[r]) * q_inv;
                s2 = fma(qn, qn, s2);
                if (lane == 0) q_wr[lo + static_cast<long long>(nst - 1) * R + r] = qn;
            }
        }
    }
    slot_io = slot;
    parity_io = parity;
    if (GRAD) {
        double* out = a.partial_g + static_cast<size_t>(cta) * a.ldv;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c0 = VEC * (tid + NT * j);
#pragma unroll
            for (int e = 0; e < VEC; e += 2)
                if (c0 + e < a.ldv) *reinterpret_cast<double2*>(out + c0 + e) = make_double2(acc[j][e], acc[j][e + 1]);
        }
    }
    if (QREC && tid == (NW - 1) * 32) a.partial_s[2 * cta + 1] = s2;
    if (tid == 0) {
        a.partial_s[2 * cta + 0] = s1;
        if (!QREC) a.partial_s[2 * cta + 1] = s2;
        if (a.cta_times) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            // end stamp in the low 48 bits, SM id above (debug only)
            a.cta_times[2 * cta + 1] = (fos_globaltimer() & 0xFFFFFFFFFFFFull) | (static_cast<unsigned long long>(smid) << 48);
        }
    }
}

// LITE: gradient-only build (modes GM_GRAD [| GM_NOB]); used for wide rows (d > 4096) by the loops
// that never need the second dot (L-BFGS, power iteration, fos_grad), where dropping v2 lets 256
// threads own 32 columns each -- half the per-element reduction cost of the 512-thread build.
template <typename T, int NT, int CPT, int R, bool LITE = false>
__global__ void __launch_bounds__(NT, 1)
grad_stream_kernel(const GradArgs a, int stage_bytes, int nstage) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(16) StreamSmem<NW, R> sm;

    const int tid = threadIdx.x;
    __shared__ int s_cta;
    // Partition slot of this CTA.  With one persistent CTA per SM the slot can be tied to the SM
    // the CTA landed on (the block scheduler's blockIdx -> SM map changes from launch to launch),
    // which lets the row blocks be weighted by the SMs' measured streaming rates.  %smid is only a
    // preference: the slot is CLAIMED with an atomic exchange of this launch's pass number, and a
    // CTA whose preferred slot is already taken (two CTAs on one SM because other work holds some
    // SMs, non-contiguous SM ids) walks on to the next free one -- P CTAs, P slots, so every row
    // block is processed exactly once per pass whatever the placement.
    uint64_t pol = 0;
    if (tid == 0) {
        int cta0 = blockIdx.x;
        if (a.sm_slot != nullptr) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int P = gridDim.x;
            int s = a.sm_slot[smid & 255u];
            for (int k = 0; k < P; ++k, s = (s + 1 == P) ? 0 : s + 1)
                if (atomicExch(&a.slot_claim[s], a.pass_no) != a.pass_no) break;
            cta0 = s;
        }
        s_cta = cta0;
        // static contiguous row partition (table built once per design): deterministic summation
        // order, bit-reproducible results for the lifetime of the design.
        const long long lo0 = a.row_lo[cta0], hi0 = a.row_lo[cta0 + 1];
        const int nst0 = static_cast<int>((hi0 - lo0 + R - 1) / R);
        for (int s = 0; s < nstage; ++s) mbar_init(&sm.full_bar[s], 1);
        mbar_fence_init();
        // prologue: fill the ring (thread 0 is also the only thread that refills it later)
        pol = l2_evict_first_policy();
        const size_t row_bytes = static_cast<size_t>(a.lda) * sizeof(T);
        for (int s = 0; s < nstage && s < nst0; ++s) {
            const long long rs = lo0 + static_cast<long long>(s) * R;
            const int nr = static_cast<int>(min(static_cast<long long>(R), hi0 - rs));
            const uint32_t bytes = static_cast<uint32_t>(nr * row_bytes);
            mbar_expect_tx(&sm.full_bar[s], bytes);
            bulk_g2s(ring + static_cast<size_t>(s) * stage_bytes,
                     static_cast<const unsigned char*>(a.A) + static_cast<size_t>(rs) * row_bytes, bytes,
                     &sm.full_bar[s], pol);
        }
    }
    __syncthreads();
    const int cta = s_cta;
    const long long lo = a.row_lo[cta];
    const long long hi = a.row_lo[cta + 1];
    const int nst = static_cast<int>((hi - lo + R - 1) / R);

    // The ring fill above only touches A, which no kernel ever writes: under programmatic
    // dependent launch it overlaps the tail of the previous epilogue.  Everything below reads
    // what that epilogue wrote (pass mode, v1, v2).
    fos_pdl_launch_dependents();
    fos_pdl_wait();
    const int mode = (a.mode_override >= 0) ? a.mode_override : a.ctrl->g_mode;
    if ((mode & (GM_GRAD | GM_DOT2 | GM_PROBE)) == 0) {
        // nothing to do (solver already finished): drain the copies in flight before exiting
        if (tid == 0)
            for (int s = 0; s < nstage && s < nst; ++s) mbar_wait(&sm.full_bar[s], 0);
        return;
    }
    if (cta == 0 && tid == 0) a.ctrl->pass_t0 = fos_globaltimer();
    if (a.cta_times && tid == 0) a.cta_times[2 * cta] = fos_globaltimer() & 0xFFFFFFFFFFFFull;

    // ===== consumers =====
    if (mode & GM_PROBE) {
        // streaming ceiling probe: same ring, same bytes, no arithmetic
        int slot = 0;
        uint32_t parity = 0;
        for (int s = 0; s < nst; ++s) {
            mbar_wait(&sm.full_bar[slot], parity);
            __syncthreads();
            if (tid == 0 && s + nstage < nst) {
                const size_t row_bytes = static_cast<size_t>(a.lda) * sizeof(T);
                const long long rs = lo + static_cast<long long>(s + nstage) * R;
                const int nr = static_cast<int>(min(static_cast<long long>(R), hi - rs));
                const uint32_t bytes = static_cast<uint32_t>(nr * row_bytes);
                mbar_expect_tx(&sm.full_bar[slot], bytes);
                bulk_g2s(ring + static_cast<size_t>(slot) * stage_bytes,
                         static_cast<const unsigned char*>(a.A) + static_cast<size_t>(rs) * row_bytes, bytes,
                         &sm.full_bar[slot], pol);
            }
            if (++slot == nstage) {
                slot = 0;
                parity ^= 1u;
            }
        }
        return;
    }
    const bool use_b = !(mode & GM_NOB);
    int slot0 = 0;
    uint32_t par0 = 0;
    if (LITE) {
        stream_consume<T, NT, CPT, R, true, false>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot0, par0);
        return;
    }
    switch (mode & (GM_GRAD | GM_DOT2)) {
        case GM_GRAD:
            if ((mode & GM_QREC) && a.qres != nullptr)
                stream_consume<T, NT, CPT, R, true, false, false, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol,
                                                                        cta, slot0, par0, a.ctrl->beta_y, a.ctrl->n_grad_calls);
            else
                stream_consume<T, NT, CPT, R, true, false>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot0, par0);
            break;
        case GM_DOT2:
            stream_consume<T, NT, CPT, R, false, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot0, par0);
            break;
        default:
            stream_consume<T, NT, CPT, R, true, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot0, par0);
            break;
    }
}


// ------------------------------------------------------------------------------------------
// Persistent solve kernel: ONE launch runs a whole proximal-gradient solve (fista / fista_delta /
// ista of the reference, iterative_solvers.py:65-344).  Every pass is the streaming loop above
// followed, in the same kernel, by what used to be a second launch (epilogue_kernel):
//
//   1. every CTA publishes its partial A^T r (and residual norms)            -> grid barrier 1
//   2. the d columns are sliced over ALL CTAs (14 column pairs each at d = 4096): a CTA sums its
//      slice over the P partials (two-level fixed order: groups of 8, then the groups)
//   3. row-sharded designs: the slice is PUSHED into every rank's window over NVLink, followed by one
//      release-flag per (rank, CTA); a CTA waits only for ITS slice from the peers and sums the
//      ranks in rank order (bit-identical on every rank) -- the all-reduce of the d-vector is done
//      by 148 independent slice exchanges that overlap each other
//   4. soft threshold / Armijo candidate for the slice (pg_logic.cuh), slice sums  -> grid barrier 2
//   5. every CTA forms the same scalars, takes the same decision, applies the momentum step to its
//      slice; CTA 0 writes the control block and the history scalars          -> grid barrier 3
//
// While 1-5 run, the bulk-copy ring is already refilling with the first stages of the next pass
// (WRAP in stream_consume), so HBM idles for a few microseconds per pass instead of a launch
// boundary + an epilogue launch.  The host launches once and synchronises once.
//
// All cross-CTA data is read through L2 (ld.global.cg / volatile) -- the L1 of an SM may hold the
// previous pass's lines.  The grid barriers are monotonic counters (never reset); every wait has a
// deadline and an abort flag so that a lost peer or a non-resident CTA ends the launch with an
// error instead of hanging the device.  Requires all CTAs co-resident: grid = one CTA per SM,
// launched cooperatively.
// ------------------------------------------------------------------------------------------
constexpr int TAIL_GS = 8;                                   // partials summed per group
constexpr int TAIL_NG = (FOS_MAX_PARTS + TAIL_GS - 1) / TAIL_GS;
constexpr int TAIL_PW = 28;                                  // most column pairs per CTA slice (ldv 8192 / 148 CTAs)

struct TailSmem {
    double2 gsum[TAIL_NG][TAIL_PW];   // group sums of the slice
    double2 ssum[TAIL_NG];            // group sums of the residual-norm pair
    double red[FOS_NSCAL][TAIL_PW];   // per-thread slice sums -> CTA sums
    double grp[TAIL_NG][FOS_NSCAL];   // group sums of the per-CTA scalars
    double tot[FOS_NSCAL];
    double s12[2];
    int ok;
    // CTA 0 only: running totals of the timing fields of the control block (stored, never re-read from HBM)
    unsigned long long acc_epi, acc_xchg, acc_grad;
    int acc_passes;
    // copy of the control block taken at the top of a pass by one coalesced warp read: 148 CTAs x ~30
    // scalar reads of the same three lines after barrier 1 were a measurable L2 hot spot (3 us)
    __align__(16) FosCtrl ctrl;
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_relaxed_sys_v2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_v2(double* p, double2 v) {
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

constexpr unsigned long long GRID_WAIT_NS = 20ull * 1000000000ull;   // a CTA of this grid never arrived
constexpr unsigned long long PEER_WAIT_NS = 60ull * 1000000000ull;   // a peer rank never arrived

// Arrive on a monotonic grid counter and wait until `target` arrivals.  Returns false when the launch
// is being abandoned (deadline passed here, or another CTA raised the abort flag).
__device__ __forceinline__ bool grid_barrier(unsigned long long* ctr, unsigned long long target, FosGridSync* gs,
                                             TailSmem& ts) {
    __syncthreads();  // every thread's writes of this phase are ordered before thread 0's release
    if (threadIdx.x == 0) {
        // release / acquire at gpu scope on the counter itself (cumulative over the CTA's writes through the
        // bar.sync above; the readers bypass L1 with ld.cg): no MEMBAR on either side
        red_release_gpu_add(ctr, 1ull);
        bool ok = true;
        if (ld_acquire_gpu_u64(ctr) < target) {
            const unsigned long long t0 = fos_globaltimer();
            unsigned spins = 0;
            while (ld_acquire_gpu_u64(ctr) < target) {
                if ((++spins & 1023u) == 0u) {
                    if (*reinterpret_cast<volatile int*>(&gs->abort) != 0 || fos_globaltimer() - t0 > GRID_WAIT_NS) {
                        atomicExch(&gs->abort, 1);
                        ok = false;
                        break;
                    }
                }
            }
        }
        ts.ok = ok ? 1 : 0;
    }
    __syncthreads();
    return ts.ok != 0;
}

// two-level fixed-order sum of P values per thread-item: groups of TAIL_GS consecutive partials, then
// the groups in order (deterministic whatever CTA runs it)
template <typename T, int NT, int CPT, int R, bool LITE>
__global__ void __launch_bounds__(NT, 1)
solve_stream_kernel(const GradArgs a, const EpiArgs e, FosGridSync* gs, int stage_bytes, int nstage, long long max_passes) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(16) StreamSmem<NW, R> sm;
    __shared__ __align__(16) TailSmem ts;
    __shared__ int s_cta;

    const int tid = threadIdx.x;
    const int P = gridDim.x;
    uint64_t pol = 0;
    if (tid == 0) {
        int cta0 = blockIdx.x;
        if (a.sm_slot != nullptr) {  // SM-indexed partition: claimed once, kept for every pass of the launch
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            int s = a.sm_slot[smid & 255u];
            for (int k = 0; k < P; ++k, s = (s + 1 == P) ? 0 : s + 1)
                if (atomicExch(&a.slot_claim[s], a.pass_no) != a.pass_no) break;
            cta0 = s;
        }
        s_cta = cta0;
        const long long lo0 = a.row_lo[cta0], hi0 = a.row_lo[cta0 + 1];
        const int nst0 = static_cast<int>((hi0 - lo0 + R - 1) / R);
        for (int s = 0; s < nstage; ++s) mbar_init(&sm.full_bar[s], 1);
        mbar_fence_init();
        pol = l2_evict_first_policy();
        const size_t row_bytes = static_cast<size_t>(a.lda) * sizeof(T);
        for (int s = 0; s < nstage && s < nst0; ++s) {  // the host guarantees nst0 >= nstage
            const long long rs = lo0 + static_cast<long long>(s) * R;
            const int nr = static_cast<int>(min(static_cast<long long>(R), hi0 - rs));
            const uint32_t bytes = static_cast<uint32_t>(nr * row_bytes);
            mbar_expect_tx(&sm.full_bar[s], bytes);
            bulk_g2s(ring + static_cast<size_t>(s) * stage_bytes,
                     static_cast<const unsigned char*>(a.A) + static_cast<size_t>(rs) * row_bytes, bytes,
                     &sm.full_bar[s], pol);
        }
    }
    __syncthreads();
    const int cta = s_cta;
    const long long lo = a.row_lo[cta];
    const long long hi = a.row_lo[cta + 1];
    const bool leader = (cta == 0 && tid == 0);
    FosCtrl* C = e.ctrl;
    const volatile FosCtrl* VC = C;

    // column slice of this CTA (in column pairs)
    const int pairs = e.ldv / 2;
    const int pw = (pairs + P - 1) / P;                       // <= TAIL_PW (host-checked)
    const int p0 = min(cta * pw, pairs);
    const int npair = min(pw, pairs - p0);
    const int NG = (P + TAIL_GS - 1) / TAIL_GS;
    const int lane_ws = e.ldv + FOS_WIN_PAD;                  // doubles per (slot, source rank) in the push region

    int slot = 0;
    uint32_t parity = 0;
    const unsigned long long gen0 = ld_acquire_gpu_u64(&gs->gen);   // passes this design's kernel has completed before
    bool alive = true;
    long long pass = 0;
    if (leader) {
        ts.acc_epi = VC->epi_ns;
        ts.acc_xchg = VC->xchg_ns;
        ts.acc_grad = VC->grad_ns;
        ts.acc_passes = VC->n_passes;
    }
    // the state the host wrote before the launch: one coalesced read per CTA into shared memory.  From
    // here on every CTA keeps its copy current itself -- the decision of a pass is computed identically
    // in every CTA -- so the control block in HBM is only written (by CTA 0, for the host), never re-read.
    if (tid < 32) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(C);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&ts.ctrl);
        for (int w = tid; w < static_cast<int>(sizeof(FosCtrl) / 8); w += 32) dst[w] = __ldcg(src + w);
    }
    __syncthreads();
    for (; pass < max_passes; ++pass) {
        const int mode = ts.ctrl.g_mode;
        if ((mode & (GM_GRAD | GM_DOT2)) == 0) break;
        const unsigned long long gen = gen0 + static_cast<unsigned long long>(pass);
        const unsigned long long target = (gen + 1ull) * static_cast<unsigned long long>(P);
        const unsigned long long epoch = (e.world > 1) ? *reinterpret_cast<volatile unsigned long long*>(e.peer.epoch) : 0ull;
        const unsigned long long t_pass0 = fos_globaltimer();
        if (leader) ts.ctrl.pass_t0 = t_pass0;

        // ---------------- the pass over this CTA's rows
        const bool use_b = !(mode & GM_NOB);
        if (LITE) {
            stream_consume<T, NT, CPT, R, true, false, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot, parity);
        } else {
            switch (mode & (GM_GRAD | GM_DOT2)) {
                case GM_GRAD:
                    if ((mode & GM_QREC) && a.qres != nullptr)
                        stream_consume<T, NT, CPT, R, true, false, true, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol,
                                                                               cta, slot, parity, ts.ctrl.beta_y, ts.ctrl.n_grad_calls);
                    else
                        stream_consume<T, NT, CPT, R, true, false, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot, parity);
                    break;
                case GM_DOT2:
                    stream_consume<T, NT, CPT, R, false, true, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot, parity);
                    break;
                default:
                    stream_consume<T, NT, CPT, R, true, true, true>(a, sm, ring, stage_bytes, nstage, lo, hi, use_b, pol, cta, slot, parity);
                    break;
            }
        }
        // ---------------- barrier 1: all partials of this pass are in L2
        const unsigned long long t_c1 = fos_globaltimer();
        if (!grid_barrier(&gs->arrive0, target, gs, ts)) { alive = false; break; }
        const unsigned long long t_b1 = fos_globaltimer();

        const PgIn in = pg_read<false>(&ts.ctrl);
        const int phase = in.phase;
        const bool has_grad = (phase == PH_GRAD);
        bool comm_ok = true;

        // ---------------- local sums over the P partials: column slice (PH_GRAD) and residual norms (every CTA)
        if (has_grad) {
            for (int item = tid; item < npair * NG; item += NT) {
                const int p = item % npair, q = item / npair;
                const double* src = a.partial_g + static_cast<size_t>(q) * TAIL_GS * a.ldv + 2 * (p0 + p);
                const int cnt = min(TAIL_GS, P - q * TAIL_GS);
                double2 v[TAIL_GS];
#pragma unroll
                for (int u = 0; u < TAIL_GS; ++u)
                    v[u] = (u < cnt) ? __ldcg(reinterpret_cast<const double2*>(src + static_cast<size_t>(u) * a.ldv))
                                     : make_double2(0.0, 0.0);
                double2 acc = v[0];
#pragma unroll
                for (int u = 1; u < TAIL_GS; ++u)
                    if (u < cnt) {
                        acc.x += v[u].x;
                        acc.y += v[u].y;
                    }
                ts.gsum[q][p] = acc;
            }
        }
        if (tid >= NT - 32 && tid - (NT - 32) < NG) {
            const int q = tid - (NT - 32);
            const int cnt = min(TAIL_GS, P - q * TAIL_GS);
            const double* src = a.partial_s + static_cast<size_t>(q) * TAIL_GS * 2;
            double2 v[TAIL_GS];   // all loads in flight before the first add (one L2 round trip, not eight)
#pragma unroll
            for (int u = 0; u < TAIL_GS; ++u)
                v[u] = (u < cnt) ? __ldcg(reinterpret_cast<const double2*>(src + 2 * u)) : make_double2(0.0, 0.0);
            double2 acc = v[0];
#pragma unroll
            for (int u = 1; u < TAIL_GS; ++u)
                if (u < cnt) {
                    acc.x += v[u].x;
                    acc.y += v[u].y;
                }
            ts.ssum[q] = acc;
        }
        __syncthreads();
        double2 gslice = make_double2(0.0, 0.0);   // thread tid < npair: its column pair of the (rank-local) gradient
        if (has_grad && tid < npair) {
            gslice = ts.gsum[0][tid];
            for (int q = 1; q < NG; ++q) {
                gslice.x += ts.gsum[q][tid].x;
                gslice.y += ts.gsum[q][tid].y;
            }
        }
        double s1 = 0.0, s2 = 0.0;
        if (tid == NT - 1) {
            double2 acc = ts.ssum[0];
            for (int q = 1; q < NG; ++q) {
                acc.x += ts.ssum[q].x;
                acc.y += ts.ssum[q].y;
            }
            ts.s12[0] = acc.x;
            ts.s12[1] = acc.y;
        }
        __syncthreads();
        s1 = ts.s12[0];
        s2 = ts.s12[1];

        const unsigned long long t_s1 = fos_globaltimer();
        // ---------------- row-sharded designs: slice exchange over peer memory
        unsigned long long t_x0 = 0, t_x1 = 0;
        if (e.world > 1) {
            if (leader) t_x0 = fos_globaltimer();
            const size_t slot_off = static_cast<size_t>(epoch & 1ull) * FOS_MAX_WORLD * lane_ws;
            const size_t mine = slot_off + static_cast<size_t>(e.rank) * lane_ws;
            if (has_grad && tid < npair) {
                for (int r = 0; r < e.world; ++r) st_relaxed_sys_v2(e.peer.fwin[r] + mine + 2 * (p0 + tid), gslice);
            }
            if (cta == 0 && tid == NT - 1) {
                for (int r = 0; r < e.world; ++r) st_relaxed_sys_v2(e.peer.fwin[r] + mine + e.ldv, make_double2(s1, s2));
            }
            __syncthreads();
            // CTAs that pushed something signal (rank, CTA); everybody else only listens to CTA 0's flags
            const bool pushed = (has_grad && npair > 0) || cta == 0;
            if (pushed && tid < e.world) {
                __threadfence_system();
                st_release_sys_u64(e.peer.fflag[tid] + static_cast<size_t>(e.rank) * FOS_MAX_PARTS + cta, epoch);
            }
            if (tid == 0) ts.ok = 1;
            __syncthreads();
            if (tid < e.world) {
                // my slice from rank `tid` (only when there is one), and the scalars from its CTA 0
                const unsigned long long* f_own = e.peer.fflag[e.rank] + static_cast<size_t>(tid) * FOS_MAX_PARTS + cta;
                const unsigned long long* f_sc = e.peer.fflag[e.rank] + static_cast<size_t>(tid) * FOS_MAX_PARTS;
                const bool need_own = has_grad && npair > 0;
                const unsigned long long t0 = fos_globaltimer();
                unsigned spins = 0;
                while ((need_own && ld_acquire_sys_u64(f_own) < epoch) || ld_acquire_sys_u64(f_sc) < epoch) {
                    if ((++spins & 1023u) == 0u) {
                        if (*reinterpret_cast<volatile int*>(&gs->abort) != 0 || fos_globaltimer() - t0 > PEER_WAIT_NS) {
                            atomicExch(&gs->abort, 1);
                            ts.ok = 0;
                            break;
                        }
                    }
                }
            }
            __syncthreads();
            comm_ok = ts.ok != 0;
            if (comm_ok) {
                if (has_grad && tid < npair) {
                    double2 acc = make_double2(0.0, 0.0);
                    double2 v[FOS_MAX_WORLD];
#pragma unroll
                    for (int r = 0; r < FOS_MAX_WORLD; ++r)
                        if (r < e.world)
                            v[r] = ld_relaxed_sys_v2(e.peer.fwin[e.rank] + slot_off + static_cast<size_t>(r) * lane_ws + 2 * (p0 + tid));
#pragma unroll
                    for (int r = 0; r < FOS_MAX_WORLD; ++r)
                        if (r < e.world) {
                            acc.x += v[r].x;
                            acc.y += v[r].y;
                        }
                    gslice = acc;
                }
                if (tid == NT - 1) {
                    double2 acc = make_double2(0.0, 0.0);
                    for (int r = 0; r < e.world; ++r) {
                        const double2 v = ld_relaxed_sys_v2(e.peer.fwin[e.rank] + slot_off + static_cast<size_t>(r) * lane_ws + e.ldv);
                        acc.x += v.x;
                        acc.y += v.y;
                    }
                    ts.s12[0] = acc.x;
                    ts.s12[1] = acc.y;
                }
            }
            __syncthreads();
            s1 = ts.s12[0];
            s2 = ts.s12[1];
            if (leader) t_x1 = fos_globaltimer();
        }

        const unsigned long long t_x2 = fos_globaltimer();
        // ---------------- elementwise 1 on the slice + slice sums
        bool accept;
        double t_new;
        pg_armijo(in, s2, accept, t_new);
        double sums[FOS_NSCAL];
#pragma unroll
        for (int k = 0; k < FOS_NSCAL; ++k) sums[k] = 0.0;
        double2 cand = make_double2(0.0, 0.0), xk = make_double2(0.0, 0.0);   // stay in registers for elementwise 2
        if (tid < npair) {
            const int c = 2 * (p0 + tid);
            if (phase == PH_GRAD) pg_elem1_grad<true>(e, in, c, gslice, sums, &cand, &xk);
            else if (phase == PH_TRIAL) pg_elem1_trial<true>(e, in, c, accept, t_new, sums, &cand, &xk);
        }
        if (tid < TAIL_PW) {
#pragma unroll
            for (int k = 0; k < FOS_NSCAL; ++k) ts.red[k][tid] = sums[k];
        }
        __syncthreads();
        if (tid < FOS_NSCAL) {
            double t = 0.0;
            for (int p = 0; p < npair; ++p) t += ts.red[tid][p];
            gs->scal[cta][tid] = t;
        }
        // ---------------- barrier 2: every CTA's slice sums are in L2
        const unsigned long long t_e1 = fos_globaltimer();
        if (!grid_barrier(&gs->arrive1, target, gs, ts)) { alive = false; break; }
        const unsigned long long t_b2 = fos_globaltimer();
        for (int item = tid; item < NG * FOS_NSCAL; item += NT) {
            const int k = item % FOS_NSCAL, q = item / FOS_NSCAL;
            const int cnt = min(TAIL_GS, P - q * TAIL_GS);
            double v[TAIL_GS];
#pragma unroll
            for (int u = 0; u < TAIL_GS; ++u) v[u] = (u < cnt) ? __ldcg(&gs->scal[q * TAIL_GS + u][k]) : 0.0;
            double t = v[0];
#pragma unroll
            for (int u = 1; u < TAIL_GS; ++u)
                if (u < cnt) t += v[u];
            ts.grp[q][k] = t;
        }
        __syncthreads();
        if (tid < FOS_NSCAL) {
            double t = 0.0;
            for (int q = 0; q < NG; ++q) t += ts.grp[q][tid];
            ts.tot[tid] = t;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FOS_NSCAL; ++k) sums[k] = ts.tot[k];

        // ---------------- scalar logic (identical everywhere), elementwise 2 on the slice, commit
        const PgOut o = pg_decide(in, sums, s1, s2, accept, t_new);
        if (o.do_update && tid < npair) pg_elem2_regs(e, in, o, 2 * (p0 + tid), cand, xk);
        if (tid == 0) pg_apply_state(&ts.ctrl, o, comm_ok);   // this CTA's copy: what the next pass starts from
        if (leader) {
            if (e.world > 1) *e.peer.epoch = epoch + 1;
            ts.acc_grad += t_b1 - t_pass0;
            ts.acc_epi += fos_globaltimer() - t_b1;
            ts.acc_xchg += t_x1 - t_x0;
            ts.acc_passes += 1;
            pg_commit<true>(C, e.hist, in, o, comm_ok, t_b1, 0ull);
            C->grad_ns = ts.acc_grad;
            C->epi_ns = ts.acc_epi;
            C->xchg_ns = ts.acc_xchg;
            C->n_passes = ts.acc_passes;
            gs->gen = gen + 1ull;
        }
        // ---------------- barrier 3: y, x_k and the control block are complete before anyone starts the next pass
        const unsigned long long t_e2 = fos_globaltimer();
        if (!grid_barrier(&gs->arrive2, target, gs, ts)) { alive = false; break; }
        if (leader) {
            const unsigned long long t_b3 = fos_globaltimer();
            gs->prof[0] += t_c1 - t_pass0;
            gs->prof[1] += t_b1 - t_c1;
            gs->prof[2] += t_s1 - t_b1;
            gs->prof[3] += t_x2 - t_s1;
            gs->prof[4] += t_e1 - t_x2;
            gs->prof[5] += t_b2 - t_e1;
            gs->prof[6] += t_e2 - t_b2;
            gs->prof[7] += t_b3 - t_e2;
            gs->prof[8] += 1;
        }
        if (!comm_ok) break;
    }

    // drain the copies still in flight (the first stages of a pass that will not run here)
    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_wait(&sm.full_bar[slot], parity);
            if (++slot == nstage) {
                slot = 0;
                parity ^= 1u;
            }
        }
    }
    if (!alive && leader) {  // abandoned: the host reports the error (FOS_ERR_COMM)
        C->stop_reason = -1;
        C->phase = PH_DONE;
        C->g_mode = GM_SKIP;
    }
}

// ------------------------------------------------------------------------------------------
// generic kernel (lda <= 512): one warp per contiguous row block, lane owns columns
// lane + 32 j.  Direct coalesced global loads, two rows in flight per warp.
// ------------------------------------------------------------------------------------------
constexpr int GEN_WARPS = 8;

template <typename T, int J>
__global__ void __launch_bounds__(GEN_WARPS * 32)
grad_generic_kernel(GradArgs a) {
    __shared__ double wacc[GEN_WARPS][J * 32];
    __shared__ double wsc[GEN_WARPS][2];

    fos_pdl_launch_dependents();
    fos_pdl_wait();
    const int mode = (a.mode_override >= 0) ? a.mode_override : a.ctrl->g_mode;
    if ((mode & (GM_GRAD | GM_DOT2)) == 0) return;
    const bool do_grad = mode & GM_GRAD;
    const bool do_dot2 = mode & GM_DOT2;
    const bool use_b = !(mode & GM_NOB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (blockIdx.x == 0 && tid == 0) a.ctrl->pass_t0 = fos_globaltimer();
    const long long W = static_cast<long long>(gridDim.x) * GEN_WARPS;
    const long long gw = static_cast<long long>(blockIdx.x) * GEN_WARPS + warp;
    const long long lo = (a.n * gw) / W, hi = (a.n * (gw + 1)) / W;

    double v1[J], v2[J], acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int c = lane + 32 * j;
        v1[j] = (do_grad && c < a.d) ? a.v1[c] : 0.0;
        v2[j] = (do_dot2 && c < a.d) ? a.v2[c] : 0.0;
        acc[j] = 0.0;
    }
    double s1 = 0.0, s2 = 0.0;
    const T* A = static_cast<const T*>(a.A);

    long long row = lo;
    for (; row + 1 < hi; row += 2) {
        double a0[J], a1[J];
        const T* p0 = A + row * a.lda;
        const T* p1 = p0 + a.lda;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int c = lane + 32 * j;
            a0[j] = (c < a.lda) ? static_cast<double>(__ldg(p0 + c)) : 0.0;
            a1[j] = (c < a.lda) ? static_cast<double>(__ldg(p1 + c)) : 0.0;
        }
        const double b0 = use_b ? __ldg(a.b + row) : 0.0;
        const double b1 = use_b ? __ldg(a.b + row + 1) : 0.0;
        double d10 = 0.0, d11 = 0.0, d20 = 0.0, d21 = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (do_grad) {
                d10 = fma(a0[j], v1[j], d10);
                d11 = fma(a1[j], v1[j], d11);
            }
            if (do_dot2) {
                d20 = fma(a0[j], v2[j], d20);
                d21 = fma(a1[j], v2[j], d21);
            }
        }
        if (do_grad) {
            d10 = fos_warp_sum(d10) - b0;
            d11 = fos_warp_sum(d11) - b1;
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = fma(d11, a1[j], fma(d10, a0[j], acc[j]));
            s1 = fma(d11, d11, fma(d10, d10, s1));
        }
        if (do_dot2) {
            d20 = fos_warp_sum(d20) - b0;
            d21 = fos_warp_sum(d21) - b1;
            s2 = fma(d21, d21, fma(d20, d20, s2));
        }
    }
    if (row < hi) {
        double a0[J];
        const T* p0 = A + row * a.lda;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int c = lane + 32 * j;
            a0[j] = (c < a.lda) ? static_cast<double>(__ldg(p0 + c)) : 0.0;
        }
        const double b0 = use_b ? __ldg(a.b + row) : 0.0;
        double d10 = 0.0, d20 = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (do_grad) d10 = fma(a0[j], v1[j], d10);
            if (do_dot2) d20 = fma(a0[j], v2[j], d20);
        }
        if (do_grad) {
            d10 = fos_warp_sum(d10) - b0;
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = fma(d10, a0[j], acc[j]);
            s1 = fma(d10, d10, s1);
        }
        if (do_dot2) {
            d20 = fos_warp_sum(d20) - b0;
            s2 = fma(d20, d20, s2);
        }
    }

#pragma unroll
    for (int j = 0; j < J; ++j) wacc[warp][lane + 32 * j] = acc[j];
    if (lane == 0) {
        wsc[warp][0] = s1;
        wsc[warp][1] = s2;
    }
    __syncthreads();
    if (do_grad) {
        for (int c = tid; c < a.ldv && c < J * 32; c += GEN_WARPS * 32) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < GEN_WARPS; ++w) t += wacc[w][c];
            a.partial_g[static_cast<size_t>(blockIdx.x) * a.ldv + c] = t;
        }
    }
    if (tid == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < GEN_WARPS; ++w) {
            t1 += wsc[w][0];
            t2 += wsc[w][1];
        }
        a.partial_s[2 * blockIdx.x + 0] = t1;
        a.partial_s[2 * blockIdx.x + 1] = t2;
    }
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
struct StreamCfg {
    const void* fn;
    int nt, cpt, r;
    const void* solve_fn = nullptr;  // persistent solve kernel of the same shape (full-mode builds only)
};

template <typename T, int NT, int CPT, int R, bool LITE = false>
StreamCfg make_cfg() {
    StreamCfg c;
    c.fn = reinterpret_cast<const void*>(&grad_stream_kernel<T, NT, CPT, R, LITE>);
    c.nt = NT;
    c.cpt = CPT;
    c.r = R;
    if (!LITE) c.solve_fn = reinterpret_cast<const void*>(&solve_stream_kernel<T, NT, CPT, R, false>);
    return c;
}

// gradient-only variant for wide rows, if one exists
bool pick_lite_cfg(int dtype, int lda, StreamCfg* out) {
    if (lda <= 4096 || lda > 8192) return false;
    *out = (dtype == FOS_F64) ? make_cfg<double, 256, 32, 1, true>() : make_cfg<float, 256, 32, 1, true>();
    return true;
}

bool pick_stream_cfg(int dtype, int lda, StreamCfg* out) {
    // FOS_ROWS_X2=1 doubles the rows handled per barrier (A/B switch for tuning)
    const char* e = getenv("FOS_ROWS_X2");
    const bool x2 = e && e[0] == '1';
    if (dtype == FOS_F64) {
        if (lda <= 512) *out = make_cfg<double, 256, 2, 8>();
        else if (lda <= 1024) *out = make_cfg<double, 256, 4, 4>();
        else if (lda <= 2048) *out = x2 ? make_cfg<double, 256, 8, 4>() : make_cfg<double, 256, 8, 2>();
        else if (lda <= 4096) *out = x2 ? make_cfg<double, 256, 16, 2>() : make_cfg<double, 256, 16, 1>();
        else if (lda <= 8192) *out = make_cfg<double, 512, 16, 1>();
        else return false;
    } else {
        if (lda <= 1024) *out = make_cfg<float, 256, 4, 8>();
        else if (lda <= 2048) *out = make_cfg<float, 256, 8, 4>();
        else if (lda <= 4096) *out = make_cfg<float, 256, 16, 2>();
        else if (lda <= 8192) *out = make_cfg<float, 512, 16, 1>();
        else return false;
    }
    return true;
}

template <typename T>
const void* pick_generic(int lda) {
    if (lda <= 32) return reinterpret_cast<const void*>(&grad_generic_kernel<T, 1>);
    if (lda <= 64) return reinterpret_cast<const void*>(&grad_generic_kernel<T, 2>);
    if (lda <= 128) return reinterpret_cast<const void*>(&grad_generic_kernel<T, 4>);
    if (lda <= 256) return reinterpret_cast<const void*>(&grad_generic_kernel<T, 8>);
    return reinterpret_cast<const void*>(&grad_generic_kernel<T, 16>);
}

constexpr int SMEM_RING_BUDGET = 200 * 1024;

}  // namespace

// Chooses the kernel and the number of partial rows.  Called once per design, before the
// workspaces are allocated (n_parts sizes them).
int fos_grad_plan(fos_design* h) {
    const char* force = getenv("FOS_FORCE_KERNEL");  // "generic" | "stream" (tests)
    bool want_stream = h->lda > 512;
    if (force && std::string(force) == "generic" && h->lda <= 512) want_stream = false;
    if (force && std::string(force) == "stream" && h->lda * (h->dtype == FOS_F64 ? 8 : 4) >= 16)
        want_stream = true;
    StreamCfg cfg;
    if (want_stream) {
        if (!pick_stream_cfg(h->dtype, h->lda, &cfg)) {
            fos_set_error("d = %d exceeds the streaming kernel's limit of 8192 columns", h->d);
            return FOS_ERR_UNSUPPORTED;
        }
        h->kern_kind = 1;
        long long per = (h->n + h->sm_count - 1) / h->sm_count;
        h->n_parts = (per >= 1) ? h->sm_count : 1;
        if (h->n < h->sm_count) h->n_parts = static_cast<int>(h->n > 0 ? h->n : 1);
        FOS_CUDA(cudaFuncSetAttribute(cfg.fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SMEM_RING_BUDGET));
        // The persistent solve kernel needs: one CTA per SM (all co-resident), every CTA's row block at
        // least as long as the ring (the ring wraps into the next pass), column slices of at most
        // TAIL_PW pairs.  FOS_FUSED=0 keeps the two-launch path (A/B, debugging).
        {
            const char* ff = getenv("FOS_FUSED");
            const int elem = (h->dtype == FOS_F64) ? 8 : 4;
            int sb = (cfg.r * h->lda * elem + 127) & ~127;
            int ns = std::min(SMEM_RING_BUDGET / sb, MAX_STAGES);
            const int pairs = h->ldv / 2;
            // (the ring depth is fitted to the shortest row block at launch time, fos_solve_stages)
            h->fused_ok = !(ff && ff[0] == '0') && cfg.solve_fn != nullptr && h->n_parts == h->sm_count &&
                          h->n_parts <= FOS_MAX_PARTS && ns >= 2 && h->n / h->n_parts >= 2LL * cfg.r &&
                          (pairs + h->n_parts - 1) / h->n_parts <= TAIL_PW;
            if (h->fused_ok) {
                int max_blocks = 0;
                cudaError_t ce = cudaFuncSetAttribute(cfg.solve_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_RING_BUDGET);
                if (ce == cudaSuccess)
                    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, cfg.solve_fn, cfg.nt,
                                                                       static_cast<size_t>(ns) * sb);
                if (ce != cudaSuccess || max_blocks < 1) {
                    cudaGetLastError();
                    h->fused_ok = false;
                }
            }
        }
        StreamCfg lite;
        const char* nl = getenv("FOS_NO_LITE");
        h->lite_ok = pick_lite_cfg(h->dtype, h->lda, &lite) && !(nl && nl[0] == '1');
        if (h->lite_ok)
            FOS_CUDA(cudaFuncSetAttribute(lite.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_RING_BUDGET));
    } else {
        h->kern_kind = 0;
        long long want = (h->n + GEN_WARPS * 16 - 1) / (GEN_WARPS * 16);
        long long cap = static_cast<long long>(h->sm_count) * 4;
        h->n_parts = static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
    }
    return FOS_OK;
}

// Ring depth of the persistent kernel for this design's current row partition: as deep as the
// shared-memory budget allows, but never deeper than the shortest row block has stages (the ring
// wraps from the end of a pass into the start of the next one).  0: the design does not qualify.
int fos_solve_stages(const fos_design* h) {
    if (!h->fused_ok || h->kern_kind != 1 || h->gsync == nullptr) return 0;
    StreamCfg cfg;
    if (!pick_stream_cfg(h->dtype, h->lda, &cfg)) return 0;
    const int elem = (h->dtype == FOS_F64) ? 8 : 4;
    const int stage_bytes = (cfg.r * h->lda * elem + 127) & ~127;
    long long min_rows = h->n;
    for (int c = 0; c < h->n_parts; ++c) min_rows = std::min(min_rows, h->row_lo_host[c + 1] - h->row_lo_host[c]);
    const long long ns = std::min<long long>(std::min(SMEM_RING_BUDGET / stage_bytes, MAX_STAGES), min_rows / cfg.r);
    return ns >= 2 ? static_cast<int>(ns) : 0;
}

// Row-sharded designs: every rank must take the SAME path (the persistent kernel pushes slices, the
// epilogue kernel pulls vectors; mixing them deadlocks until the wait deadlines fire).  The ranks therefore
// agree once, when the windows are attached (multigpu.attach): a rank is a candidate if its design
// qualifies with a margin that survives the rate-weighted re-partition (blocks between 0.3 x and 1.6 x the
// mean from 4 ranks on), and the kernel is used only if ALL ranks are candidates.
extern "C" int fos_design_solve_kernel_ok(const fos_design* h, int world, int* ok) {
    FOS_REQUIRE(h && ok, "null pointer argument");
    *ok = 0;
    if (!h->fused_ok || h->kern_kind != 1 || h->gsync == nullptr) return FOS_OK;
    StreamCfg cfg;
    if (!pick_stream_cfg(h->dtype, h->lda, &cfg)) return FOS_OK;
    const char* e = getenv("FOS_BALANCE");
    const bool may_rebalance = (world >= 4 && !(e && e[0] == '0')) || (e && e[0] == '1');
    const long long need = (may_rebalance ? 8LL : 2LL) * cfg.r;
    *ok = (h->n / h->n_parts >= need) ? 1 : 0;
    return FOS_OK;
}

extern "C" int fos_design_solve_kernel_disable(fos_design* h) {
    FOS_REQUIRE(h, "null design");
    h->fused_ok = false;
    return FOS_OK;
}

int fos_launch_solve(fos_design* h, const FosHist& hist, long long max_passes) {
    int nstage = fos_solve_stages(h);
    if (nstage == 0) {
        fos_set_error("this design does not qualify for the persistent solve kernel");
        return FOS_ERR_UNSUPPORTED;
    }
    StreamCfg cfg;
    pick_stream_cfg(h->dtype, h->lda, &cfg);
    GradArgs a;
    a.A = h->A;
    a.b = h->b;
    a.v1 = h->y;
    a.v2 = h->xc;
    a.partial_g = h->partial_g;
    a.partial_s = h->partial_s;
    a.ctrl = h->ctrl;
    a.n = h->n;
    a.d = h->d;
    a.lda = h->lda;
    a.ldv = h->ldv;
    a.mode_override = -1;
    a.cta_times = nullptr;
    a.qres = h->qres;
    a.row_lo = h->row_lo;
    a.sm_slot = h->sm_slot;
    a.slot_claim = a.sm_slot ? reinterpret_cast<unsigned*>(h->sm_slot + 256) : nullptr;
    a.pass_no = static_cast<unsigned>(h->launches & 0x7fffffff);
    EpiArgs e{};
    e.ctrl = h->ctrl;
    e.hist = hist;
    e.partial_g = h->partial_g;
    e.partial_s = h->partial_s;
    e.n_parts = h->n_parts;
    e.d = h->d;
    e.ldv = h->ldv;
    e.op = EOP_PG;
    e.y = h->y;
    e.xc = h->xc;
    e.xk = h->xk;
    e.g = h->g;
    e.world = h->world;
    e.rank = h->rank;
    e.peer = h->peer;
    const int elem = (h->dtype == FOS_F64) ? 8 : 4;
    int stage_bytes = (cfg.r * h->lda * elem + 127) & ~127;
    FosGridSync* gs = h->gsync;
    void* params[6] = {&a, &e, &gs, &stage_bytes, &nstage, &max_passes};
    // cooperative launch: the driver refuses the launch unless all CTAs can be resident at once
    FOS_CUDA(fos_launch_ex(cfg.solve_fn, dim3(h->n_parts), dim3(cfg.nt), static_cast<size_t>(nstage) * stage_bytes,
                           h->stream, params, false, -1));
    h->launches++;
    return FOS_OK;
}

int fos_launch_grad(fos_design* h, int mode_override) {
    GradArgs a;
    a.A = h->A;
    a.b = h->b;
    a.v1 = h->y;
    a.v2 = h->xc;
    a.partial_g = h->partial_g;
    a.partial_s = h->partial_s;
    a.ctrl = h->ctrl;
    a.n = h->n;
    a.d = h->d;
    a.lda = h->lda;
    a.ldv = h->ldv;
    a.mode_override = mode_override;
    a.cta_times = h->cta_times;
    a.qres = h->qres;
    a.row_lo = h->row_lo;
    a.sm_slot = (h->kern_kind == 1) ? h->sm_slot : nullptr;
    a.slot_claim = a.sm_slot ? reinterpret_cast<unsigned*>(h->sm_slot + 256) : nullptr;
    a.pass_no = static_cast<unsigned>(h->launches & 0x7fffffff);
    void* params[3];
    params[0] = &a;
    if (h->kern_kind == 1) {
        StreamCfg cfg;
        pick_stream_cfg(h->dtype, h->lda, &cfg);
        const bool grad_only = (mode_override >= 0) ? ((mode_override & (GM_GRAD | GM_DOT2 | GM_PROBE)) == GM_GRAD)
                                                    : h->grad_only_hint;
        if (grad_only && h->lite_ok) pick_lite_cfg(h->dtype, h->lda, &cfg);
        const int elem = (h->dtype == FOS_F64) ? 8 : 4;
        int stage_bytes = cfg.r * h->lda * elem;
        stage_bytes = (stage_bytes + 127) & ~127;
        int nstage = SMEM_RING_BUDGET / stage_bytes;
        if (nstage > MAX_STAGES) nstage = MAX_STAGES;
        if (nstage < 2) {
            fos_set_error("row too large for the shared-memory ring (stage %d bytes)", stage_bytes);
            return FOS_ERR_UNSUPPORTED;
        }
        params[1] = &stage_bytes;
        params[2] = &nstage;
        FOS_CUDA(fos_launch_ex(cfg.fn, dim3(h->n_parts), dim3(cfg.nt), static_cast<size_t>(nstage) * stage_bytes,
                               h->stream, params, h->pdl, 0));
    } else {
        const void* fn = (h->dtype == FOS_F64) ? pick_generic<double>(h->lda) : pick_generic<float>(h->lda);
        FOS_CUDA(fos_launch_ex(fn, dim3(h->n_parts), dim3(GEN_WARPS * 32), 0, h->stream, params, h->pdl, 0));
    }
    h->launches++;
    return FOS_OK;
}
