// epilogue_kernels.cu -- everything that happens on d-length vectors after a pass over A:
// ordered cross-CTA reduction of the A^T r partials, (+alpha2 y), soft threshold
// (prox_operators.py:3-8), Armijo test (iterative_solvers.py:187-194), step norms and
// ratio (:204-206), Nesterov / delta momentum with adaptive restart (:209-221, :330-331),
// history and objective bookkeeping (:224-232, :319-322), stop rules (:179, :238, :242),
// and the power-iteration normalisation (:55-59).  All solver scalars live in FosCtrl on
// the device; the host never takes part in a decision.
//
// One launch = one thread-block cluster of 8 CTAs: the d columns are split across the
// cluster, scalar sums are combined through distributed shared memory in fixed rank order,
// so every CTA takes the same branch with the same numbers.
//
// Elementwise formulas use explicit round-to-nearest intrinsics (no FMA contraction) so that,
// given the same gradient, they reproduce numpy's two-rounding arithmetic bit for bit.
#include "epilogue_common.cuh"
#include "pg_logic.cuh"

namespace {

__global__ void __cluster_dims__(FOS_EPI_CLUSTER, 1, 1) __launch_bounds__(FOS_EPI_THREADS)
epilogue_kernel(const EpiArgs e) {
    __shared__ Shared sh;
    cg::cluster_group cluster = cg::this_cluster();
    // let the next gradient kernel start filling its ring, then wait for this pass's kernel
    fos_pdl_launch_dependents();
    fos_pdl_wait();
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    FosCtrl* C = e.ctrl;

    double sums[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) sums[k] = 0.0;

    // ------------------------------------------------------------------ one-shot ops
    if (e.op == EOP_POWER) {
        if (C->g_mode == GM_SKIP) return;
        const unsigned long long epoch = (e.world > 1) ? *e.peer.epoch : 0ull;
        bool comm_ok = true;
        if (e.world > 1) comm_ok = peer_exchange(e, sh, true, 0.0, 0.0, epoch);
        FOR_MY_COLUMN_PAIRS(c) {
            const double2 w = reduced_column(e, c, epoch);
            *reinterpret_cast<double2*>(e.g + c) = w;
            if (c < e.d) sums[0] = fma(w.x, w.x, sums[0]);
            if (c + 1 < e.d) sums[0] = fma(w.y, w.y, sums[0]);
        }
        const double L_prev = C->L_prev, ptol = C->ptol;
        const int pit = C->pit, pit_max = C->pit_max;
        cluster_sum(sums, sh);
        const double L = sqrt(sums[0]);
        FOR_MY_COLUMN_PAIRS(c) {
            const double2 w = *reinterpret_cast<const double2*>(e.g + c);
            *reinterpret_cast<double2*>(e.y + c) = make_double2(__ddiv_rn(w.x, L), __ddiv_rn(w.y, L));
        }
        if (leader) {
            const bool done = (fabs(L - L_prev) < ptol) || (pit + 1 >= pit_max);
            C->L = L;
            C->L_prev = L;
            C->pit = pit + 1;
            C->g_mode = (done || !comm_ok) ? GM_SKIP : (GM_GRAD | GM_NOB);
            C->n_passes += 1;
            if (!comm_ok) C->stop_reason = -1;
            if (e.world > 1) *e.peer.epoch = epoch + 1;
        }
        return;
    }
    if (e.op == EOP_FG || e.op == EOP_OBJ) {
        double s1, s2;
        load_pass_scalars(e, sh, s1, s2);
        const unsigned long long epoch = (e.world > 1) ? *e.peer.epoch : 0ull;
        bool comm_ok = true;
        if (e.world > 1) {
            comm_ok = peer_exchange(e, sh, e.op == EOP_FG, s1, s2, epoch);
            reduced_scalars(e, sh, epoch, s1, s2);
        }
        if (e.op == EOP_FG) {
            FOR_MY_COLUMN_PAIRS(c) {
                double2 g = reduced_column(e, c, epoch);
                const double2 x = *reinterpret_cast<const double2*>(e.y + c);
                if (e.op_bits & 2) {
                    g.x = __dadd_rn(g.x, __dmul_rn(e.op_a2, x.x));
                    g.y = __dadd_rn(g.y, __dmul_rn(e.op_a2, x.y));
                }
                *reinterpret_cast<double2*>(e.g + c) = g;
                sums[0] = fma(x.y, x.y, fma(x.x, x.x, sums[0]));  // padding columns hold 0
            }
        } else {
            FOR_MY_COLUMN_PAIRS(c) {
                const double2 x = *reinterpret_cast<const double2*>(e.xc + c);
                sums[0] = fma(x.y, x.y, fma(x.x, x.x, sums[0]));
                sums[1] += fabs(x.x) + fabs(x.y);
            }
        }
        cluster_sum(sums, sh);
        if (leader) {
            double val = 0.5 * ((e.op == EOP_FG) ? s1 : s2);
            if (e.op_bits & 2) val = __dadd_rn(val, __dmul_rn(0.5 * e.op_a2, sums[0]));
            if (e.op == EOP_OBJ && (e.op_bits & 1)) val = __dadd_rn(val, __dmul_rn(e.op_a1, sums[1]));
            C->out[0] = val;
            C->out[1] = s1;
            C->out[2] = s2;
            C->n_passes += 1;
            C->stop_reason = comm_ok ? 0 : -1;
            if (e.world > 1) *e.peer.epoch = epoch + 1;
        }
        return;
    }

    // ------------------------------------------------------------------ proximal-gradient engine
    // (state machine and elementwise formulas: pg_logic.cuh, shared with the persistent solve kernel)
    if (C->phase == PH_DONE) return;
    const PgIn in = pg_read<false>(C);
    const int phase = in.phase;
    const unsigned long long t_epi0 = fos_globaltimer();

    double s1, s2;
    load_pass_scalars(e, sh, s1, s2);
    const unsigned long long epoch = (e.world > 1) ? *e.peer.epoch : 0ull;
    bool comm_ok = true;
    unsigned long long t_x0 = 0, t_x1 = 0;
    if (e.world > 1) {
        if (leader) t_x0 = fos_globaltimer();
        comm_ok = peer_exchange(e, sh, phase == PH_GRAD, s1, s2, epoch);
        if (leader) t_x1 = fos_globaltimer();
        reduced_scalars(e, sh, epoch, s1, s2);
    }

    bool accept;
    double t_new;
    pg_armijo(in, s2, accept, t_new);

    // ---- elementwise 1: gradient, candidate point, local sums
    if (phase == PH_GRAD) {
        FOR_MY_COLUMN_PAIRS(c) pg_elem1_grad<false>(e, in, c, reduced_column(e, c, epoch), sums);
    } else if (phase == PH_TRIAL) {
        FOR_MY_COLUMN_PAIRS(c) pg_elem1_trial<false>(e, in, c, accept, t_new, sums);
    }
    cluster_sum(sums, sh);

    // ---- scalar logic (identical in every thread)
    const PgOut o = pg_decide(in, sums, s1, s2, accept, t_new);

    // ---- elementwise 2: momentum point, roll the iterate, history row
    if (o.do_update) {
        FOR_MY_COLUMN_PAIRS(c) pg_elem2<false>(e, in, o, c);
    }

    if (leader) {
        if (e.world > 1) *e.peer.epoch = epoch + 1;
        pg_commit(C, e.hist, in, o, comm_ok, t_epi0, t_x1 - t_x0);
    }
}

// standalone soft threshold (prox_operators.py:3-16) on a flat buffer
__global__ void prox_kernel(const double* __restrict__ v, double* __restrict__ out, long long len,
                            double thresh, double scale, int do_scale) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        double s = fos_soft_threshold(v[i], thresh);
        out[i] = do_scale ? __ddiv_rn(s, scale) : s;
    }
}

}  // namespace

int fos_launch_epilogue(fos_design* h, int op, int g_mode_ran, const FosHist& hist, double a1,
                        double a2, int bits) {
    EpiArgs e{};
    e.ctrl = h->ctrl;
    e.hist = hist;
    e.partial_g = h->partial_g;
    e.partial_s = h->partial_s;
    e.n_parts = h->n_parts;
    e.d = h->d;
    e.ldv = h->ldv;
    e.op = op;
    e.g_mode_ran = g_mode_ran;
    e.y = h->y;
    e.xc = h->xc;
    e.xk = h->xk;
    e.g = h->g;
    e.op_a1 = a1;
    e.op_a2 = a2;
    e.op_bits = bits;
    e.world = h->world;
    e.rank = h->rank;
    e.peer = h->peer;
    void* params[1] = {&e};
    FOS_CUDA(fos_launch_ex(reinterpret_cast<const void*>(&epilogue_kernel), dim3(FOS_EPI_CLUSTER),
                           dim3(FOS_EPI_THREADS), 0, h->stream, params, h->pdl, 0));
    h->launches++;
    return FOS_OK;
}

int fos_launch_prox(const double* v_dev, double* out_dev, long long len, double thresh, double scale,
                    cudaStream_t stream) {
    if (len <= 0) return FOS_OK;
    long long blocks = (len + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    prox_kernel<<<dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream>>>(v_dev, out_dev, len, thresh, scale,
                                                                            scale != 1.0 ? 1 : 0);
    FOS_CUDA(cudaGetLastError());
    return FOS_OK;
}
