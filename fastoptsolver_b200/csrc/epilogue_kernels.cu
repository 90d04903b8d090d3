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
    const int phase = C->phase;
    if (phase == PH_DONE) return;
    const int scheme = C->scheme, backtracking = C->backtracking, k = C->k;
    const double a1 = C->alpha1, a2 = C->alpha2;
    const double tau = C->tau, trial_t = C->trial_t;
    const int obj_pending = C->obj_pending, want_obj = C->want_obj, obj_terms = C->obj_terms;
    const int shrinks = C->shrinks, n_grad_calls = C->n_grad_calls;
    const double gy_saved = C->gy, gd_saved = C->gd, cand_xx_saved = C->cand_xx;
    const double pend_l2 = C->pend_l2, pend_l1 = C->pend_l1;
    const double t_mom = C->t_mom, prev_step = C->prev_step;
    const unsigned long long pass_t0 = C->pass_t0;
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

    // Armijo test of the candidate evaluated by the pass that just ran (:191 / :306 / :101)
    bool accept = false;
    double t_new = trial_t;
    if (phase == PH_TRIAL) {
        double lhs = 0.5 * s2;
        if (a2 > 0.0) lhs = __dadd_rn(lhs, __dmul_rn(0.5 * a2, cand_xx_saved));
        const double rhs = __dadd_rn(gy_saved, __dmul_rn(C->armijo_c, gd_saved));
        accept = lhs <= rhs;
        if (!accept) t_new = __dmul_rn(trial_t, C->eta);
    }

    // ---- elementwise 1: gradient, candidate point, local sums
    enum { S_GG = 0, S_DX2 = 1, S_L1 = 2, S_XX = 3, S_GD = 4, S_YY = 5 };
    if (phase == PH_GRAD) {
        FOR_MY_COLUMN_PAIRS(c) {
            double2 g = reduced_column(e, c, epoch);
            const double2 y = *reinterpret_cast<const double2*>(e.y + c);
            const double2 xk = *reinterpret_cast<const double2*>(e.xk + c);
            if (a2 > 0.0) {
                g.x = __dadd_rn(g.x, __dmul_rn(a2, y.x));
                g.y = __dadd_rn(g.y, __dmul_rn(a2, y.y));
            }
            *reinterpret_cast<double2*>(e.g + c) = g;
            double2 cand;
            cand.x = (c < e.d) ? prox_point(y.x, tau, g.x, a1) : 0.0;
            cand.y = (c + 1 < e.d) ? prox_point(y.y, tau, g.y, a1) : 0.0;
            *reinterpret_cast<double2*>(e.xc + c) = cand;
            const double dx = cand.x - xk.x, dy = cand.y - xk.y;
            sums[S_GG] = fma(g.y, g.y, fma(g.x, g.x, sums[S_GG]));
            sums[S_DX2] = fma(dy, dy, fma(dx, dx, sums[S_DX2]));
            sums[S_L1] += fabs(cand.x) + fabs(cand.y);
            sums[S_XX] = fma(cand.y, cand.y, fma(cand.x, cand.x, sums[S_XX]));
            sums[S_GD] = fma(g.y, cand.y - y.y, fma(g.x, cand.x - y.x, sums[S_GD]));
            sums[S_YY] = fma(y.y, y.y, fma(y.x, y.x, sums[S_YY]));
        }
    } else if (phase == PH_TRIAL) {
        FOR_MY_COLUMN_PAIRS(c) {
            const double2 xk = *reinterpret_cast<const double2*>(e.xk + c);
            double2 cand;
            if (accept) {
                cand = *reinterpret_cast<const double2*>(e.xc + c);
            } else {
                const double2 g = *reinterpret_cast<const double2*>(e.g + c);
                const double2 y = *reinterpret_cast<const double2*>(e.y + c);
                cand.x = (c < e.d) ? prox_point(y.x, t_new, g.x, a1) : 0.0;
                cand.y = (c + 1 < e.d) ? prox_point(y.y, t_new, g.y, a1) : 0.0;
                *reinterpret_cast<double2*>(e.xc + c) = cand;
                sums[S_GD] = fma(g.y, cand.y - y.y, fma(g.x, cand.x - y.x, sums[S_GD]));
            }
            const double dx = cand.x - xk.x, dy = cand.y - xk.y;
            sums[S_DX2] = fma(dy, dy, fma(dx, dx, sums[S_DX2]));
            sums[S_L1] += fabs(cand.x) + fabs(cand.y);
            sums[S_XX] = fma(cand.y, cand.y, fma(cand.x, cand.x, sums[S_XX]));
        }
    }
    cluster_sum(sums, sh);

    // ---- scalar logic (identical in every thread)
    int n_phase = phase, n_gmode = GM_SKIP, n_k = k, n_shrinks = shrinks;
    int n_obj_pending = obj_pending, n_stop = C->stop_reason, n_ngrad = n_grad_calls;
    double n_tau = tau, n_trial = trial_t, n_gy = gy_saved, n_gd = gd_saved, n_cxx = cand_xx_saved;
    double n_pl2 = pend_l2, n_pl1 = pend_l1, n_tmom = t_mom, n_prev = prev_step;
    bool do_update = false, obj_known = false;
    double resolved_obj = 0.0;
    bool resolve_obj = false;
    bool ls_done = false;

    if (phase == PH_GRAD) {
        n_ngrad = n_grad_calls + 1;
        if (obj_pending) {
            resolved_obj = __dadd_rn(__dadd_rn(0.5 * s2, pend_l2), pend_l1);
            resolve_obj = true;
            n_obj_pending = 0;
        }
        if (scheme == FOS_SCHEME_NESTEROV && C->tol > 0.0 && sqrt(sums[S_GG]) < C->tol) {
            n_stop = FOS_STOP_GRADNORM;
            n_phase = PH_DONE;
            n_gmode = GM_SKIP;
        } else if (backtracking) {
            n_trial = tau;
            n_shrinks = 0;
            n_gy = 0.5 * s1;
            if (a2 > 0.0) n_gy = __dadd_rn(n_gy, __dmul_rn(0.5 * a2, sums[S_YY]));
            n_gd = sums[S_GD];
            n_cxx = sums[S_XX];
            n_phase = PH_TRIAL;
            n_gmode = GM_DOT2;
        } else {
            do_update = true;
        }
    } else if (phase == PH_TRIAL) {
        if (accept) {
            n_tau = trial_t;
            ls_done = true;
            do_update = true;
            obj_known = true;
        } else {
            n_trial = t_new;
            n_shrinks = shrinks + 1;
            n_gd = sums[S_GD];
            n_cxx = sums[S_XX];
            n_gmode = GM_DOT2;
        }
    } else {  // PH_FINALOBJ
        resolved_obj = __dadd_rn(__dadd_rn(0.5 * s2, pend_l2), pend_l1);
        resolve_obj = true;
        n_obj_pending = 0;
        n_phase = PH_DONE;
        n_gmode = GM_SKIP;
    }

    double beta = 0.0, this_step = 0.0;
    bool plain_copy = false;
    double new_obj = 0.0;
    bool write_new_obj = false;
    if (do_update) {
        this_step = sqrt(sums[S_DX2]);
        const double ratio = (prev_step > 0.0) ? this_step / prev_step : INFINITY;
        if (scheme == FOS_SCHEME_NESTEROV) {
            if (C->adaptive_restart && ratio > C->restart_thr) {
                n_tmom = 1.0;
                plain_copy = true;
            } else {
                n_tmom = 0.5 * (1.0 + sqrt(1.0 + 4.0 * (t_mom * t_mom)));
                beta = (t_mom - 1.0) / n_tmom;
            }
        } else if (scheme == FOS_SCHEME_DELTA) {
            const double kk = static_cast<double>(k + 1);
            beta = kk / ((kk + 1.0) + C->delta);
        } else {
            plain_copy = true;
        }
        if (want_obj) {
            const double l2t = (obj_terms & 2) ? __dmul_rn(0.5 * a2, sums[S_XX]) : 0.0;
            const double l1t = (obj_terms & 1) ? __dmul_rn(a1, sums[S_L1]) : 0.0;
            if (obj_known) {
                new_obj = __dadd_rn(__dadd_rn(0.5 * s2, l2t), l1t);
                write_new_obj = true;
            } else {
                n_pl2 = l2t;
                n_pl1 = l1t;
                n_obj_pending = 1;
            }
        }
        n_k = k + 1;
        n_prev = this_step;
        bool stop = false;
        if (C->tol > 0.0 && this_step < C->tol) {
            stop = true;
            n_stop = FOS_STOP_STEP;
        } else if (scheme != FOS_SCHEME_ISTA && C->tol_ratio > 0.0 && ratio < C->tol_ratio) {
            stop = true;
            n_stop = FOS_STOP_RATIO;
        } else if (n_k >= C->max_iter) {
            stop = true;
            n_stop = FOS_STOP_MAXITER;
        }
        if (stop) {
            n_phase = n_obj_pending ? PH_FINALOBJ : PH_DONE;
            n_gmode = n_obj_pending ? GM_DOT2 : GM_SKIP;
        } else {
            n_phase = PH_GRAD;
            n_gmode = GM_GRAD | (n_obj_pending ? GM_DOT2 : 0);
        }

        // ---- elementwise 2: momentum point, roll the iterate, history row
        double* hrow = (e.hist.x_hist != nullptr) ? e.hist.x_hist + static_cast<size_t>(k + 1) * e.d : nullptr;
        FOR_MY_COLUMN_PAIRS(c) {
            const double2 cand = *reinterpret_cast<const double2*>(e.xc + c);
            const double2 xk = *reinterpret_cast<const double2*>(e.xk + c);
            double2 yn;
            if (plain_copy) {
                yn = cand;
            } else {
                yn.x = __dadd_rn(cand.x, __dmul_rn(beta, __dsub_rn(cand.x, xk.x)));
                yn.y = __dadd_rn(cand.y, __dmul_rn(beta, __dsub_rn(cand.y, xk.y)));
            }
            *reinterpret_cast<double2*>(e.y + c) = yn;
            *reinterpret_cast<double2*>(e.xk + c) = cand;
            if (hrow != nullptr) {
                if (c < e.d) hrow[c] = cand.x;
                if (c + 1 < e.d) hrow[c + 1] = cand.y;
            }
        }
    }

    if (leader) {
        const float dt_ms = static_cast<float>(static_cast<double>(fos_globaltimer() - pass_t0) * 1e-6);
        if (phase == PH_GRAD && e.hist.grad_ms) e.hist.grad_ms[n_grad_calls] = dt_ms;
        if (phase == PH_TRIAL && e.hist.ls_ms) e.hist.ls_ms[k] += dt_ms;
        if (resolve_obj && e.hist.obj_hist && k >= 1) e.hist.obj_hist[k - 1] = resolved_obj;
        if (write_new_obj && e.hist.obj_hist) e.hist.obj_hist[k] = new_obj;
        if (ls_done && e.hist.ls_iters) e.hist.ls_iters[k] = shrinks;
        if (do_update) {
            if (e.hist.t_hist) e.hist.t_hist[k + 1] = n_tau;
            if (e.hist.step_hist) e.hist.step_hist[k] = this_step;
        }
        C->epi_ns += fos_globaltimer() - t_epi0;
        C->xchg_ns += t_x1 - t_x0;
        if (e.world > 1) *e.peer.epoch = epoch + 1;
        if (!comm_ok) {  // a peer never arrived: abort the solve, the host reports FOS_ERR_COMM
            n_phase = PH_DONE;
            n_gmode = GM_SKIP;
            n_stop = -1;
        }
        C->phase = n_phase;
        C->g_mode = n_gmode;
        C->k = n_k;
        C->shrinks = n_shrinks;
        C->obj_pending = n_obj_pending;
        C->stop_reason = n_stop;
        C->n_grad_calls = n_ngrad;
        C->n_passes += 1;
        C->tau = n_tau;
        C->trial_t = n_trial;
        C->gy = n_gy;
        C->gd = n_gd;
        C->cand_xx = n_cxx;
        C->pend_l2 = n_pl2;
        C->pend_l1 = n_pl1;
        C->t_mom = n_tmom;
        C->prev_step = n_prev;
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
