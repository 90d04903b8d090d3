// pg_logic.cuh -- the proximal-gradient state machine (fista / fista_delta / ista of the reference,
// iterative_solvers.py:65-344) as device functions shared by its two hosts:
//   * epilogue_kernel (epilogue_kernels.cu): one cluster of 8 CTAs launched after every pass;
//   * solve_stream_kernel (grad_kernels.cu): the persistent kernel that runs a whole solve, where
//     the same logic is the tail of every pass, column-sliced over all CTAs of the grid.
// Both hosts feed it the reduced gradient per column pair and the NS reduced scalars; everything
// here is elementwise or scalar, written with explicit round-to-nearest intrinsics so that, given
// the same gradient, the iterates reproduce numpy's two-rounding arithmetic bit for bit.
#pragma once

#include "fos_common.cuh"

namespace {

enum { S_GG = 0, S_DX2 = 1, S_L1 = 2, S_XX = 3, S_GD = 4, S_YY = 5, S_R1 = 6, S_R2 = 7 };

// loads of d-vectors: CG = true reads through L2 (ld.global.cg) -- required when another CTA of the
// SAME launch wrote the data (the L1 of this SM may hold the previous pass's line)
template <bool CG>
__device__ __forceinline__ double2 pg_ld2(const double* p) {
    if (CG) return __ldcg(reinterpret_cast<const double2*>(p));
    return *reinterpret_cast<const double2*>(p);
}

__device__ __forceinline__ double pg_prox_point(double y, double t, double g, double a1) {
    double v = __dsub_rn(y, __dmul_rn(t, g));
    if (a1 > 0.0) v = fos_soft_threshold(v, __dmul_rn(t, a1));
    return v;
}

// snapshot of the control block taken at the start of a pass's epilogue
struct PgIn {
    int phase, scheme, backtracking, adaptive_restart, k, obj_pending, want_obj, obj_terms, shrinks, n_grad_calls;
    int max_iter, stop_reason, use_qrec;
    double beta_y;
    double a1, a2, tau, trial_t, gy, gd, cand_xx, pend_l2, pend_l1, t_mom, prev_step;
    double eta, armijo_c, tol, tol_ratio, restart_thr, delta;
    unsigned long long pass_t0;
};

template <bool CG>
__device__ __forceinline__ PgIn pg_read(const FosCtrl* C) {
    PgIn in;
    if (CG) {
        const volatile FosCtrl* V = C;
        in.phase = V->phase; in.scheme = V->scheme; in.backtracking = V->backtracking;
        in.adaptive_restart = V->adaptive_restart; in.k = V->k; in.obj_pending = V->obj_pending;
        in.want_obj = V->want_obj; in.obj_terms = V->obj_terms; in.shrinks = V->shrinks;
        in.n_grad_calls = V->n_grad_calls; in.max_iter = V->max_iter; in.stop_reason = V->stop_reason;
        in.a1 = V->alpha1; in.a2 = V->alpha2; in.tau = V->tau; in.trial_t = V->trial_t; in.gy = V->gy;
        in.gd = V->gd; in.cand_xx = V->cand_xx; in.pend_l2 = V->pend_l2; in.pend_l1 = V->pend_l1;
        in.t_mom = V->t_mom; in.prev_step = V->prev_step; in.eta = V->eta; in.armijo_c = V->armijo_c;
        in.tol = V->tol; in.tol_ratio = V->tol_ratio; in.restart_thr = V->restart_thr; in.delta = V->delta;
        in.pass_t0 = V->pass_t0;
        in.use_qrec = V->use_qrec; in.beta_y = V->beta_y;
    } else {
        in.phase = C->phase; in.scheme = C->scheme; in.backtracking = C->backtracking;
        in.adaptive_restart = C->adaptive_restart; in.k = C->k; in.obj_pending = C->obj_pending;
        in.want_obj = C->want_obj; in.obj_terms = C->obj_terms; in.shrinks = C->shrinks;
        in.n_grad_calls = C->n_grad_calls; in.max_iter = C->max_iter; in.stop_reason = C->stop_reason;
        in.a1 = C->alpha1; in.a2 = C->alpha2; in.tau = C->tau; in.trial_t = C->trial_t; in.gy = C->gy;
        in.gd = C->gd; in.cand_xx = C->cand_xx; in.pend_l2 = C->pend_l2; in.pend_l1 = C->pend_l1;
        in.t_mom = C->t_mom; in.prev_step = C->prev_step; in.eta = C->eta; in.armijo_c = C->armijo_c;
        in.tol = C->tol; in.tol_ratio = C->tol_ratio; in.restart_thr = C->restart_thr; in.delta = C->delta;
        in.pass_t0 = C->pass_t0;
        in.use_qrec = C->use_qrec; in.beta_y = C->beta_y;
    }
    return in;
}

// Armijo test of the candidate evaluated by the pass that just ran (:191 / :306 / :101); s2 = its
// squared residual norm
__device__ __forceinline__ void pg_armijo(const PgIn& in, double s2, bool& accept, double& t_new) {
    accept = false;
    t_new = in.trial_t;
    if (in.phase == PH_TRIAL) {
        double lhs = 0.5 * s2;
        if (in.a2 > 0.0) lhs = __dadd_rn(lhs, __dmul_rn(0.5 * in.a2, in.cand_xx));
        const double rhs = __dadd_rn(in.gy, __dmul_rn(in.armijo_c, in.gd));
        accept = lhs <= rhs;
        if (!accept) t_new = __dmul_rn(in.trial_t, in.eta);
    }
}

// ---- elementwise 1 (PH_GRAD): gradient (+a2 y), candidate point, local sums; g = reduced A^T r
template <bool CG>
__device__ __forceinline__ void pg_elem1_grad(const EpiArgs& e, const PgIn& in, int c, double2 g, double (&sums)[FOS_NSCAL],
                                              double2* cand_out = nullptr, double2* xk_out = nullptr) {
    const double2 y = pg_ld2<CG>(e.y + c);
    const double2 xk = pg_ld2<CG>(e.xk + c);
    if (in.a2 > 0.0) {
        g.x = __dadd_rn(g.x, __dmul_rn(in.a2, y.x));
        g.y = __dadd_rn(g.y, __dmul_rn(in.a2, y.y));
    }
    *reinterpret_cast<double2*>(e.g + c) = g;
    double2 cand;
    cand.x = (c < e.d) ? pg_prox_point(y.x, in.tau, g.x, in.a1) : 0.0;
    cand.y = (c + 1 < e.d) ? pg_prox_point(y.y, in.tau, g.y, in.a1) : 0.0;
    *reinterpret_cast<double2*>(e.xc + c) = cand;
    const double dx = cand.x - xk.x, dy = cand.y - xk.y;
    sums[S_GG] = fma(g.y, g.y, fma(g.x, g.x, sums[S_GG]));
    sums[S_DX2] = fma(dy, dy, fma(dx, dx, sums[S_DX2]));
    sums[S_L1] += fabs(cand.x) + fabs(cand.y);
    sums[S_XX] = fma(cand.y, cand.y, fma(cand.x, cand.x, sums[S_XX]));
    sums[S_GD] = fma(g.y, cand.y - y.y, fma(g.x, cand.x - y.x, sums[S_GD]));
    sums[S_YY] = fma(y.y, y.y, fma(y.x, y.x, sums[S_YY]));
    if (cand_out) *cand_out = cand;
    if (xk_out) *xk_out = xk;
}

// ---- elementwise 1 (PH_TRIAL): keep the accepted candidate or form the next one with the shrunk step
template <bool CG>
__device__ __forceinline__ void pg_elem1_trial(const EpiArgs& e, const PgIn& in, int c, bool accept, double t_new,
                                               double (&sums)[FOS_NSCAL], double2* cand_out = nullptr,
                                               double2* xk_out = nullptr) {
    const double2 xk = pg_ld2<CG>(e.xk + c);
    double2 cand;
    if (accept) {
        cand = pg_ld2<CG>(e.xc + c);
    } else {
        const double2 g = pg_ld2<CG>(e.g + c);
        const double2 y = pg_ld2<CG>(e.y + c);
        cand.x = (c < e.d) ? pg_prox_point(y.x, t_new, g.x, in.a1) : 0.0;
        cand.y = (c + 1 < e.d) ? pg_prox_point(y.y, t_new, g.y, in.a1) : 0.0;
        *reinterpret_cast<double2*>(e.xc + c) = cand;
        sums[S_GD] = fma(g.y, cand.y - y.y, fma(g.x, cand.x - y.x, sums[S_GD]));
    }
    const double dx = cand.x - xk.x, dy = cand.y - xk.y;
    sums[S_DX2] = fma(dy, dy, fma(dx, dx, sums[S_DX2]));
    sums[S_L1] += fabs(cand.x) + fabs(cand.y);
    sums[S_XX] = fma(cand.y, cand.y, fma(cand.x, cand.x, sums[S_XX]));
    if (cand_out) *cand_out = cand;
    if (xk_out) *xk_out = xk;
}

// everything the scalar logic decides (identical in every thread that runs it)
struct PgOut {
    int n_phase, n_gmode, n_k, n_shrinks, n_obj_pending, n_stop, n_ngrad;
    double n_tau, n_trial, n_gy, n_gd, n_cxx, n_pl2, n_pl1, n_tmom, n_prev, n_beta;
    bool do_update, obj_known, resolve_obj, ls_done, plain_copy, write_new_obj;
    double resolved_obj, beta, this_step, new_obj;
};

// sums: the NS column sums reduced over all columns; s1 / s2: squared residual norms of the pass
__device__ __forceinline__ PgOut pg_decide(const PgIn& in, const double (&sums)[FOS_NSCAL], double s1, double s2,
                                           bool accept, double t_new) {
    PgOut o;
    o.n_phase = in.phase; o.n_gmode = GM_SKIP; o.n_k = in.k; o.n_shrinks = in.shrinks;
    o.n_obj_pending = in.obj_pending; o.n_stop = in.stop_reason; o.n_ngrad = in.n_grad_calls;
    o.n_tau = in.tau; o.n_trial = in.trial_t; o.n_gy = in.gy; o.n_gd = in.gd; o.n_cxx = in.cand_xx;
    o.n_pl2 = in.pend_l2; o.n_pl1 = in.pend_l1; o.n_tmom = in.t_mom; o.n_prev = in.prev_step; o.n_beta = in.beta_y;
    o.do_update = false; o.obj_known = false; o.resolve_obj = false; o.ls_done = false;
    o.plain_copy = false; o.write_new_obj = false;
    o.resolved_obj = 0.0; o.beta = 0.0; o.this_step = 0.0; o.new_obj = 0.0;

    if (in.phase == PH_GRAD) {
        o.n_ngrad = in.n_grad_calls + 1;
        if (in.obj_pending) {
            o.resolved_obj = __dadd_rn(__dadd_rn(0.5 * s2, in.pend_l2), in.pend_l1);
            o.resolve_obj = true;
            o.n_obj_pending = 0;
        }
        if (in.scheme == FOS_SCHEME_NESTEROV && in.tol > 0.0 && sqrt(sums[S_GG]) < in.tol) {
            o.n_stop = FOS_STOP_GRADNORM;
            o.n_phase = PH_DONE;
            o.n_gmode = GM_SKIP;
        } else if (in.backtracking) {
            o.n_trial = in.tau;
            o.n_shrinks = 0;
            o.n_gy = 0.5 * s1;
            if (in.a2 > 0.0) o.n_gy = __dadd_rn(o.n_gy, __dmul_rn(0.5 * in.a2, sums[S_YY]));
            o.n_gd = sums[S_GD];
            o.n_cxx = sums[S_XX];
            o.n_phase = PH_TRIAL;
            o.n_gmode = GM_DOT2;
        } else {
            o.do_update = true;
        }
    } else if (in.phase == PH_TRIAL) {
        if (accept) {
            o.n_tau = in.trial_t;
            o.ls_done = true;
            o.do_update = true;
            o.obj_known = true;
        } else {
            o.n_trial = t_new;
            o.n_shrinks = in.shrinks + 1;
            o.n_gd = sums[S_GD];
            o.n_cxx = sums[S_XX];
            o.n_gmode = GM_DOT2;
        }
    } else {  // PH_FINALOBJ
        o.resolved_obj = __dadd_rn(__dadd_rn(0.5 * s2, in.pend_l2), in.pend_l1);
        o.resolve_obj = true;
        o.n_obj_pending = 0;
        o.n_phase = PH_DONE;
        o.n_gmode = GM_SKIP;
    }

    if (o.do_update) {
        o.this_step = sqrt(sums[S_DX2]);
        const double ratio = (in.prev_step > 0.0) ? o.this_step / in.prev_step : INFINITY;
        if (in.scheme == FOS_SCHEME_NESTEROV) {
            if (in.adaptive_restart && ratio > in.restart_thr) {
                o.n_tmom = 1.0;
                o.plain_copy = true;
            } else {
                o.n_tmom = 0.5 * (1.0 + sqrt(1.0 + 4.0 * (in.t_mom * in.t_mom)));
                o.beta = (in.t_mom - 1.0) / o.n_tmom;
            }
        } else if (in.scheme == FOS_SCHEME_DELTA) {
            const double kk = static_cast<double>(in.k + 1);
            o.beta = kk / ((kk + 1.0) + in.delta);
        } else {
            o.plain_copy = true;
        }
        if (in.want_obj) {
            const double l2t = (in.obj_terms & 2) ? __dmul_rn(0.5 * in.a2, sums[S_XX]) : 0.0;
            const double l1t = (in.obj_terms & 1) ? __dmul_rn(in.a1, sums[S_L1]) : 0.0;
            if (o.obj_known) {
                o.new_obj = __dadd_rn(__dadd_rn(0.5 * s2, l2t), l1t);
                o.write_new_obj = true;
            } else {
                o.n_pl2 = l2t;
                o.n_pl1 = l1t;
                o.n_obj_pending = 1;
            }
        }
        o.n_k = in.k + 1;
        o.n_prev = o.this_step;
        bool stop = false;
        if (in.tol > 0.0 && o.this_step < in.tol) {
            stop = true;
            o.n_stop = FOS_STOP_STEP;
        } else if (in.scheme != FOS_SCHEME_ISTA && in.tol_ratio > 0.0 && ratio < in.tol_ratio) {
            stop = true;
            o.n_stop = FOS_STOP_RATIO;
        } else if (o.n_k >= in.max_iter) {
            stop = true;
            o.n_stop = FOS_STOP_MAXITER;
        }
        if (stop) {
            o.n_phase = o.n_obj_pending ? PH_FINALOBJ : PH_DONE;
            o.n_gmode = o.n_obj_pending ? GM_DOT2 : GM_SKIP;
        } else {
            o.n_phase = PH_GRAD;
            // the next pass evaluates the gradient at the new y and, for the history, the residual norm of the
            // iterate just accepted: from the recurrence (every pass keeps q current) or from a second dot
            o.n_gmode = GM_GRAD | (in.use_qrec ? GM_QREC : (o.n_obj_pending ? GM_DOT2 : 0));
        }
        o.n_beta = o.plain_copy ? 0.0 : o.beta;   // what y_{k+1} is formed with (elementwise 2)
    }
    return o;
}

// ---- elementwise 2: momentum point, roll the iterate, history row (only when o.do_update)
// (the candidate and the current iterate of the column pair: still in the caller's registers, or re-read)
__device__ __forceinline__ void pg_elem2_regs(const EpiArgs& e, const PgIn& in, const PgOut& o, int c, double2 cand, double2 xk);
template <bool CG>
__device__ __forceinline__ void pg_elem2(const EpiArgs& e, const PgIn& in, const PgOut& o, int c) {
    pg_elem2_regs(e, in, o, c, pg_ld2<CG>(e.xc + c), pg_ld2<CG>(e.xk + c));
}
__device__ __forceinline__ void pg_elem2_regs(const EpiArgs& e, const PgIn& in, const PgOut& o, int c, double2 cand, double2 xk) {
    double* hrow = (e.hist.x_hist != nullptr) ? e.hist.x_hist + static_cast<size_t>(in.k + 1) * e.d : nullptr;
    double2 yn;
    if (o.plain_copy) {
        yn = cand;
    } else {
        yn.x = __dadd_rn(cand.x, __dmul_rn(o.beta, __dsub_rn(cand.x, xk.x)));
        yn.y = __dadd_rn(cand.y, __dmul_rn(o.beta, __dsub_rn(cand.y, xk.y)));
    }
    *reinterpret_cast<double2*>(e.y + c) = yn;
    *reinterpret_cast<double2*>(e.xk + c) = cand;
    if (hrow != nullptr) {
        if (c < e.d) hrow[c] = cand.x;
        if (c + 1 < e.d) hrow[c + 1] = cand.y;
    }
}

// ---- the dynamic state the next pass starts from (the persistent kernel applies it to every CTA's
// shared-memory copy of the control block as well: the decision is identical everywhere)
__device__ __forceinline__ void pg_apply_state(FosCtrl* C, PgOut o, bool comm_ok) {
    if (!comm_ok) {  // a peer never arrived: abort the solve, the host reports FOS_ERR_COMM
        o.n_phase = PH_DONE;
        o.n_gmode = GM_SKIP;
        o.n_stop = -1;
    }
    C->phase = o.n_phase;
    C->g_mode = o.n_gmode;
    C->k = o.n_k;
    C->shrinks = o.n_shrinks;
    C->obj_pending = o.n_obj_pending;
    C->stop_reason = o.n_stop;
    C->n_grad_calls = o.n_ngrad;
    C->tau = o.n_tau;
    C->trial_t = o.n_trial;
    C->gy = o.n_gy;
    C->gd = o.n_gd;
    C->cand_xx = o.n_cxx;
    C->pend_l2 = o.n_pl2;
    C->pend_l1 = o.n_pl1;
    C->t_mom = o.n_tmom;
    C->prev_step = o.n_prev;
    C->beta_y = o.n_beta;
}

// ---- the one thread that owns the control block writes the new state and the history scalars
// TOTALS_BY_CALLER: the caller keeps running totals of epi_ns / xchg_ns / n_passes and stores them itself
// (the persistent kernel: no read-modify-write round trips to HBM while 147 CTAs wait at a barrier)
template <bool TOTALS_BY_CALLER = false>
__device__ __forceinline__ void pg_commit(FosCtrl* C, const FosHist& hist, const PgIn& in, PgOut o, bool comm_ok,
                                          unsigned long long t_epi0, unsigned long long xchg_ns) {
    const float dt_ms = static_cast<float>(static_cast<double>(fos_globaltimer() - in.pass_t0) * 1e-6);
    if (in.phase == PH_GRAD && hist.grad_ms) hist.grad_ms[in.n_grad_calls] = dt_ms;
    if (in.phase == PH_TRIAL && hist.ls_ms) hist.ls_ms[in.k] += dt_ms;
    if (o.resolve_obj && hist.obj_hist && in.k >= 1) hist.obj_hist[in.k - 1] = o.resolved_obj;
    if (o.write_new_obj && hist.obj_hist) hist.obj_hist[in.k] = o.new_obj;
    if (o.ls_done && hist.ls_iters) hist.ls_iters[in.k] = in.shrinks;
    if (o.do_update) {
        if (hist.t_hist) hist.t_hist[in.k + 1] = o.n_tau;
        if (hist.step_hist) hist.step_hist[in.k] = o.this_step;
    }
    if (!TOTALS_BY_CALLER) {
        C->epi_ns += fos_globaltimer() - t_epi0;
        C->xchg_ns += xchg_ns;
    }
    if (!TOTALS_BY_CALLER) C->n_passes += 1;
    pg_apply_state(C, o, comm_ok);
}

}  // namespace
