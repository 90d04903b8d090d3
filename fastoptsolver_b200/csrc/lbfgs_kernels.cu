// lbfgs_kernels.cu -- device-resident L-BFGS (north_star item 3): the m-vector (S, Y) history
// lives in HBM, the two-loop recursion and the line-search bookkeeping run in one cluster
// kernel after every loss+gradient pass, so an iteration costs the passes over A and nothing
// else (no per-evaluation host round trip as with the scipy driver the reference uses,
// lbfgs.py:64-70).
//
// The algorithm is the unconstrained path of L-BFGS-B 3.0, which is what
// scipy.optimize.fmin_l_bfgs_b runs for the reference: steepest descent with first step
// 1/||g|| on the first iteration, afterwards the quasi-Newton direction with H0 = (s.y/y.y) I
// (theta = y.y/s.y), memory m = 10, the More'-Thuente line search dcsrch/dcstep with
// ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmax = 1e10, at most 20 evaluations per search,
// curvature skip when s.y <= eps*(-g.d)*stp, and the stop tests ||g||_inf <= pgtol and
// (f_old - f)/max(|f_old|,|f|,1) <= factr*eps.  In exact arithmetic the iterates equal
// scipy's; in floating point the compact-representation algebra of L-BFGS-B and the two-loop
// recursion round differently, so traces agree to ~1e-10 on well-conditioned designs and
// drift on ill-conditioned ones (SURVEY.md section 4 measures the same drift CPU vs CPU).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "epilogue_common.cuh"

struct LbfgsCtrl {
    // configuration
    int m, max_iter, maxfun, maxls, obj_terms, pad;
    double alpha1, alpha2, pgtol, factr;
    // state
    int stage;  // 0: first evaluation at x0, 1: inside a line search, 2: finished
    int iter, col, head, nfg, ifun, stop, nskip;
    double f, fold, stp, dnorm, gd, gdold, theta;
    // dcsrch state (MINPACK-2)
    int brackt, ls_stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    double rho[16];
};

struct LbfgsArgs {
    EpiArgs e;        // partials, d, ldv, y (trial point = v1 of the next pass), g (trial gradient), peers
    LbfgsCtrl* L;
    double* S;        // [m][ldv]
    double* Y;        // [m][ldv]
    double* x;        // accepted iterate
    double* gacc;     // gradient at the accepted iterate
    double* dvec;     // search direction
    double* obj_hist; // full objective of every accepted iterate (lbfgs.py:56-61)
};

namespace {

constexpr int LB_MAXP = 2;  // column pairs per thread: ldv <= 2*2*2048 = 8192
constexpr double LB_EPS = 2.220446049250313e-16;

struct LbShared {
    double wred[EW][NS];
    double cl[2][FOS_EPI_CLUSTER][NS];
};

// Ordered cluster reduction of K <= NS values; slots >= first_max are combined with max, the
// others with +.  `round` alternates the DSMEM buffer so that consecutive reductions need a
// single cluster barrier each.
template <int K>
__device__ __forceinline__ void lb_reduce(double (&v)[K], LbShared& sh, int& round, int first_max = K) {
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, v[k], o);
            v[k] = (k >= first_max) ? fmax(v[k], t) : v[k] + t;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sh.wred[warp][k] = v[k];
    }
    __syncthreads();
    const int buf = round & 1;
    if (tid < K) {
        double t = sh.wred[0][tid];
        for (int w = 1; w < EW; ++w) t = (tid >= first_max) ? fmax(t, sh.wred[w][tid]) : t + sh.wred[w][tid];
        const unsigned me = cluster.block_rank();
        for (unsigned r = 0; r < cluster.num_blocks(); ++r) *cluster.map_shared_rank(&sh.cl[buf][me][tid], r) = t;
    }
    cluster.sync();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = sh.cl[buf][0][k];
        for (unsigned r = 1; r < cluster.num_blocks(); ++r) t = (k >= first_max) ? fmax(t, sh.cl[buf][r][k]) : t + sh.cl[buf][r][k];
        v[k] = t;
    }
    ++round;
}

// MINPACK-2 dcstep: safeguarded cubic/quadratic step for the More'-Thuente search
__device__ void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                       double fp, double dp, int& brackt, double stpmin, double stpmax) {
    double gamma, p, q, r, s, stpc, stpf, stpq, theta;
    const double sgnd = dp * (dx / fabs(dx));
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp;
        fy = fp;
        dy = dp;
    } else {
        if (sgnd < 0.0) {
            sty = stx;
            fy = fx;
            dy = dx;
        }
        stx = stp;
        fx = fp;
        dx = dp;
    }
    stp = stpf;
}

struct LsState {  // register copy of the dcsrch fields of LbfgsCtrl
    int brackt, stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

__device__ void dcsrch_start(LsState& s, double stp, double f, double g, double stpmin, double stpmax) {
    s.brackt = 0;
    s.stage = 1;
    s.finit = f;
    s.ginit = g;
    s.gtest = 1e-3 * g;  // ftol
    s.width = stpmax - stpmin;
    s.width1 = s.width / 0.5;
    s.stx = 0.0;
    s.fx = f;
    s.gx = g;
    s.sty = 0.0;
    s.fy = f;
    s.gy = g;
    s.stmin = 0.0;
    s.stmax = stp + 4.0 * stp;
}

// returns 0: evaluate at the new stp ("FG"), 1: converged, 2: warning (search ends, step accepted)
__device__ int dcsrch_step(LsState& s, double& stp, double f, double g, double stpmin, double stpmax) {
    const double gtol = 0.9, xtol = 0.1;   // (ftol = 1e-3 is folded into s.gtest when the search starts)
    const double ftest = s.finit + stp * s.gtest;
    if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
    int task = 0;
    if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) task = 2;
    if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) task = 2;
    if (stp == stpmax && f <= ftest && g <= s.gtest) task = 2;
    if (stp == stpmin && (f > ftest || g >= s.gtest)) task = 2;
    if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) task = 1;
    if (task != 0) return task;
    if (s.stage == 1 && f <= s.fx && f > ftest) {
        const double fm = f - stp * s.gtest;
        double fxm = s.fx - s.stx * s.gtest, fym = s.fy - s.sty * s.gtest;
        const double gm = g - s.gtest;
        double gxm = s.gx - s.gtest, gym = s.gy - s.gtest;
        dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * s.gtest;
        s.fy = fym + s.sty * s.gtest;
        s.gx = gxm + s.gtest;
        s.gy = gym + s.gtest;
    } else {
        dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
    }
    if (s.brackt) {
        if (fabs(s.sty - s.stx) >= 0.66 * s.width1) stp = s.stx + 0.5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
    }
    if (s.brackt) {
        s.stmin = fmin(s.stx, s.sty);
        s.stmax = fmax(s.stx, s.sty);
    } else {
        s.stmin = stp + 1.1 * (stp - s.stx);
        s.stmax = stp + 4.0 * (stp - s.stx);
    }
    stp = fmax(stp, stpmin);
    stp = fmin(stp, stpmax);
    if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) || (s.brackt && s.stmax - s.stmin <= xtol * s.stmax))
        stp = s.stx;
    return 0;
}

__global__ void __cluster_dims__(FOS_EPI_CLUSTER, 1, 1) __launch_bounds__(FOS_EPI_THREADS)
lbfgs_epilogue_kernel(const LbfgsArgs a) {
    __shared__ Shared sh0;   // for the pass scalars / exchange helpers
    __shared__ LbShared sh;
    cg::cluster_group cluster = cg::this_cluster();
    fos_pdl_launch_dependents();
    fos_pdl_wait();
    const EpiArgs& e = a.e;
    LbfgsCtrl* L = a.L;
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    if (L->stage == 2) return;
    const int stage = L->stage, m = L->m;
    int col = L->col, head = L->head, iter = L->iter, nfg = L->nfg, ifun = L->ifun, stop = 0, nskip = L->nskip;
    const double a1 = L->alpha1, a2 = L->alpha2;
    double f_acc = L->f, stp = L->stp, gdold = L->gdold, theta = L->theta, dnorm = L->dnorm;
    LsState ls{L->brackt, L->ls_stage, L->ginit, L->gtest, L->gx, L->gy, L->finit, L->fx,
               L->fy,     L->stx,      L->sty,   L->stmin, L->stmax, L->width, L->width1};
    double rho[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rho[i] = L->rho[i];
    const double stpmax = 1e10, stpmin = 0.0;
    int round = 0;

    double s1, s2;
    load_pass_scalars(e, sh0, s1, s2);
    const unsigned long long epoch = (e.world > 1) ? *e.peer.epoch : 0ull;
    bool comm_ok = true;
    if (e.world > 1) {
        comm_ok = peer_exchange(e, sh0, true, s1, s2, epoch);
        reduced_scalars(e, sh0, epoch, s1, s2);
    }

    // this thread's column pairs
    int cols[LB_MAXP];
    int np = 0;
    FOR_MY_COLUMN_PAIRS(c) {
        if (np < LB_MAXP) cols[np++] = c;
    }

    // ---- round 1: trial gradient and every scalar that depends on it
    double2 gt[LB_MAXP], xt[LB_MAXP];
    double r1[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) r1[k] = 0.0;
    enum { R_XX = 0, R_L1 = 1, R_GD = 2, R_YY = 3, R_GG = 4, R_GINF = 5 };
    for (int i = 0; i < np; ++i) {
        const int c = cols[i];
        double2 g = reduced_column(e, c, epoch);
        const double2 x = *reinterpret_cast<const double2*>(e.y + c);
        if (a2 != 0.0) {
            g.x = __dadd_rn(g.x, __dmul_rn(a2, x.x));
            g.y = __dadd_rn(g.y, __dmul_rn(a2, x.y));
        }
        gt[i] = g;
        xt[i] = x;
        const double2 dv = (stage == 0) ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(a.dvec + c);
        const double2 go = (stage == 0) ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(a.gacc + c);
        r1[R_XX] = fma(x.y, x.y, fma(x.x, x.x, r1[R_XX]));
        r1[R_L1] += fabs(x.x) + fabs(x.y);
        r1[R_GD] = fma(g.y, dv.y, fma(g.x, dv.x, r1[R_GD]));
        const double yx = g.x - go.x, yy = g.y - go.y;
        r1[R_YY] = fma(yy, yy, fma(yx, yx, r1[R_YY]));
        r1[R_GG] = fma(g.y, g.y, fma(g.x, g.x, r1[R_GG]));
        r1[R_GINF] = fmax(r1[R_GINF], fmax(fabs(g.x), fabs(g.y)));
    }
    lb_reduce<6>(r1, sh, round, R_GINF);
    double f = 0.5 * s1;
    if (a2 != 0.0) f = __dadd_rn(f, __dmul_rn(0.5 * a2, r1[R_XX]));
    nfg += 1;

    bool accept = false, new_search = false, restart = false;
    if (stage == 0) {
        // first evaluation: x0 is the accepted point
        f_acc = f;
        accept = true;
        if (r1[R_GINF] <= L->pgtol) stop = 1;
    } else {
        ifun += 1;
        const int task = dcsrch_step(ls, stp, f, r1[R_GD], stpmin, stpmax);
        if (task != 0) {
            accept = true;
        } else if (ifun >= L->maxls) {
            restart = true;  // too many evaluations in one search (lnsrlb: info = -3)
        }
    }

    double gd_new = r1[R_GD];
    if (accept && stage == 1) {
        iter += 1;
        // full objective of the accepted iterate, the reference's callback (lbfgs.py:56-61)
        double obj = 0.5 * s1;
        if (L->obj_terms & 2) obj = __dadd_rn(obj, __dmul_rn(0.5 * a2, r1[R_XX]));
        if (L->obj_terms & 1) obj = __dadd_rn(obj, __dmul_rn(a1, r1[R_L1]));
        if (leader && a.obj_hist) a.obj_hist[iter - 1] = obj;
        // stop tests in L-BFGS-B's order
        const double ddum = fmax(fabs(f_acc), fmax(fabs(f), 1.0));
        if (r1[R_GINF] <= L->pgtol) stop = 1;
        else if ((f_acc - f) <= LB_EPS * L->factr * ddum) stop = 2;
        if (stop == 0 && iter >= L->max_iter) stop = 3;
        if (stop == 0 && nfg > L->maxfun) stop = 4;
        // curvature pair: s = stp d, y = g_new - g_old;  s.y from the directional derivatives
        double dr, dd;
        if (stp == 1.0) {
            dr = gd_new - gdold;
            dd = -gdold;
        } else {
            dr = (gd_new - gdold) * stp;
            dd = -gdold * stp;
        }
        const double rr = r1[R_YY];
        if (dr <= LB_EPS * dd) {
            nskip += 1;
        } else {
            int slot;
            if (col < m) {
                slot = (head + col) % m;
                col += 1;
            } else {
                slot = head;
                head = (head + 1) % m;
            }
            for (int i = 0; i < np; ++i) {
                const int c = cols[i];
                const double2 xo = *reinterpret_cast<const double2*>(a.x + c);
                const double2 go = *reinterpret_cast<const double2*>(a.gacc + c);
                *reinterpret_cast<double2*>(a.S + static_cast<size_t>(slot) * e.ldv + c) =
                    make_double2(xt[i].x - xo.x, xt[i].y - xo.y);
                *reinterpret_cast<double2*>(a.Y + static_cast<size_t>(slot) * e.ldv + c) =
                    make_double2(gt[i].x - go.x, gt[i].y - go.y);
            }
            rho[slot] = 1.0 / dr;
            theta = rr / dr;
        }
        f_acc = f;
    }
    if (accept) {
        for (int i = 0; i < np; ++i) {
            const int c = cols[i];
            *reinterpret_cast<double2*>(a.x + c) = xt[i];
            *reinterpret_cast<double2*>(a.gacc + c) = gt[i];
        }
        new_search = (stop == 0);
    }
    if (restart) {
        // abandon the search: back to the accepted point with an empty memory, or give up
        if (col == 0) {
            stop = 5;
        } else {
            col = 0;
            head = 0;
            theta = 1.0;
            new_search = true;
        }
        for (int i = 0; i < np; ++i) {
            gt[i] = *reinterpret_cast<const double2*>(a.gacc + cols[i]);
            xt[i] = *reinterpret_cast<const double2*>(a.x + cols[i]);
        }
    }
    // S and Y written above are read below by other threads of the cluster
    __threadfence();
    cluster.sync();

    if (new_search) {
        // ---- two-loop recursion on this thread's columns; one cluster reduction per history dot
        double2 q[LB_MAXP];
        for (int i = 0; i < np; ++i) q[i] = gt[i];
        double alpha[16];
        for (int j = col - 1; j >= 0; --j) {
            const int slot = (head + j) % m;
            double part[1] = {0.0};
            for (int i = 0; i < np; ++i) {
                const double2 sv = *reinterpret_cast<const double2*>(a.S + static_cast<size_t>(slot) * e.ldv + cols[i]);
                part[0] = fma(sv.y, q[i].y, fma(sv.x, q[i].x, part[0]));
            }
            lb_reduce<1>(part, sh, round);
            alpha[j] = rho[slot] * part[0];
            for (int i = 0; i < np; ++i) {
                const double2 yv = *reinterpret_cast<const double2*>(a.Y + static_cast<size_t>(slot) * e.ldv + cols[i]);
                q[i].x = fma(-alpha[j], yv.x, q[i].x);
                q[i].y = fma(-alpha[j], yv.y, q[i].y);
            }
        }
        if (col > 0) {
            const double gam = 1.0 / theta;
            for (int i = 0; i < np; ++i) {
                q[i].x *= gam;
                q[i].y *= gam;
            }
        }
        for (int j = 0; j < col; ++j) {
            const int slot = (head + j) % m;
            double part[1] = {0.0};
            for (int i = 0; i < np; ++i) {
                const double2 yv = *reinterpret_cast<const double2*>(a.Y + static_cast<size_t>(slot) * e.ldv + cols[i]);
                part[0] = fma(yv.y, q[i].y, fma(yv.x, q[i].x, part[0]));
            }
            lb_reduce<1>(part, sh, round);
            const double beta = rho[slot] * part[0];
            for (int i = 0; i < np; ++i) {
                const double2 sv = *reinterpret_cast<const double2*>(a.S + static_cast<size_t>(slot) * e.ldv + cols[i]);
                q[i].x = fma(alpha[j] - beta, sv.x, q[i].x);
                q[i].y = fma(alpha[j] - beta, sv.y, q[i].y);
            }
        }
        // d = -H g; its norm and slope
        double r2[2] = {0.0, 0.0};
        for (int i = 0; i < np; ++i) {
            const double2 dv = make_double2(-q[i].x, -q[i].y);
            *reinterpret_cast<double2*>(a.dvec + cols[i]) = dv;
            r2[0] = fma(dv.y, dv.y, fma(dv.x, dv.x, r2[0]));
            r2[1] = fma(gt[i].y, dv.y, fma(gt[i].x, dv.x, r2[1]));
            q[i] = dv;
        }
        lb_reduce<2>(r2, sh, round);
        dnorm = sqrt(r2[0]);
        double gd = r2[1];
        if (gd >= 0.0) {
            // not a descent direction (lnsrlb: info = -4): drop the memory, steepest descent
            if (col == 0) {
                stop = 5;
            } else {
                col = 0;
                head = 0;
                theta = 1.0;
                double r3[1] = {0.0};
                for (int i = 0; i < np; ++i) {
                    q[i] = make_double2(-gt[i].x, -gt[i].y);
                    *reinterpret_cast<double2*>(a.dvec + cols[i]) = q[i];
                    r3[0] = fma(q[i].y, q[i].y, fma(q[i].x, q[i].x, r3[0]));
                }
                lb_reduce<1>(r3, sh, round);
                dnorm = sqrt(r3[0]);
                gd = -r3[0];
            }
        }
        if (stop == 0) {
            stp = (iter == 0) ? fmin(1.0 / dnorm, stpmax) : 1.0;
            dcsrch_start(ls, stp, f_acc, gd, stpmin, stpmax);
            gdold = gd;
            ifun = 0;
            for (int i = 0; i < np; ++i) {
                const double2 xo = *reinterpret_cast<const double2*>(a.x + cols[i]);
                *reinterpret_cast<double2*>(e.y + cols[i]) =
                    make_double2(fma(stp, q[i].x, xo.x), fma(stp, q[i].y, xo.y));
            }
        }
    } else if (stop == 0) {
        // same search, new trial step: x + stp d
        for (int i = 0; i < np; ++i) {
            const int c = cols[i];
            const double2 xo = *reinterpret_cast<const double2*>(a.x + c);
            const double2 dv = *reinterpret_cast<const double2*>(a.dvec + c);
            *reinterpret_cast<double2*>(e.y + c) = make_double2(fma(stp, dv.x, xo.x), fma(stp, dv.y, xo.y));
        }
    }

    if (leader) {
        if (!comm_ok) stop = 6;
        if (e.world > 1) *e.peer.epoch = epoch + 1;
        L->stage = (stop != 0) ? 2 : 1;
        L->iter = iter;
        L->col = col;
        L->head = head;
        L->nfg = nfg;
        L->ifun = ifun;
        L->stop = stop;
        L->nskip = nskip;
        L->f = f_acc;
        L->stp = stp;
        L->dnorm = dnorm;
        L->gdold = gdold;
        L->theta = theta;
        L->brackt = ls.brackt;
        L->ls_stage = ls.stage;
        L->ginit = ls.ginit;
        L->gtest = ls.gtest;
        L->gx = ls.gx;
        L->gy = ls.gy;
        L->finit = ls.finit;
        L->fx = ls.fx;
        L->fy = ls.fy;
        L->stx = ls.stx;
        L->sty = ls.sty;
        L->stmin = ls.stmin;
        L->stmax = ls.stmax;
        L->width = ls.width;
        L->width1 = ls.width1;
        for (int i = 0; i < 16; ++i) L->rho[i] = rho[i];
        e.ctrl->g_mode = (stop != 0) ? GM_SKIP : GM_GRAD;
        e.ctrl->n_passes += 1;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
extern "C" int fos_lbfgs(fos_design* h, const fos_lbfgs_params* p, fos_lbfgs_result* r) {
    FOS_REQUIRE(h && p && r, "null pointer argument");
    FOS_REQUIRE(p->m >= 1 && p->m <= 16, "history size m must be in 1..16");
    FOS_REQUIRE(p->max_iter >= 0 && p->maxfun >= 1 && p->maxls >= 1, "bad iteration limits");
    FOS_REQUIRE(h->ldv <= 2 * LB_MAXP * FOS_EPI_THREADS * FOS_EPI_CLUSTER, "d too large for the device L-BFGS");
    FOS_CUDA(cudaSetDevice(h->device));
    const size_t vb = static_cast<size_t>(h->ldv) * sizeof(double);
    double *S = nullptr, *Y = nullptr, *x = nullptr, *gacc = nullptr, *dvec = nullptr, *oh = nullptr;
    LbfgsCtrl* L = nullptr;
    const long long launches0 = h->launches;
    static_assert(2 * sizeof(LbfgsCtrl) <= FOS_PIN_SCRATCH, "pinned scratch too small for the L-BFGS snapshots");
    auto cleanup = [&]() {};
    auto body = [&]() -> int {
        // per-call arrays from the design's grow-only workspace (no cudaMalloc / cudaFree per fit)
        const size_t ohb = (static_cast<size_t>(std::max(p->max_iter, 1)) * sizeof(double) + 255) & ~static_cast<size_t>(255);
        const size_t lcb = (sizeof(LbfgsCtrl) + 255) & ~static_cast<size_t>(255);
        void* base = nullptr;
        FOS_TRY(fos_arena_reserve(h, vb * (2 * static_cast<size_t>(p->m) + 3) + ohb + lcb, &base));
        S = static_cast<double*>(base);
        Y = S + static_cast<size_t>(p->m) * h->ldv;
        x = Y + static_cast<size_t>(p->m) * h->ldv;
        gacc = x + h->ldv;
        dvec = gacc + h->ldv;
        oh = dvec + h->ldv;
        L = reinterpret_cast<LbfgsCtrl*>(reinterpret_cast<char*>(oh) + ohb);
        for (double* q : {S, Y}) FOS_CUDA(cudaMemsetAsync(q, 0, vb * p->m, h->stream));
        for (double* q : {x, gacc, dvec}) FOS_CUDA(cudaMemsetAsync(q, 0, vb, h->stream));
        LbfgsCtrl lc{};
        lc.m = p->m;
        lc.max_iter = p->max_iter;
        lc.maxfun = p->maxfun;
        lc.maxls = p->maxls;
        lc.obj_terms = p->obj_terms;
        lc.alpha1 = p->alpha1;
        lc.alpha2 = p->alpha2;
        lc.pgtol = p->pgtol;
        lc.factr = p->factr;
        lc.theta = 1.0;
        FOS_CUDA(cudaMemcpyAsync(L, &lc, sizeof(lc), cudaMemcpyHostToDevice, h->stream));
        FosCtrl* c = h->ctrl_host;
        memset(c, 0, sizeof(FosCtrl));
        c->g_mode = GM_GRAD;
        c->phase = PH_DONE;
        FOS_CUDA(cudaMemcpyAsync(h->ctrl, c, sizeof(FosCtrl), cudaMemcpyHostToDevice, h->stream));
        FOS_CUDA(cudaMemsetAsync(h->y, 0, vb, h->stream));  // x0 = 0 (lbfgs.py:63)
        if (p->x0) {
            memcpy(h->vec_host, p->x0, static_cast<size_t>(h->d) * sizeof(double));
            for (int q = h->d; q < h->ldv; ++q) h->vec_host[q] = 0.0;
            FOS_CUDA(cudaMemcpyAsync(h->y, h->vec_host, vb, cudaMemcpyHostToDevice, h->stream));
        }

        LbfgsArgs a{};
        a.e.ctrl = h->ctrl;
        a.e.partial_g = h->partial_g;
        a.e.partial_s = h->partial_s;
        a.e.n_parts = h->n_parts;
        a.e.d = h->d;
        a.e.ldv = h->ldv;
        a.e.y = h->y;
        a.e.xc = h->xc;
        a.e.xk = h->xk;
        a.e.g = h->g;
        a.e.world = h->world;
        a.e.rank = h->rank;
        a.e.peer = h->peer;
        a.L = L;
        a.S = S;
        a.Y = Y;
        a.x = x;
        a.gacc = gacc;
        a.dvec = dvec;
        a.obj_hist = oh;
        void* params[1] = {&a};

        FOS_CUDA(cudaEventRecord(h->ev0, h->stream));
        // launch passes in batches; poll the (small) L-BFGS control block between batches
        LbfgsCtrl* snap = static_cast<LbfgsCtrl*>(h->pin_scratch);
        memset(snap, 0, 2 * sizeof(LbfgsCtrl));
        cudaEvent_t ev[2];
        cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        const double bytes = static_cast<double>(h->n) * h->lda * (h->dtype == FOS_F64 ? 8 : 4);
        const int B = std::max(1, std::min(16, static_cast<int>(0.5e-3 / (bytes / 5.0e12 + 30e-6))));
        const long long max_pairs = static_cast<long long>(p->maxfun) + 2;
        long long launched = 0;
        int batch = 0, status = FOS_OK;
        bool finished = false;
        while (!finished && status == FOS_OK) {
            if (batch >= 2) {
                cudaEventSynchronize(ev[batch & 1]);
                if (snap[batch & 1].stage == 2) break;
            }
            if (launched >= max_pairs) break;
            for (int i = 0; i < B && launched < max_pairs && status == FOS_OK; ++i, ++launched) {
                h->grad_only_hint = true;
                status = fos_launch_grad(h, -1);
                h->grad_only_hint = false;
                if (status == FOS_OK) {
                    cudaError_t le = fos_launch_ex(reinterpret_cast<const void*>(&lbfgs_epilogue_kernel),
                                                   dim3(FOS_EPI_CLUSTER), dim3(FOS_EPI_THREADS), 0, h->stream, params,
                                                   h->pdl, 0);
                    if (le != cudaSuccess) {
                        fos_set_error("L-BFGS epilogue launch failed: %s", cudaGetErrorString(le));
                        status = FOS_ERR_CUDA;
                    }
                    h->launches++;
                }
            }
            cudaMemcpyAsync(&snap[batch & 1], L, sizeof(LbfgsCtrl), cudaMemcpyDeviceToHost, h->stream);
            cudaEventRecord(ev[batch & 1], h->stream);
            ++batch;
        }
        cudaError_t se = cudaStreamSynchronize(h->stream);
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
        if (status != FOS_OK) return status;
        if (se != cudaSuccess) {
            fos_set_error("L-BFGS loop failed: %s", cudaGetErrorString(se));
            return FOS_ERR_CUDA;
        }
        FOS_CUDA(cudaEventRecord(h->ev1, h->stream));
        FOS_CUDA(cudaMemcpyAsync(&lc, L, sizeof(lc), cudaMemcpyDeviceToHost, h->stream));
        FOS_CUDA(cudaMemcpyAsync(h->vec_host, x, vb, cudaMemcpyDeviceToHost, h->stream));
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        if (lc.stop == 6) {
            fos_set_error("multi-GPU exchange timed out: a peer rank never arrived");
            return FOS_ERR_COMM;
        }
        if (lc.stage != 2) {
            fos_set_error("device L-BFGS did not finish within %lld evaluations", max_pairs);
            return FOS_ERR_INVALID;
        }
        if (r->x) memcpy(r->x, h->vec_host, static_cast<size_t>(h->d) * sizeof(double));
        if (r->obj_hist && lc.iter > 0)
            FOS_CUDA(cudaMemcpy(r->obj_hist, oh, static_cast<size_t>(lc.iter) * sizeof(double), cudaMemcpyDeviceToHost));
        r->f_final = lc.f;
        r->n_iters = lc.iter;
        r->n_fg = lc.nfg;
        r->n_skipped = lc.nskip;
        r->stop_reason = lc.stop;
        FOS_CUDA(cudaEventElapsedTime(&r->loop_ms, h->ev0, h->ev1));
        r->kernel_launches = h->launches - launches0;
        return FOS_OK;
    };
    const int st = body();
    cleanup();
    return st;
}
