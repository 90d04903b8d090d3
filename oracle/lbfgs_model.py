"""TEST INFRASTRUCTURE ONLY -- numpy model of the device-resident L-BFGS driver
(fastoptsolver_b200/csrc/lbfgs_kernels.cu: lbfgs_epilogue_kernel), evaluation by evaluation.

The reference's L-BFGS is scipy's L-BFGS-B (lbfgs.py:64-70: m = 10, factr = 1e7, pgtol = tol,
maxfun = 15000, maxls = 20).  The device driver runs the unconstrained path of that algorithm as a
two-loop recursion with H0 = (s.y / y.y) I and the MINPACK-2 line search dcsrch / dcstep
(ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmax = 1e10).  This file restates the driver's state
machine -- the same decisions in the same order, one call of ``fg`` per pass over A -- so that the
algorithm can be pinned against scipy on the CPU (tests/test_lbfgs_model_cpu.py) independently of
any GPU.  Only tests may import it.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 2.220446049250313e-16
FTOL, GTOL, XTOL = 1e-3, 0.9, 0.1
STPMIN, STPMAX = 0.0, 1e10


class _Search:
    """dcsrch state (MINPACK-2), fields as in LbfgsCtrl."""

    def __init__(self, stp, f, g):
        self.brackt = False
        self.stage = 1
        self.finit, self.ginit = f, g
        self.gtest = FTOL * g
        self.width = STPMAX - STPMIN
        self.width1 = self.width / 0.5
        self.stx, self.fx, self.gx = 0.0, f, g
        self.sty, self.fy, self.gy = 0.0, f, g
        self.stmin = 0.0
        self.stmax = stp + 4.0 * stp


def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = max(stpmin, min(stpmax, stpf))
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * math.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


def _dcsrch_step(s: _Search, stp, f, g):
    """-> (task, stp): task 0 = evaluate at the new stp, 1 = converged, 2 = warning (search ends)."""
    ftest = s.finit + stp * s.gtest
    if s.stage == 1 and f <= ftest and g >= 0.0:
        s.stage = 2
    task = 0
    if s.brackt and (stp <= s.stmin or stp >= s.stmax):
        task = 2
    if s.brackt and s.stmax - s.stmin <= XTOL * s.stmax:
        task = 2
    if stp == STPMAX and f <= ftest and g <= s.gtest:
        task = 2
    if stp == STPMIN and (f > ftest or g >= s.gtest):
        task = 2
    if f <= ftest and abs(g) <= GTOL * (-s.ginit):
        task = 1
    if task != 0:
        return task, stp
    if s.stage == 1 and f <= s.fx and f > ftest:
        fm = f - stp * s.gtest
        fxm, fym = s.fx - s.stx * s.gtest, s.fy - s.sty * s.gtest
        gm = g - s.gtest
        gxm, gym = s.gx - s.gtest, s.gy - s.gtest
        s.stx, fxm, gxm, s.sty, fym, gym, stp, s.brackt = _dcstep(
            s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax)
        s.fx = fxm + s.stx * s.gtest
        s.fy = fym + s.sty * s.gtest
        s.gx = gxm + s.gtest
        s.gy = gym + s.gtest
    else:
        s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, s.brackt = _dcstep(
            s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax)
    if s.brackt:
        if abs(s.sty - s.stx) >= 0.66 * s.width1:
            stp = s.stx + 0.5 * (s.sty - s.stx)
        s.width1 = s.width
        s.width = abs(s.sty - s.stx)
        s.stmin = min(s.stx, s.sty)
        s.stmax = max(s.stx, s.sty)
    else:
        s.stmin = stp + 1.1 * (stp - s.stx)
        s.stmax = stp + 4.0 * (stp - s.stx)
    stp = min(max(stp, STPMIN), STPMAX)
    if (s.brackt and (stp <= s.stmin or stp >= s.stmax)) or (s.brackt and s.stmax - s.stmin <= XTOL * s.stmax):
        stp = s.stx
    return 0, stp


def lbfgs_device_model(fg, x0, m=10, max_iter=500, maxfun=15000, maxls=20, pgtol=1e-6, factr=1e7, callback=None):
    """Run the device driver's state machine with ``fg(x) -> (f, g)`` as the pass over A.

    Returns dict(x, f, n_iters, n_fg, n_skipped, stop) with stop as in fos_lbfgs: 1 projected
    gradient, 2 relative reduction, 3 max_iter, 4 maxfun, 5 abnormal line-search termination.
    ``callback(x)`` fires once per accepted iterate, like scipy's."""
    x_acc = np.array(x0, dtype=np.float64)
    d = x_acc.size
    S = np.zeros((m, d))
    Y = np.zeros((m, d))
    rho = np.zeros(m)
    col = head = it = nfg = ifun = nskip = stop = 0
    theta = 1.0
    stage = 0
    f_acc = 0.0
    g_acc = np.zeros(d)
    dvec = np.zeros(d)
    stp = gdold = 0.0
    ls = None
    y_trial = x_acc.copy()
    while True:
        f, g = fg(y_trial)                       # one pass over A
        g = np.asarray(g, dtype=np.float64)
        nfg += 1
        gd_new = float(g @ dvec) if stage == 1 else 0.0
        ginf = float(np.max(np.abs(g))) if d else 0.0
        accept = new_search = restart = False
        if stage == 0:
            f_acc = f
            accept = True
            if ginf <= pgtol:
                stop = 1
        else:
            ifun += 1
            task, stp = _dcsrch_step(ls, stp, f, gd_new)
            if task != 0:
                accept = True
            elif ifun >= maxls:
                restart = True
        if accept and stage == 1:
            it += 1
            if callback is not None:
                callback(y_trial.copy())
            ddum = max(abs(f_acc), abs(f), 1.0)
            if ginf <= pgtol:
                stop = 1
            elif (f_acc - f) <= EPS * factr * ddum:
                stop = 2
            if stop == 0 and it >= max_iter:
                stop = 3
            if stop == 0 and nfg > maxfun:
                stop = 4
            if stp == 1.0:
                dr, dd = gd_new - gdold, -gdold
            else:
                dr, dd = (gd_new - gdold) * stp, -gdold * stp
            yv = g - g_acc
            rr = float(yv @ yv)
            if dr <= EPS * dd:
                nskip += 1
            else:
                if col < m:
                    slot = (head + col) % m
                    col += 1
                else:
                    slot = head
                    head = (head + 1) % m
                S[slot] = y_trial - x_acc
                Y[slot] = yv
                rho[slot] = 1.0 / dr
                theta = rr / dr
            f_acc = f
        g_dir = g
        if accept:
            x_acc = y_trial.copy()
            g_acc = g.copy()
            new_search = (stop == 0)
        if restart:
            if col == 0:
                stop = 5
            else:
                col = head = 0
                theta = 1.0
                new_search = True
            g_dir = g_acc
        if stop == 0 and new_search:
            q = g_dir.copy()
            alpha = np.zeros(m)
            for j in range(col - 1, -1, -1):
                slot = (head + j) % m
                alpha[j] = rho[slot] * float(S[slot] @ q)
                q -= alpha[j] * Y[slot]
            if col > 0:
                q *= 1.0 / theta
            for j in range(col):
                slot = (head + j) % m
                beta = rho[slot] * float(Y[slot] @ q)
                q += (alpha[j] - beta) * S[slot]
            dvec = -q
            dnorm = math.sqrt(float(dvec @ dvec))
            gd = float(g_dir @ dvec)
            if gd >= 0.0:
                if col == 0:
                    stop = 5
                else:
                    col = head = 0
                    theta = 1.0
                    dvec = -g_dir
                    dnorm = math.sqrt(float(dvec @ dvec))
                    gd = -float(dvec @ dvec)
            if stop == 0:
                stp = min(1.0 / dnorm, STPMAX) if it == 0 else 1.0
                ls = _Search(stp, f_acc, gd)
                gdold = gd
                ifun = 0
                y_trial = x_acc + stp * dvec
        elif stop == 0:
            y_trial = x_acc + stp * dvec
        if stop != 0:
            break
        stage = 1
    return {"x": x_acc, "f": f_acc, "n_iters": it, "n_fg": nfg, "n_skipped": nskip, "stop": stop}
