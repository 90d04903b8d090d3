"""numpy restatement of the reference solvers -- TEST INFRASTRUCTURE, not product code.

Every routine keeps the reference's floating-point expression order (so that on
the same BLAS it reproduces the reference bit for bit) but is organised
differently: one accelerated-proximal-gradient engine serves both FISTA
flavours, and bookkeeping lives in one METRICS dict instead of module lists.

Reference citations are into /root/reference (read-only, absent on the GPU box).
Pinned by tests/golden/*.npz (see oracle/__init__.py).
"""
from __future__ import annotations

import time

import numpy as np

ARMIJO_C = 1e-2  # iterative_solvers.py:11 (module global ``C``)

# Filled by every solver call: per-gradient timings, per-line-search timings and
# shrink counts (iterative_solvers.py:16-18), plus the Lipschitz estimate used.
METRICS = {"grad_times": [], "ls_times": [], "ls_iters": [], "L": None}


def _reset():
    METRICS["grad_times"] = []
    METRICS["ls_times"] = []
    METRICS["ls_iters"] = []
    METRICS["L"] = None


# --------------------------------------------------------------------------- operators
def prox_l1(v, tau):
    """Soft threshold, prox_operators.py:3-8: sign(v)*max(|v|-tau, 0)."""
    mag = np.abs(v) - tau
    return np.sign(v) * np.maximum(mag, 0.0)


def prox_elastic_net(v, tau, alpha1, alpha2):
    """prox_operators.py:10-16: soft threshold at tau*alpha1, then /(1+tau*alpha2)."""
    return prox_l1(v, tau * alpha1) / (1 + tau * alpha2)


def compute_objective(x, A, b, reg_type, alpha1, alpha2):
    """objective_functions.py:3-30.  The residual is formed before reg_type is
    validated (``:13`` precedes ``:28``), so a bad reg_type still costs one pass."""
    res = A @ x - b
    smooth = 0.5 * res.dot(res)
    if reg_type in ("ridge", "elasticnet"):
        smooth += 0.5 * alpha2 * x.dot(x)
    if reg_type in ("lasso", "elasticnet"):
        nonsmooth = alpha1 * np.linalg.norm(x, 1)
    elif reg_type == "ridge":
        nonsmooth = 0.0
    else:
        raise ValueError(f"Unsupported reg_type='{reg_type}'")
    return smooth + nonsmooth


def smooth_value_and_grad(x, A, b, alpha2=0.0):
    """lbfgs.py:46-51 (``fg``): loss = 0.5 r.r (+0.5 a2 x.x), grad = A^T r (+a2 x)."""
    res = A @ x - b
    loss = 0.5 * res.dot(res)
    grad = A.T @ res
    if alpha2 != 0.0:
        loss += 0.5 * alpha2 * x.dot(x)
        grad += alpha2 * x
    return loss, grad


def estimate_lipschitz(A, n_iter=100, tol=1e-6):
    """iterative_solvers.py:45-60.  Power iteration on A^T A started from the
    *legacy global* numpy RNG (``:50``); absolute-difference stop (``:57``)."""
    v = np.random.randn(A.shape[1])
    v /= np.linalg.norm(v)
    last = 0.0
    L = None
    for _ in range(n_iter):
        w = A.T @ (A @ v)
        L = np.linalg.norm(w)
        v = w / L
        if abs(L - last) < tol:
            break
        last = L
    return L


# --------------------------------------------------------------------------- ISTA
def ista(x0, g, grad_g, prox_h, L, backtracking=False, eta=0.5, t_init_factor=1.0,
         max_iter=500, tol=0.0, return_history=False):
    """iterative_solvers.py:65-125.  Generic proximal gradient on user callables;
    ``prox_h`` receives the step size (``:98,111``)."""
    _reset()
    x = x0.copy()
    step = t_init_factor / L
    log = {"x": [x.copy()], "t": [step], "delta": []} if return_history else None
    for _ in range(max_iter):
        t0 = time.perf_counter()
        grad = grad_g(x)
        METRICS["grad_times"].append(time.perf_counter() - t0)
        if backtracking:
            shrinks = 0
            t0 = time.perf_counter()
            trial = step
            while True:
                x_new = prox_h(x - trial * grad, trial)
                move = x_new - x
                if g(x_new) <= g(x) + ARMIJO_C * grad.dot(move):
                    break
                trial *= eta
                shrinks += 1
            METRICS["ls_times"].append(time.perf_counter() - t0)
            METRICS["ls_iters"].append(shrinks)
            step = trial
        else:
            x_new = prox_h(x - step * grad, step)
        delta = np.linalg.norm(x_new - x)
        x = x_new
        if return_history:
            log["x"].append(x.copy())
            log["t"].append(step)
            log["delta"].append(delta)
        if tol > 0.0 and delta < tol:
            break
    return (x, log) if return_history else x


# --------------------------------------------------------------------------- FISTA engine
def _apg(A, b, reg_type, alpha1, alpha2, *, scheme, delta, backtracking, eta,
         t_init_factor, max_iter, tol, tol_ratio, adaptive_restart,
         restart_threshold, return_history):
    """Shared engine for ``fista`` (scheme='nesterov', iterative_solvers.py:132-245)
    and ``fista_delta`` (scheme='delta', iterative_solvers.py:251-344)."""
    _reset()
    d = A.shape[1]
    x_cur = np.zeros(d)          # always float64 (``:150``, ``:270``)
    y = x_cur.copy()
    x_old = x_cur.copy()
    t_mom = 1.0
    L = estimate_lipschitz(A)
    if alpha2 > 0:
        L += alpha2
    METRICS["L"] = L
    tau = t_init_factor / L

    nesterov = scheme == "nesterov"
    if return_history:
        hist = {"x": [x_cur.copy()] if nesterov else [], "obj": []}
    else:
        hist = None

    def smooth(z):               # ``g_smooth`` closures (``:163-168``, ``:282-287``)
        res = A @ z - b
        val = 0.5 * res.dot(res)
        if alpha2 > 0:
            val += 0.5 * alpha2 * z.dot(z)
        return val

    for it in range(max_iter):
        t0 = time.perf_counter()
        grad = A.T @ (A @ y - b)
        if alpha2 > 0:
            grad += alpha2 * y
        METRICS["grad_times"].append(time.perf_counter() - t0)

        # gradient-norm stop exists only in fista (``:179``)
        if nesterov and tol > 0.0 and np.linalg.norm(grad) < tol:
            break

        if backtracking:
            shrinks = 0
            t0 = time.perf_counter()
            trial = tau
            while True:
                cand = y - trial * grad
                if alpha1 > 0:
                    cand = prox_l1(cand, trial * alpha1)
                move = cand - y
                if smooth(cand) <= smooth(y) + ARMIJO_C * grad.dot(move):
                    break
                trial *= eta
                shrinks += 1
            METRICS["ls_times"].append(time.perf_counter() - t0)
            METRICS["ls_iters"].append(shrinks)
            tau = trial          # never grows back (``:197``, ``:312``)

        x_new = y - tau * grad
        if alpha1 > 0:
            x_new = prox_l1(x_new, tau * alpha1)

        if not nesterov and return_history:
            # fista_delta logs before the step norms, through compute_objective,
            # so reg_type matters there (``:319-322``)
            hist["x"].append(x_new.copy())
            hist["obj"].append(compute_objective(x_new, A, b, reg_type, alpha1, alpha2))

        step_now = np.linalg.norm(x_new - x_cur)
        step_before = np.linalg.norm(x_cur - x_old)
        ratio = step_now / step_before if step_before > 0 else np.inf

        if nesterov:
            if adaptive_restart and ratio > restart_threshold:
                t_next = 1.0
                y_new = x_new.copy()
            else:
                t_next = 0.5 * (1 + np.sqrt(1 + 4 * t_mom ** 2))
                beta = (t_mom - 1) / t_next
                y_new = x_new + beta * (x_new - x_cur)
        else:
            k = it + 1                       # loop runs k = 1..max_iter (``:289``)
            theta = k / (k + 1 + delta)
            t_next = t_mom
            y_new = x_new + theta * (x_new - x_cur)

        if nesterov and return_history:
            res = A @ x_new - b              # inlined objective (``:225-231``)
            obj = 0.5 * res.dot(res)
            if alpha2 > 0:
                obj += 0.5 * alpha2 * x_new.dot(x_new)
            if alpha1 > 0:
                obj += alpha1 * np.linalg.norm(x_new, 1)
            hist["obj"].append(obj)
            hist["x"].append(x_new.copy())

        x_old, x_cur, y, t_mom = x_cur, x_new, y_new, t_next

        if tol > 0.0 and step_now < tol:
            break
        if tol_ratio > 0.0 and ratio < tol_ratio:
            break

    return (x_cur, hist) if return_history else x_cur


def fista(A, b, reg_type, alpha1, alpha2, backtracking=False, eta=0.5, t_init_factor=1.0,
          max_iter=500, tol=0.0, tol_ratio=0.0, adaptive_restart=False,
          restart_threshold=1.0, return_history=False):
    """iterative_solvers.py:132-245.  ``reg_type`` is accepted and never read."""
    return _apg(A, b, reg_type, alpha1, alpha2, scheme="nesterov", delta=None,
                backtracking=backtracking, eta=eta, t_init_factor=t_init_factor,
                max_iter=max_iter, tol=tol, tol_ratio=tol_ratio,
                adaptive_restart=adaptive_restart, restart_threshold=restart_threshold,
                return_history=return_history)


def fista_delta(A, b, reg_type, alpha1, alpha2, delta, backtracking=False, eta=0.5,
                t_init_factor=1.0, max_iter=500, tol=0.0, tol_ratio=0.0,
                return_history=False):
    """iterative_solvers.py:251-344.  theta_k = k/(k+1+delta), delta > 2 asserted."""
    _reset()
    assert delta > 2, "In FISTA-Δ, delta must be > 2 for convergence (course requirement)"
    return _apg(A, b, reg_type, alpha1, alpha2, scheme="delta", delta=delta,
                backtracking=backtracking, eta=eta, t_init_factor=t_init_factor,
                max_iter=max_iter, tol=tol, tol_ratio=tol_ratio,
                adaptive_restart=False, restart_threshold=1.0,
                return_history=return_history)


# --------------------------------------------------------------------------- L-BFGS
class LBFGSSolver:
    """lbfgs.py:7-73.  Thin wrapper over scipy.optimize.fmin_l_bfgs_b (third-party,
    unpinned in the reference; scipy 1.18.1 in this image).  The L1 term never
    enters ``fg``; ``history_`` records the full objective and is only cleared in
    ``__init__`` (``:39``), so repeated fits accumulate."""

    def __init__(self, reg_type, alpha1, alpha2, max_iter=500, tol=1e-6, eps=1e-8):
        kind, a1, a2 = reg_type, alpha1, alpha2
        if kind == "lasso":
            a2 = 0.0
        elif kind == "ridge":
            a1 = 0.0
        elif kind == "elasticnet":
            if alpha1 < eps:
                kind, a1 = "ridge", 0.0
            elif alpha2 < eps:
                kind, a2 = "lasso", 0.0
        else:
            raise ValueError(f"Unsupported reg_type='{reg_type}'")
        self.reg_type, self.alpha1, self.alpha2 = kind, a1, a2
        self.max_iter = max_iter
        self.tol = tol
        self.history_ = []

    def fit(self, A, b):
        from scipy.optimize import fmin_l_bfgs_b

        _reset()
        ridge_like = self.reg_type in ("ridge", "elasticnet")

        def fg(x):
            t0 = time.perf_counter()
            res = A @ x - b
            loss = 0.5 * res.dot(res)
            grad = A.T @ res
            if ridge_like:
                loss += 0.5 * self.alpha2 * x.dot(x)
                grad += self.alpha2 * x
            METRICS["grad_times"].append(time.perf_counter() - t0)
            return loss, grad

        def on_iter(xk):
            self.history_.append(
                compute_objective(xk, A, b, self.reg_type, self.alpha1, self.alpha2))

        out = fmin_l_bfgs_b(func=fg, x0=np.zeros(A.shape[1]), maxiter=self.max_iter,
                            pgtol=self.tol, callback=on_iter)
        self.x_ = out[0]
        self.final_obj_ = out[1]
        return self
