"""TEST INFRASTRUCTURE ONLY -- numpy model of the contract of ``fos_prox_grad`` (include/fos.h): the
proximal-gradient engine behind fista / fista_delta / ista, stated in terms of the C structs'
fields (scheme, alpha1, alpha2, obj_terms, delta, backtracking, eta, armijo_c, step0, max_iter, tol,
tol_ratio, adaptive_restart, restart_threshold, want_history, x0 -> x, x_hist, obj_hist, t_hist,
step_hist, ls_iters, n_iters, n_grad_calls, stop_reason).

It follows the reference loops (iterative_solvers.py:85-125, :170-243, :289-342) with the step
size handed in instead of estimated, and records the objective selected by ``obj_terms`` for every
iterate.  tests/test_host_logic_cpu.py plugs it in underneath the product's Python layer (in place
of the CUDA library) to check that layer against the golden traces on a CPU.  Only tests may
import it.
"""
from __future__ import annotations

import numpy as np

NESTEROV, DELTA, ISTA = 0, 1, 2
STOP_MAXITER, STOP_GRADNORM, STOP_STEP, STOP_RATIO = 0, 1, 2, 3


def _soft(v, thr):
    return np.sign(v) * np.maximum(np.abs(v) - thr, 0.0)


def prox_grad_model(A, b, *, scheme, alpha1, alpha2, obj_terms, delta, backtracking, eta, armijo_c, step0,
                    max_iter, tol, tol_ratio, adaptive_restart, restart_threshold, x0=None):
    A = np.asarray(A, dtype=np.float64)
    d = A.shape[1]
    x = np.zeros(d) if x0 is None else np.array(x0, dtype=np.float64)
    y = x.copy()
    x_old = x.copy()
    tau = step0
    t_mom = 1.0

    def smooth(z):
        r = A @ z - b
        v = 0.5 * r.dot(r)
        if alpha2 > 0:
            v += 0.5 * alpha2 * z.dot(z)
        return v

    def objective(z):
        r = A @ z - b
        v = 0.5 * r.dot(r)
        if obj_terms & 2:
            v += 0.5 * alpha2 * z.dot(z)
        if obj_terms & 1:
            v += alpha1 * np.abs(z).sum()
        return v

    out = {"x_hist": [x.copy()], "obj_hist": [], "t_hist": [tau], "step_hist": [], "ls_iters": [],
           "n_grad_calls": 0, "stop_reason": STOP_MAXITER}
    it = 0
    for k in range(max_iter):
        at = x if scheme == ISTA else y
        grad = A.T @ (A @ at - b)
        if alpha2 > 0:
            grad = grad + alpha2 * at
        out["n_grad_calls"] += 1
        if scheme == NESTEROV and tol > 0.0 and np.linalg.norm(grad) < tol:
            out["stop_reason"] = STOP_GRADNORM
            break
        shrinks = 0
        if backtracking:
            trial = tau
            while True:
                cand = at - trial * grad
                if alpha1 > 0:
                    cand = _soft(cand, trial * alpha1)
                if smooth(cand) <= smooth(at) + armijo_c * grad.dot(cand - at):
                    break
                trial *= eta
                shrinks += 1
            tau = trial
        x_new = at - tau * grad
        if alpha1 > 0:
            x_new = _soft(x_new, tau * alpha1)
        step_now = np.linalg.norm(x_new - x)
        step_before = np.linalg.norm(x - x_old)
        ratio = step_now / step_before if step_before > 0 else np.inf
        if scheme == NESTEROV:
            if adaptive_restart and ratio > restart_threshold:
                t_next, y_new = 1.0, x_new.copy()
            else:
                t_next = 0.5 * (1 + np.sqrt(1 + 4 * t_mom ** 2))
                y_new = x_new + ((t_mom - 1) / t_next) * (x_new - x)
        elif scheme == DELTA:
            kk = k + 1
            t_next, y_new = t_mom, x_new + (kk / (kk + 1 + delta)) * (x_new - x)
        else:
            t_next, y_new = t_mom, x_new
        it += 1
        out["x_hist"].append(x_new.copy())
        out["obj_hist"].append(objective(x_new))
        out["t_hist"].append(tau)
        out["step_hist"].append(step_now)
        out["ls_iters"].append(shrinks)
        x_old, x, y, t_mom = x, x_new, y_new, t_next
        if tol > 0.0 and step_now < tol:
            out["stop_reason"] = STOP_STEP
            break
        if scheme != ISTA and tol_ratio > 0.0 and ratio < tol_ratio:
            out["stop_reason"] = STOP_RATIO
            break
    out["x"] = x
    out["n_iters"] = it
    return out
