"""TEST INFRASTRUCTURE ONLY -- numpy model of the residual recurrence the streaming kernels use to record the
objective of fixed-step runs without a second dot product (csrc/grad_kernels.cu: GM_QREC):

    y_k = x_k + beta_k (x_k - x_{k-1})   =>   q_k := A x_k - b = (r_k + beta_k q_{k-1}) / (1 + beta_k),   r_k = A y_k - b

so the objective of iterate x_k (iterative_solvers.py:225-231 / :321: compute_objective after the update) needs no
product with x_k: the gradient pass at y_k already has r_k.  ``fista_with_recurrence`` runs the reference's fixed-step
loop (iterative_solvers.py:170-243 and :289-342, no backtracking) this way -- one product A y per iteration, plus one
trailing product for the last iterate, exactly the kernel's pass count -- and returns the same (x, history) as the
reference so that tests can compare objective traces with the golden traces.  Only tests may import it.
"""
from __future__ import annotations

import numpy as np


def _soft(v, thr):
    return np.sign(v) * np.maximum(np.abs(v) - thr, 0.0)


def fista_with_recurrence(A, b, alpha1, alpha2, L, max_iter, *, scheme="nesterov", delta=None, obj_terms=3,
                          adaptive_restart=False, restart_threshold=1.0):
    A = np.asarray(A, dtype=np.float64)
    d = A.shape[1]
    x = np.zeros(d)
    y = x.copy()
    t = 1.0 / L
    t_mom = 1.0
    prev_step = 0.0
    beta_y = 0.0                      # what the current y was formed with
    q = None                          # A x_k - b of the current iterate
    pend = None                       # (l2 term, l1 term) of the iterate whose residual norm is still pending
    xs, objs = [x.copy()], []
    for k in range(max_iter):
        r = A @ y - b                 # the pass at y_k
        q = r if beta_y == 0.0 else (r + beta_y * q) / (1.0 + beta_y)
        if pend is not None:          # objective of x_k, recorded one pass late like the kernel does
            objs.append(0.5 * q.dot(q) + pend[0] + pend[1])
        grad = A.T @ r
        if alpha2 > 0:
            grad = grad + alpha2 * y
        v = y - t * grad
        x_new = _soft(v, t * alpha1) if alpha1 > 0 else v
        step = np.linalg.norm(x_new - x)
        ratio = step / prev_step if prev_step > 0 else np.inf
        if scheme == "nesterov":
            if adaptive_restart and ratio > restart_threshold:
                t_mom, beta = 1.0, 0.0
            else:
                t_next = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t_mom * t_mom))
                beta = (t_mom - 1.0) / t_next
                t_mom = t_next
        else:
            kk = float(k + 1)
            beta = kk / (kk + 1.0 + delta)
        y = x_new + beta * (x_new - x) if beta != 0.0 else x_new.copy()
        beta_y = beta
        pend = (0.5 * alpha2 * x_new.dot(x_new) if obj_terms & 2 else 0.0,
                alpha1 * np.abs(x_new).sum() if obj_terms & 1 else 0.0)
        x, prev_step = x_new, step
        xs.append(x.copy())
    if pend is not None:              # trailing pass: residual of the last iterate directly
        r = A @ x - b
        objs.append(0.5 * r.dot(r) + pend[0] + pend[1])
    return x, {"x": xs, "obj": objs}
