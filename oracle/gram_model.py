"""TEST INFRASTRUCTURE ONLY -- numpy model of the product's Gram-matrix formulation
(fastoptsolver_b200/csrc/gram_kernels.cu) of two reference routines:

* estimate_lipschitz (iterative_solvers.py:45-60) with w = (A^T A) v in place of w = A^T (A v)
  (gram_matvec_kernel + gram_power_finish_kernel; used when the Gram matrix was accumulated under
  the host->device upload), and
* fista with fixed step (iterative_solvers.py:132-245) with Grad = G y - c in place of
  A^T (A y - b) and the objective 0.5 x^T G x - c^T x + 0.5 b^T b (path_step_kernel).

The claims to pin on the CPU (tests/test_gram_model_cpu.py): same Lipschitz estimate to 1e-12 with
the same number of steps, same iterates to 1e-9, same sparsity pattern.  Only tests may import it.
"""
from __future__ import annotations

import numpy as np


def gram(A, b):
    A = np.asarray(A, dtype=np.float64)
    return A.T @ A, A.T @ b, float(b @ b)


def estimate_lipschitz_gram(G, n_iter=100, tol=1e-6):
    """Power iteration of the reference on G = A^T A; start vector from numpy's global stream,
    drawn exactly like the reference does (iterative_solvers.py:50-51).  Returns (L, steps)."""
    v = np.random.randn(G.shape[0])
    v /= np.linalg.norm(v)
    prev = 0.0
    steps = 0
    L = 0.0
    for _ in range(n_iter):
        w = G @ v
        L = np.linalg.norm(w)
        v = w / L
        steps += 1
        if abs(L - prev) < tol:
            break
        prev = L
    return L, steps


def _soft(v, thr):
    return np.sign(v) * np.maximum(np.abs(v) - thr, 0.0)


def fista_gram(G, c, btb, alpha1, alpha2, L, max_iter, t_init_factor=1.0):
    """Fixed-step FISTA on (G, c): returns (x, [objective of x_1..x_K])."""
    d = G.shape[0]
    tau = t_init_factor / L
    x = np.zeros(d)
    y = np.zeros(d)
    t_prev = 1.0
    objs = []
    for _ in range(max_iter):
        grad = G @ y - c
        if alpha2 > 0:
            grad = grad + alpha2 * y
        v = y - tau * grad
        x_new = _soft(v, tau * alpha1) if alpha1 > 0 else v
        t_cur = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t_prev * t_prev))
        beta = (t_prev - 1.0) / t_cur
        y = x_new + beta * (x_new - x)
        x, t_prev = x_new, t_cur
        obj = 0.5 * float(x @ (G @ x)) - float(c @ x) + 0.5 * btb
        if alpha2 > 0:
            obj += 0.5 * alpha2 * float(x @ x)
        if alpha1 > 0:
            obj += alpha1 * float(np.abs(x).sum())
        objs.append(obj)
    return x, objs


def fista_gram_batch(G, c, btb, alphas1, alpha2, L, max_iter, t_init_factor=1.0):
    """``fista_gram`` for a vector of L1 penalties at once (one column per penalty, the layout of
    path_step_kernel): returns (X [n_lambda, d], objectives of the final iterates).  Column j is
    elementwise the same recurrence as ``fista_gram(G, c, btb, alphas1[j], ...)``; only the matrix
    product is batched."""
    alphas1 = np.asarray(alphas1, dtype=np.float64)
    d, m = G.shape[0], alphas1.size
    tau = t_init_factor / L
    X = np.zeros((d, m))
    Y = np.zeros((d, m))
    t_prev = 1.0
    thr = tau * alphas1[None, :]
    for _ in range(max_iter):
        grad = G @ Y - c[:, None]
        if alpha2 > 0:
            grad = grad + alpha2 * Y
        V = Y - tau * grad
        X_new = np.where(alphas1[None, :] > 0, _soft(V, thr), V)
        t_cur = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t_prev * t_prev))
        beta = (t_prev - 1.0) / t_cur
        Y = X_new + beta * (X_new - X)
        X, t_prev = X_new, t_cur
    obj = 0.5 * np.einsum("ij,ij->j", X, G @ X) - c @ X + 0.5 * btb
    if alpha2 > 0:
        obj = obj + 0.5 * alpha2 * np.einsum("ij,ij->j", X, X)
    obj = obj + alphas1 * np.abs(X).sum(axis=0)
    return X.T.copy(), obj
