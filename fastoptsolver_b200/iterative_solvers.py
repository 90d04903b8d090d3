"""ISTA / FISTA / FISTA-delta on the GPU with the reference's module API
(iterative_solvers.py of ElBaldo1/FastOptSolver): same names, signatures, defaults,
history layout, module-global metric lists and Armijo constant ``C``.

Every solver call turns into ONE call of ``fos_prox_grad`` (include/fos.h): the loop, the
Armijo decisions, restarts and stop rules all run on the device.
"""
from __future__ import annotations

import ctypes as C_
import time
from typing import Callable

import numpy as np

from . import _lib
from .design import as_design, cache_enabled, find_by_matrix
from .operators import L1Prox, SmoothGrad, SmoothValue, prox_l1  # noqa: F401

# ---------------------------------------------------------------------
# module globals of the reference (iterative_solvers.py:11, :16-18); ``C`` is read at call
# time, the lists are cleared in place so that ``lbfgs.grad_call_times is grad_call_times``
# ---------------------------------------------------------------------
C: float = 1e-2
grad_call_times = []
ls_call_times = []
ls_call_iters = []

# device-side facts about the last call that the reference API has no slot for
last_run = {}


def reset_metrics() -> None:
    grad_call_times.clear()
    ls_call_times.clear()
    ls_call_iters.clear()


def get_metrics():
    """Same keys as the reference (iterative_solvers.py:32-40); times are device times of
    the passes (read from %globaltimer inside the kernels), in seconds."""
    return {
        'grad_num_calls': len(grad_call_times),
        'grad_time_total': sum(grad_call_times),
        'grad_time_mean': np.mean(grad_call_times) if grad_call_times else 0.0,
        'ls_num_calls': len(ls_call_times),
        'ls_time_total': sum(ls_call_times),
        'ls_time_mean': np.mean(ls_call_times) if ls_call_times else 0.0,
        'ls_iters_total': sum(ls_call_iters),
    }


def estimate_lipschitz(A, n_iter: int = 100, tol: float = 1e-6) -> float:
    """Power iteration for lambda_max(A^T A) (iterative_solvers.py:45-60).  The start vector
    is drawn on the host from numpy's legacy global RNG so the stream advances exactly as in
    the reference (``:50``); the <= n_iter passes over A run on the device."""
    des = find_by_matrix(A)
    if des is None:
        des = as_design(A, _b_placeholder(A))
    v = np.random.randn(des.shape[1])
    v /= np.linalg.norm(v)
    L, iters, ms = des.power_iter(v, n_iter, tol)
    # "gram": the products ran on G = A^T A accumulated under the upload (include/fos.h,
    # fos_design_upload_gram); "stream": two fused passes over A per step
    st, world = des.upload_gram()["state"], des.comm_info()[1]
    on_gram = (st == 2) if world > 1 else (st == 1)
    last_run["lipschitz"] = {"iters": iters, "gpu_ms": ms, "via": "gram" if on_gram else "stream"}
    return np.float64(L)


def _b_placeholder(A):
    # estimate_lipschitz takes A only; when A is a bare host array we need some b to build
    # the design.  A zero vector is fine (the power iteration never reads b).
    if hasattr(A, "handle"):
        return None
    return _zeros_for(A)


_ZERO_B = {}


def _zeros_for(A):
    n = A.shape[0]
    z = _ZERO_B.get(n)
    if z is None:
        _ZERO_B.clear()
        z = _ZERO_B[n] = np.zeros(n)
    return z


def _run(des, *, scheme, alpha1, alpha2, obj_terms, delta, backtracking, eta, step0, max_iter, tol,
         tol_ratio, adaptive_restart, restart_threshold, want_history, x0=None):
    lib = _lib.load()
    d = des.shape[1]
    K = int(max_iter)
    p = _lib.PGParams()
    p.scheme = scheme
    p.alpha1, p.alpha2 = float(alpha1), float(alpha2)
    p.obj_terms = obj_terms
    p.delta = float(delta)
    p.backtracking = int(bool(backtracking))
    p.eta = float(eta)
    p.armijo_c = float(C)
    p.step0 = float(step0)
    p.max_iter = K
    p.tol, p.tol_ratio = float(tol), float(tol_ratio)
    p.adaptive_restart = int(bool(adaptive_restart))
    p.restart_threshold = float(restart_threshold)
    p.want_history = int(bool(want_history))
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        p.x0 = x0.ctypes.data_as(_lib.c_double_p)

    x = np.empty(d)
    xh = np.empty((K + 1, d)) if want_history else None
    oh = np.empty(max(K, 1))
    th = np.empty(K + 1)
    sh = np.empty(max(K, 1))
    li = np.zeros(max(K, 1), dtype=np.int32)
    gm = np.zeros(K + 1, dtype=np.float32)
    lm = np.zeros(max(K, 1), dtype=np.float32)
    r = _lib.PGResult()
    r.x = x.ctypes.data_as(_lib.c_double_p)
    if xh is not None:
        r.x_hist = xh.ctypes.data_as(_lib.c_double_p)
    r.obj_hist = oh.ctypes.data_as(_lib.c_double_p)
    r.t_hist = th.ctypes.data_as(_lib.c_double_p)
    r.step_hist = sh.ctypes.data_as(_lib.c_double_p)
    r.ls_iters = li.ctypes.data_as(_lib.c_int_p)
    r.grad_ms = gm.ctypes.data_as(_lib.c_float_p)
    r.ls_ms = lm.ctypes.data_as(_lib.c_float_p)
    _lib.check(lib.fos_prox_grad(des.handle, C_.byref(p), C_.byref(r)))

    it = r.n_iters
    grad_call_times.extend((gm[: r.n_grad_calls].astype(np.float64) * 1e-3).tolist())
    if backtracking:
        ls_call_times.extend((lm[:it].astype(np.float64) * 1e-3).tolist())
        ls_call_iters.extend(int(v) for v in li[:it])
    last_run["solver"] = {
        "iters": it, "grad_calls": r.n_grad_calls, "passes": r.n_passes, "stop_reason": r.stop_reason,
        "loop_ms": r.loop_ms, "kernel_launches": r.kernel_launches,
        "grad_kernel_ms": r.grad_kernel_ms, "grad_kernel_launches": r.grad_kernel_launches,
        "epilogue_ms": r.epilogue_ms, "exchange_ms": r.exchange_ms,
        "host_ms": {"setup": r.host_setup_ms, "loop": r.host_loop_ms, "finish": r.host_finish_ms},
    }
    return x, it, xh, oh, th, sh


# ---------------------------------------------------------------------
# ISTA
# ---------------------------------------------------------------------
def ista(
    x0: np.ndarray,
    g: Callable[[np.ndarray], float],
    grad_g: Callable[[np.ndarray], np.ndarray],
    prox_h: Callable[[np.ndarray, float], np.ndarray],
    L: float,
    backtracking: bool = False,
    eta: float = 0.5,
    t_init_factor: float = 1.0,
    max_iter: int = 500,
    tol: float = 0.0,
    return_history: bool = False,
):
    """Proximal gradient on (g, grad_g, prox_h) (iterative_solvers.py:65-125).

    With the operator objects of ``operators.ista_callables`` the loop runs on the device.
    Arbitrary Python callables keep the reference's loop semantics on the host: the loop
    then only sequences the user's own functions (nothing of this package's arithmetic is
    replaced by numpy there)."""
    reset_metrics()
    fused = (isinstance(g, SmoothValue) and isinstance(grad_g, SmoothGrad) and isinstance(prox_h, L1Prox)
             and g.design is grad_g.design is prox_h.design
             and g.alpha2 == grad_g.alpha2 and grad_g.alpha1 == prox_h.alpha1)
    if not fused:
        return _ista_callbacks(x0, g, grad_g, prox_h, L, backtracking, eta, t_init_factor, max_iter, tol,
                               return_history)
    x0 = np.asarray(x0, dtype=np.float64)
    step0 = t_init_factor / L
    terms = (1 if prox_h.alpha1 > 0 else 0) | (2 if grad_g.alpha2 > 0 else 0)
    x, it, xh, oh, th, sh = _run(
        g.design, scheme=_lib.SCHEME_ISTA, alpha1=prox_h.alpha1, alpha2=grad_g.alpha2, obj_terms=terms, delta=0.0,
        backtracking=backtracking, eta=eta, step0=step0, max_iter=max_iter, tol=tol, tol_ratio=0.0,
        adaptive_restart=False, restart_threshold=1.0, want_history=return_history, x0=x0)
    if not return_history:
        return x
    # extra (not part of the reference's log): objective of x_1..x_k, free by-product of the pass
    last_run["ista_obj"] = [np.float64(v) for v in oh[:it]]
    log = {"x": [xh[i].copy() for i in range(it + 1)],
           "t": [step0] + [float(v) for v in th[1: it + 1]],
           "delta": [np.float64(v) for v in sh[:it]]}
    return x, log


def _ista_callbacks(x0, g, grad_g, prox_h, L, backtracking, eta, t_init_factor, max_iter, tol, return_history):
    import time
    x = x0.copy()
    t = t_init_factor / L
    log = {"x": [x.copy()], "t": [t], "delta": []} if return_history else None
    for _ in range(max_iter):
        t0 = time.perf_counter()
        grad = grad_g(x)
        grad_call_times.append(time.perf_counter() - t0)
        if backtracking:
            shrinks, t0, t_k = 0, time.perf_counter(), t
            while True:
                x_new = prox_h(x - t_k * grad, t_k)
                if g(x_new) <= g(x) + C * grad.dot(x_new - x):
                    break
                t_k *= eta
                shrinks += 1
            ls_call_times.append(time.perf_counter() - t0)
            ls_call_iters.append(shrinks)
            t = t_k
        else:
            x_new = prox_h(x - t * grad, t)
        delta = np.linalg.norm(x_new - x)
        x = x_new
        if return_history:
            log["x"].append(x.copy())
            log["t"].append(t)
            log["delta"].append(delta)
        if tol > 0.0 and delta < tol:
            break
    return (x, log) if return_history else x


# ---------------------------------------------------------------------
# FISTA
# ---------------------------------------------------------------------
def fista(
    A,
    b,
    reg_type: str,
    alpha1: float,
    alpha2: float,
    backtracking: bool = False,
    eta: float = 0.5,
    t_init_factor: float = 1.0,
    max_iter: int = 500,
    tol: float = 0.0,
    tol_ratio: float = 0.0,
    adaptive_restart: bool = False,
    restart_threshold: float = 1.0,
    return_history: bool = False,
):
    """Accelerated proximal gradient (iterative_solvers.py:132-245).  ``reg_type`` is
    accepted and ignored exactly as in the reference; alpha1 > 0 / alpha2 > 0 decide."""
    reset_metrics()
    t0 = time.perf_counter()
    des = as_design(A, b)
    t1 = time.perf_counter()
    if des is not A:      # uploaded (or found) by this call: what the upload did (include/fos.h, fos_design_upload_gram)
        ug = des.upload_gram()
        last_run["upload_gram"] = {"state": ug["state"], "copy_ms": ug["copy_ms"], "tail_ms": ug["tail_ms"]}
    L_val = estimate_lipschitz(des)
    t2 = time.perf_counter()
    if alpha2 > 0:
        L_val += alpha2
    last_run["L"] = float(L_val)
    terms = (1 if alpha1 > 0 else 0) | (2 if alpha2 > 0 else 0)
    x, it, xh, oh, _, _ = _run(
        des, scheme=_lib.SCHEME_NESTEROV, alpha1=alpha1, alpha2=alpha2, obj_terms=terms, delta=0.0,
        backtracking=backtracking, eta=eta, step0=t_init_factor / L_val, max_iter=max_iter, tol=tol,
        tol_ratio=tol_ratio, adaptive_restart=adaptive_restart, restart_threshold=restart_threshold,
        want_history=return_history)
    t3 = time.perf_counter()
    if des is not A and not cache_enabled():
        des.close()       # uploaded by this call and owned by nobody else: release the device copy now, not at GC time
    t4 = time.perf_counter()
    # host wall clock of the stages of this call (seconds): design lookup / upload, Lipschitz
    # estimate, solver loop incl. result download, release of the device copy
    last_run["host_s"] = {"design": t1 - t0, "lipschitz": t2 - t1, "solve": t3 - t2, "teardown": t4 - t3}
    if not return_history:
        return x
    history = {"x": [xh[i].copy() for i in range(it + 1)], "obj": [np.float64(v) for v in oh[:it]]}
    return x, history


# ---------------------------------------------------------------------
# FISTA-delta
# ---------------------------------------------------------------------
def fista_delta(
    A,
    b,
    reg_type: str,
    alpha1: float,
    alpha2: float,
    delta: float,
    backtracking: bool = False,
    eta: float = 0.5,
    t_init_factor: float = 1.0,
    max_iter: int = 500,
    tol: float = 0.0,
    tol_ratio: float = 0.0,
    return_history: bool = False,
):
    """FISTA with theta_k = k/(k+1+delta) (iterative_solvers.py:251-344).  The recorded
    objective goes through ``compute_objective(reg_type)`` in the reference (``:321``), so
    ``reg_type`` selects its terms here and an unknown one raises ValueError -- only when
    history is requested, like the reference."""
    reset_metrics()
    assert delta > 2, "In FISTA-Δ, delta must be > 2 for convergence (course requirement)"
    des = as_design(A, b)
    L_val = estimate_lipschitz(des)
    if alpha2 > 0:
        L_val += alpha2
    last_run["L"] = float(L_val)
    terms = {"lasso": 1, "ridge": 2, "elasticnet": 3}.get(reg_type)
    if terms is None:
        if return_history and max_iter > 0:
            raise ValueError(f"Unsupported reg_type='{reg_type}'")
        terms = 0
    x, it, xh, oh, _, _ = _run(
        des, scheme=_lib.SCHEME_DELTA, alpha1=alpha1, alpha2=alpha2, obj_terms=terms, delta=delta,
        backtracking=backtracking, eta=eta, step0=t_init_factor / L_val, max_iter=max_iter, tol=tol,
        tol_ratio=tol_ratio, adaptive_restart=False, restart_threshold=1.0, want_history=return_history)
    if not return_history:
        return x
    history = {"x": [xh[i].copy() for i in range(1, it + 1)], "obj": [np.float64(v) for v in oh[:it]]}
    return x, history
