"""LBFGSSolver with the reference's API (lbfgs.py:7-73).

``fit`` keeps scipy's L-BFGS-B as the host driver exactly like the reference
(lbfgs.py:64-70) -- scipy.optimize.fmin_l_bfgs_b is the reference's own third-party
dependency for this path -- while every evaluation of ``fg`` (loss + gradient, the two
dgemv calls of lbfgs.py:46-48) is ONE fused pass of the CUDA gradient kernel, and every
callback objective (lbfgs.py:57) is one dot-only pass.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
from scipy.optimize import fmin_l_bfgs_b

from . import _lib

from .design import as_design
from .iterative_solvers import grad_call_times, last_run, reset_metrics  # noqa: F401  (shared lists)

_BITS = {"lasso": 1, "ridge": 2, "elasticnet": 3}
DRIVER = "scipy"


class LBFGSSolver:
    """L-BFGS for Ridge and smooth Elastic-Net, with the reference's tiny-alpha shortcut."""

    def __init__(self, reg_type, alpha1, alpha2, max_iter=500, tol=1e-6, eps=1e-8, *, driver=None):
        # driver: "scipy" (default; the reference's own driver, every fg one fused GPU pass) or
        # "device" (two-loop recursion + More'-Thuente search on the GPU, no host round trips);
        # the module global DRIVER / env FOS_LBFGS_DRIVER set the default.
        self.driver = driver
        if reg_type == "lasso":
            kind, a1, a2 = "lasso", alpha1, 0.0
        elif reg_type == "ridge":
            kind, a1, a2 = "ridge", 0.0, alpha2
        elif reg_type == "elasticnet":
            if alpha1 < eps:
                kind, a1, a2 = "ridge", 0.0, alpha2
            elif alpha2 < eps:
                kind, a1, a2 = "lasso", alpha1, 0.0
            else:
                kind, a1, a2 = "elasticnet", alpha1, alpha2
        else:
            raise ValueError(f"Unsupported reg_type='{reg_type}'")
        self.reg_type, self.alpha1, self.alpha2 = kind, a1, a2
        self.max_iter = max_iter
        self.tol = tol
        self.history_ = []      # only reset here: repeated fits accumulate (lbfgs.py:39)

    def fit(self, A, b=None):
        reset_metrics()
        des = as_design(A, b)
        ridge_like = self.reg_type in ("ridge", "elasticnet")
        a2 = float(self.alpha2) if ridge_like else 0.0
        bits = _BITS[self.reg_type]
        driver = self.driver or os.environ.get("FOS_LBFGS_DRIVER") or DRIVER
        if driver == "device":
            return self._fit_device(des, a2, bits)
        if driver != "scipy":
            raise ValueError(f"unknown L-BFGS driver {driver!r}")
        gpu_ms = []

        def fg(x):
            import time
            t0 = time.perf_counter()
            loss, grad = des.grad(x, a2)
            grad_call_times.append(time.perf_counter() - t0)
            return loss, grad

        def callback(xk):
            self.history_.append(np.float64(des.objective(xk, bits, self.alpha1, self.alpha2)))

        res = fmin_l_bfgs_b(func=fg, x0=np.zeros(des.shape[1]), maxiter=self.max_iter, pgtol=self.tol,
                            callback=callback)
        self.x_ = res[0]
        self.final_obj_ = res[1]
        last_run["lbfgs"] = {"fg_calls": len(grad_call_times), "iters": res[2].get("nit"), "gpu_ms": gpu_ms}
        return self

    def _fit_device(self, des, a2, bits):
        """scipy defaults of fmin_l_bfgs_b (m=10, factr=1e7, maxfun=15000, maxls=20) on the device."""
        d = des.shape[1]
        K = int(self.max_iter)
        p = _lib.LbfgsParams(m=10, max_iter=K, maxfun=15000, maxls=20, obj_terms=bits, alpha1=float(self.alpha1),
                             alpha2=a2, pgtol=float(self.tol), factr=1e7)
        x = np.empty(d)
        oh = np.empty(max(K, 1))
        r = _lib.LbfgsResult()
        r.x = x.ctypes.data_as(_lib.c_double_p)
        r.obj_hist = oh.ctypes.data_as(_lib.c_double_p)
        _lib.check(_lib.load().fos_lbfgs(des.handle, C.byref(p), C.byref(r)))
        self.history_.extend(np.float64(v) for v in oh[: r.n_iters])
        self.x_ = x
        self.final_obj_ = r.f_final
        per = r.loop_ms * 1e-3 / max(r.n_fg, 1)
        grad_call_times.extend([per] * r.n_fg)
        last_run["lbfgs"] = {"fg_calls": r.n_fg, "iters": r.n_iters, "skipped": r.n_skipped, "stop_reason": r.stop_reason,
                             "loop_ms": r.loop_ms, "kernel_launches": r.kernel_launches, "driver": "device"}
        return self
