"""Device-backed operators with the reference's signatures.

prox_l1 / prox_elastic_net   <- prox_operators.py:3-16
compute_objective            <- objective_functions.py:3-30
ista_callables               the (g, grad_g, prox_h) triple ``ista`` takes
                             (iterative_solvers.py:65-70), as framework-owned objects so
                             that ``ista`` can run its loop on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .design import DeviceDesign, as_design

_REG_BITS = {"lasso": 1, "ridge": 2, "elasticnet": 3}


def _prox_buffers(v):
    arr = np.asarray(v, dtype=np.float64)
    flat = np.ascontiguousarray(arr).reshape(-1)
    return arr, flat, np.empty_like(flat)


def prox_l1(v, tau, device=0):
    """Soft threshold sign(v)*max(|v|-tau, 0) on the GPU; any shape (prox_operators.py:3-8)."""
    arr, flat, out = _prox_buffers(v)
    _lib.check(_lib.load().fos_prox_l1(C.c_void_p(flat.ctypes.data), flat.size, float(tau),
                                       C.c_void_p(out.ctypes.data), device))
    res = out.reshape(arr.shape)
    return res if arr.ndim else np.float64(res)


def prox_elastic_net(v, tau, alpha1, alpha2, device=0):
    """prox of alpha1||.||_1 + 0.5 alpha2 ||.||^2 (prox_operators.py:10-16)."""
    arr, flat, out = _prox_buffers(v)
    _lib.check(_lib.load().fos_prox_elastic_net(C.c_void_p(flat.ctypes.data), flat.size, float(tau),
                                                float(alpha1), float(alpha2), C.c_void_p(out.ctypes.data), device))
    res = out.reshape(arr.shape)
    return res if arr.ndim else np.float64(res)


def compute_objective(x, A, b, reg_type, alpha1, alpha2):
    """0.5||Ax-b||^2 [+0.5 a2 ||x||^2] [+a1 ||x||_1] in one pass over A
    (objective_functions.py:3-30).  As in the reference the residual pass runs before
    ``reg_type`` is validated (``:13`` precedes ``:28``)."""
    des = as_design(A, b)
    bits = _REG_BITS.get(reg_type)
    val = des.objective(x, bits if bits is not None else 0, alpha1, alpha2)
    if bits is None:
        raise ValueError(f"Unsupported reg_type='{reg_type}'")
    return np.float64(val)


# ------------------------------------------------------------------------- ISTA closures
class _LeastSquaresPart:
    def __init__(self, design: DeviceDesign, alpha1: float, alpha2: float):
        self.design, self.alpha1, self.alpha2 = design, float(alpha1), float(alpha2)


class SmoothValue(_LeastSquaresPart):
    """g(x) = 0.5||Ax-b||^2 (+0.5 a2 ||x||^2)"""

    def __call__(self, x):
        return np.float64(self.design.objective(x, 2 if self.alpha2 > 0 else 0, 0.0, self.alpha2))


class SmoothGrad(_LeastSquaresPart):
    """grad g(x) = A^T(Ax-b) (+a2 x), A read once"""

    def __call__(self, x):
        return self.design.grad(x, self.alpha2 if self.alpha2 > 0 else 0.0)[1]


class L1Prox(_LeastSquaresPart):
    """prox_h(v, t) = soft_threshold(v, t*alpha1); identity when alpha1 == 0"""

    def __call__(self, v, t):
        return prox_l1(v, t * self.alpha1, device=self.design.device) if self.alpha1 > 0 else v


def ista_callables(A, b, alpha1, alpha2):
    """(g, grad_g, prox_h) for ``ista``.  When all three come from one call of this function
    ``ista`` recognises them and keeps the whole loop on the device."""
    des = as_design(A, b)
    return SmoothValue(des, alpha1, alpha2), SmoothGrad(des, alpha1, alpha2), L1Prox(des, alpha1, alpha2)
