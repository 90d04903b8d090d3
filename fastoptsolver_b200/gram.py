"""Gram-matrix mode and the batched regularisation path (north_star item 4).

For n >> d the reference's per-iteration cost -- two (three with history) passes over A --
can be paid once: G = A^T A, c = A^T b, b^T b.  ``GramDesign`` builds them on the GPU with
fp64 tensor-core MMA; ``fista_path`` then runs FISTA for a whole vector of L1 penalties at
once, each iteration being one d x d x Lambda contraction with the prox / momentum update
fused into its epilogue.  Column l reproduces ``fista(A, b, "lasso"|"elasticnet",
alphas1[l], alpha2, max_iter=..., backtracking=False)`` of the reference
(iterative_solvers.py:132-245) up to the rounding of G y - c versus A^T(A y - b).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from . import iterative_solvers as S
from .design import DeviceDesign, as_design


def _destroy(handle):
    try:
        _lib.load().fos_gram_destroy(C.c_void_p(handle))
    except Exception:
        pass


class GramDesign:
    """G = A^T A, c = A^T b, b^T b resident in HBM."""

    def __init__(self, design: DeviceDesign):
        lib = _lib.load()
        out = C.c_void_p()
        _lib.check(lib.fos_gram_create(design.handle, C.byref(out)))
        self._h = C.c_void_p(out.value)
        self._finalizer = weakref.finalize(self, _destroy, out.value)
        self.design = design
        d, btb, ms, ns = C.c_int(), C.c_double(), C.c_float(), C.c_int()
        _lib.check(lib.fos_gram_info(self._h, C.byref(d), C.byref(btb), C.byref(ms), C.byref(ns)))
        self.d, self.btb, self.build_ms, self.nsplit = d.value, btb.value, ms.value, ns.value

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h is not None:
            self._finalizer()
            self._h = None

    def download(self):
        G = np.empty((self.d, self.d))
        c = np.empty(self.d)
        _lib.check(_lib.load().fos_gram_download(self._h, C.c_void_p(G.ctypes.data), C.c_void_p(c.ctypes.data)))
        return G, c

    def device_arrays(self):
        """(G, c) as zero-copy objects exposing __cuda_array_interface__ (e.g. for torch.as_tensor)."""
        g, c = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.load().fos_gram_pointers(self._h, C.byref(g), C.byref(c)))
        return _CudaArray(g.value, (self.d, self.d)), _CudaArray(c.value, (self.d,))

    def allreduce(self, dist, group=None):
        """Row-sharded ranks: sum the local Gram matrices over the ranks (the one
        bandwidth-relevant collective of the whole path: d^2 doubles, once; a plain library
        all-reduce).  c = A^T b and b^T b are already global: they come from a pass of the
        streaming kernel, whose epilogue does the peer-memory exchange."""
        import torch
        gd, _ = self.device_arrays()
        G = torch.as_tensor(gd, device=f"cuda:{self.design.device}")
        dist.all_reduce(G, group=group)
        torch.cuda.synchronize(G.device)


class _CudaArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def fista_path(A, b, alphas1, alpha2=0.0, t_init_factor=1.0, max_iter=500, L=None, gram=None, tol=0.0,
               check_every=10, X0=None):
    """FISTA for every alpha1 in ``alphas1`` at once (fixed step, no restart).

    Returns (X, info): X has one row per penalty; info holds the objectives of the final
    iterates, the Lipschitz estimate and timings.  ``L`` defaults to the reference's estimate
    (``estimate_lipschitz``: power iteration started from numpy's global RNG) + alpha2.
    ``tol`` > 0 stops once every column's step norm ||x_{k+1}-x_k|| is below it (the step-norm
    rule of fista, iterative_solvers.py:238, tested every ``check_every`` iterations); ``X0``
    (one row per penalty) warm-starts the columns, e.g. from a neighbouring path."""
    des = as_design(A, b)
    own = gram is None
    if own:
        gram = GramDesign(des)
    alphas1 = np.ascontiguousarray(alphas1, dtype=np.float64).reshape(-1)
    if L is None:
        L = S.estimate_lipschitz(des)
        if alpha2 > 0:
            L += alpha2
    X = np.empty((alphas1.size, gram.d))
    obj = np.empty(alphas1.size)
    p = _lib.PathParams(alphas1=alphas1.ctypes.data_as(_lib.c_double_p), n_lambda=alphas1.size, alpha2=float(alpha2),
                        step=float(t_init_factor / L), max_iter=int(max_iter), tol=float(tol),
                        check_every=int(check_every))
    if X0 is not None:
        X0 = np.ascontiguousarray(X0, dtype=np.float64)
        if X0.shape != X.shape:
            raise ValueError(f"X0 must have shape {X.shape}, got {X0.shape}")
        p.X0 = X0.ctypes.data_as(_lib.c_double_p)
    r = _lib.PathResult(X=X.ctypes.data_as(_lib.c_double_p), obj=obj.ctypes.data_as(_lib.c_double_p))
    _lib.check(_lib.load().fos_gram_path_fista(gram.handle, C.byref(p), C.byref(r)))
    info = {"obj": obj, "L": float(L), "loop_ms": r.loop_ms, "build_ms": gram.build_ms, "launches": r.kernel_launches,
            "nsplit": gram.nsplit, "iters": r.n_iters, "last_max_step": r.last_max_step, "tile_rows": r.tile_rows}
    if own:
        gram.close()
    return X, info


def fista_path_warm(A, b, alphas1, alpha2=0.0, chunk=32, tol=1e-8, max_iter=5000, check_every=10, L=None, gram=None):
    """The path solved to tolerance in decreasing-penalty chunks, each chunk warm-started from the
    previous chunk's last (smallest-penalty) solution -- the sequential-with-warm-start strategy,
    batched ``chunk`` penalties at a time so the tensor-core contraction stays busy.  Returns
    (X, info) like ``fista_path``; info["iters"] lists the iterations each chunk needed."""
    des = as_design(A, b)
    own = gram is None
    if own:
        gram = GramDesign(des)
    alphas1 = np.ascontiguousarray(alphas1, dtype=np.float64).reshape(-1)
    order = np.argsort(-alphas1)                      # largest penalty (sparsest solution) first
    if L is None:
        L = S.estimate_lipschitz(des)
        if alpha2 > 0:
            L += alpha2
    X = np.zeros((alphas1.size, gram.d))
    obj = np.zeros(alphas1.size)
    iters, ms = [], 0.0
    warm = None
    for lo in range(0, alphas1.size, chunk):
        idx = order[lo: lo + chunk]
        X0 = None if warm is None else np.tile(warm, (idx.size, 1))
        Xc, info = fista_path(des, None, alphas1[idx], alpha2=alpha2, max_iter=max_iter, L=L, gram=gram, tol=tol,
                              check_every=check_every, X0=X0)
        X[idx] = Xc
        obj[idx] = info["obj"]
        iters.append(info["iters"])
        ms += info["loop_ms"]
        warm = Xc[-1]
    if own:
        gram.close()
    return X, {"obj": obj, "L": float(L), "iters": iters, "loop_ms": ms, "build_ms": gram.build_ms}
