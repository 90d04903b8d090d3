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

    def __init__(self, design: DeviceDesign, _handle=None):
        lib = _lib.load()
        if _handle is None:
            out = C.c_void_p()
            _lib.check(lib.fos_gram_create(design.handle, C.byref(out)))
            _handle = out.value
        self._h = C.c_void_p(_handle)
        self._finalizer = weakref.finalize(self, _destroy, _handle)
        self.design = design
        d, btb, ms, ns = C.c_int(), C.c_double(), C.c_float(), C.c_int()
        _lib.check(lib.fos_gram_info(self._h, C.byref(d), C.byref(btb), C.byref(ms), C.byref(ns)))
        self.d, self.btb, self.build_ms, self.nsplit = d.value, btb.value, ms.value, ns.value

    def subset(self, idx):
        """The Gram system restricted to the (strictly increasing) feature indices ``idx``, zero
        padded to the tile width (include/fos.h: fos_gram_subset).  Used by the screened path."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = C.c_void_p()
        _lib.check(_lib.load().fos_gram_subset(self._h, C.c_void_p(idx.ctypes.data), int(idx.size), C.byref(out)))
        return GramDesign(self.design, _handle=out.value)

    def apply(self, X):
        """G x - c for every row x of X (include/fos.h: fos_gram_apply): the gradient of the smooth
        part, A^T(A x - b), without the alpha2 term."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        if X.shape[1] != self.d:
            raise ValueError(f"X must have {self.d} columns, got {X.shape[1]}")
        out = np.empty_like(X)
        _lib.check(_lib.load().fos_gram_apply(self._h, C.c_void_p(X.ctypes.data), int(X.shape[0]),
                                              C.c_void_p(out.ctypes.data)))
        return out

    def solve(self, alphas1, alpha2, step, max_iter, tol, check_every, X0=None):
        """Batched fixed-step FISTA on this system: (X, objectives, iterations, device ms)."""
        alphas1 = np.ascontiguousarray(alphas1, dtype=np.float64).reshape(-1)
        X = np.empty((alphas1.size, self.d))
        obj = np.empty(alphas1.size)
        p = _lib.PathParams(alphas1=alphas1.ctypes.data_as(_lib.c_double_p), n_lambda=alphas1.size,
                            alpha2=float(alpha2), step=float(step), max_iter=int(max_iter), tol=float(tol),
                            check_every=int(check_every))
        if X0 is not None:
            X0 = np.ascontiguousarray(X0, dtype=np.float64)
            if X0.shape != X.shape:
                raise ValueError(f"X0 must have shape {X.shape}, got {X0.shape}")
            p.X0 = X0.ctypes.data_as(_lib.c_double_p)
        r = _lib.PathResult(X=X.ctypes.data_as(_lib.c_double_p), obj=obj.ctypes.data_as(_lib.c_double_p))
        _lib.check(_lib.load().fos_gram_path_fista(self._h, C.byref(p), C.byref(r)))
        return X, obj, r.n_iters, r.loop_ms

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


def fista_path_stream(A, b, alphas1, alpha2=0.0, t_init_factor=1.0, max_iter=500, L=None, tol=0.0, check_every=10,
                      X0=None):
    """``fista_path`` without the Gram matrix: every iteration streams A once per batch of 8 penalties
    and forms A Y - b 1^T and A^T(.) on the fp64 tensor cores inside one kernel (include/fos.h:
    fos_mrhs_fista).  For designs where d^2 doubles do not fit, or n is not >> d so that building G
    costs more than the iterations it saves.  Same arguments and return value as ``fista_path``; column
    l reproduces ``fista(A, b, ..., alphas1[l], alpha2, backtracking=False)`` of the reference
    (iterative_solvers.py:132-245)."""
    des = as_design(A, b)
    alphas1 = np.ascontiguousarray(alphas1, dtype=np.float64).reshape(-1)
    if L is None:
        L = S.estimate_lipschitz(des)
        if alpha2 > 0:
            L += alpha2
    d = des.shape[1]
    X = np.empty((alphas1.size, d))
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
    _lib.check(_lib.load().fos_mrhs_fista(des.handle, C.byref(p), C.byref(r)))
    info = {"obj": obj, "L": float(L), "loop_ms": r.loop_ms, "launches": r.kernel_launches, "iters": r.n_iters,
            "last_max_step": r.last_max_step, "batches": (alphas1.size + 7) // 8}
    if des is not A:
        des.close()
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


def screened_path(system, alphas1, alpha2, step, chunk=8, tol=1e-8, max_iter=5000, check_every=10, rule_scale=1.0,
                  kkt_slack=1e-7):
    """The warm-started path with sequential strong-rule screening (Tibshirani et al. 2012) on a Gram
    system: penalties in decreasing order, ``chunk`` at a time; before a chunk, feature j is
    discarded when

        |(G x(lam_prev) - c)_j| < rule_scale * (2 lam_min_of_chunk - lam_prev)      and x_j(lam_prev) == 0,

    the chunk is solved on the kept features only (the d x d x Lambda contraction shrinks to
    |S| x |S| x Lambda), and the KKT condition |(G x - c)_j| <= alpha1 is then re-checked on the
    discarded features for every column; violators are added and the chunk is solved again (warm),
    so the result is the solution of the unscreened problem.  Every column runs fista's fixed-step
    iteration (iterative_solvers.py:199-221) on its restricted system.

    ``system`` provides subset(idx) / apply(X) / solve(...) / d (GramDesign; the CPU tests pass a numpy
    stand-in).  The rule is evaluated once per chunk with the chunk's SMALLEST penalty, so it only
    discards while 2 lam_min > lam_prev: a chunk must not span more than a factor 2 of the grid
    (256 log-spaced penalties over three decades: chunk <= 25), otherwise every feature is kept
    and the chunk costs what the unscreened path costs.  rule_scale > 1 discards more than the rule allows (tests use it to force the KKT
    repair loop).  Returns (X, info)."""
    alphas1 = np.ascontiguousarray(alphas1, dtype=np.float64).reshape(-1)
    d = system.d
    order = np.argsort(-alphas1, kind="stable")
    X = np.zeros((alphas1.size, d))
    obj = np.zeros(alphas1.size)
    grad_prev = system.apply(np.zeros((1, d)))[0]          # = -c
    lam_prev = float(np.max(np.abs(grad_prev)))            # lambda_max: x(lam) = 0 for lam >= it
    x_prev = np.zeros(d)
    ever_active = np.zeros(d, dtype=bool)
    log = {"kept": [], "iters": [], "kkt_rounds": [], "violations": [], "loop_ms": 0.0}
    for lo in range(0, alphas1.size, chunk):
        idx = order[lo: lo + chunk]
        lam_min = float(alphas1[idx].min())
        thr = rule_scale * (2.0 * lam_min - lam_prev)
        keep = ever_active | (np.abs(grad_prev) >= thr) if thr > 0 else np.ones(d, dtype=bool)
        warm = np.tile(x_prev, (idx.size, 1))
        rounds, iters, n_viol = 0, 0, 0
        while True:
            rounds += 1
            S = np.flatnonzero(keep)
            if S.size == 0:
                Xc = np.zeros((idx.size, d))
                oc = None
            else:
                sub = system.subset(S) if S.size < d else system
                X0 = np.zeros((idx.size, sub.d))
                X0[:, : S.size] = warm[:, S]
                Xs, oc, it, ms = sub.solve(alphas1[idx], alpha2, step, max_iter, tol, check_every, X0)
                iters += it
                log["loop_ms"] += ms
                Xc = np.zeros((idx.size, d))
                Xc[:, S] = Xs[:, : S.size]
                if sub is not system:
                    sub.close()
            G = system.apply(Xc)                            # (chunk, d): G x - c
            if S.size == d:
                break
            out = ~keep
            viol = (np.abs(G[:, out]) > alphas1[idx][:, None] * (1.0 + kkt_slack)).any(axis=0)
            if not viol.any():
                break
            n_viol += int(viol.sum())
            keep[np.flatnonzero(out)[viol]] = True
            warm = Xc
        if oc is None:                                      # nothing kept: objective of x = 0 (0.5 b^T b), via a solve of 0 iterations
            _, oc, _, _ = system.solve(alphas1[idx], alpha2, step, 0, 0.0, 1)
        X[idx] = Xc
        obj[idx] = oc
        last = int(np.argmin(alphas1[idx]))
        x_prev, grad_prev, lam_prev = Xc[last], G[last], lam_min
        ever_active |= (Xc != 0.0).any(axis=0)
        log["kept"].append(int(keep.sum()))
        log["iters"].append(iters)
        log["kkt_rounds"].append(rounds)
        log["violations"].append(n_viol)
    log["obj"] = obj
    return X, log


def fista_path_screened(A, b, alphas1, alpha2=0.0, chunk=8, tol=1e-8, max_iter=5000, check_every=10, L=None, gram=None,
                        rule_scale=1.0):
    """``fista_path_warm`` with strong-rule screening (see ``screened_path``): same solutions, the
    batched contraction runs on the kept features only.  Returns (X, info); info["kept"] lists the
    number of features each chunk was solved on."""
    des = as_design(A, b)
    own = gram is None
    if own:
        gram = GramDesign(des)
    if L is None:
        L = S.estimate_lipschitz(des)
        if alpha2 > 0:
            L += alpha2
    X, log = screened_path(gram, alphas1, alpha2, 1.0 / L, chunk=chunk, tol=tol, max_iter=max_iter,
                           check_every=check_every, rule_scale=rule_scale)
    log.update(L=float(L), build_ms=gram.build_ms, d=gram.d)
    if own:
        gram.close()
    return X, log
