"""Run a parity case through a backend that exposes the reference's module API
(the oracle, or the CUDA drop-in) and compare with the golden traces."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable

import numpy as np

import cases


@dataclass
class Backend:
    name: str
    fista: Callable
    fista_delta: Callable
    ista: Callable
    estimate_lipschitz: Callable
    lbfgs_cls: type
    ista_callables: Callable      # (A, b, a1, a2) -> (g, grad_g, prox_h)
    ls_iters: Callable            # () -> list of shrink counts of the last call
    grad_calls: Callable          # () -> number of gradient calls of the last call


def oracle_backend():
    import oracle

    return Backend(
        name="oracle",
        fista=oracle.fista, fista_delta=oracle.fista_delta, ista=oracle.ista,
        estimate_lipschitz=oracle.estimate_lipschitz, lbfgs_cls=oracle.LBFGSSolver,
        ista_callables=lambda A, b, a1, a2: cases.ista_callables_numpy(A, b, a1, a2, oracle.prox_l1),
        ls_iters=lambda: list(oracle.METRICS["ls_iters"]),
        grad_calls=lambda: len(oracle.METRICS["grad_times"]),
    )


def cuda_backend():
    """The product: drop-in modules backed by libfos_b200.so."""
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import lbfgs as LB
    from fastoptsolver_b200 import operators as OPS

    return Backend(
        name="cuda",
        fista=S.fista, fista_delta=S.fista_delta, ista=S.ista,
        estimate_lipschitz=S.estimate_lipschitz, lbfgs_cls=LB.LBFGSSolver,
        ista_callables=lambda A, b, a1, a2: OPS.ista_callables(A, b, a1, a2),
        ls_iters=lambda: list(S.ls_call_iters),
        grad_calls=lambda: len(S.grad_call_times),
    )


_GOLD = {}


def golden(name):
    if name not in _GOLD:
        with np.load(os.path.join(cases.GOLDEN_DIR, f"traces_{name}.npz")) as z:
            _GOLD[name] = {k: z[k] for k in z.files}
    return _GOLD[name]


def golden_keys(name):
    g = golden(name)
    return sorted({k.rsplit("/", 1)[0] for k in g if "/" in k})


def all_case_ids(solvers=None, designs=None, include_cpu_only=False):
    """(design, key) pairs of the golden traces.  Designs flagged ``cpu_only`` (staged: pinned on the
    CPU, not yet run on a GPU) are left out unless asked for."""
    out = []
    for name in cases.DESIGNS:
        if designs and name not in designs:
            continue
        if cases.DESIGNS[name].get("cpu_only") and not include_cpu_only:
            continue
        for key in golden_keys(name):
            if solvers and key.split("/")[0] not in solvers:
                continue
            out.append((name, key))
    return out


def run_case(backend: Backend, name: str, key: str):
    A, b = cases.design(name)
    spec = cases.solver_specs(name, A, b)[key]
    g = golden(name)
    a1, a2 = (float(v) for v in g[f"{key}/alpha"])
    d = A.shape[1]
    kind = spec["solver"]
    out = {}
    np.random.seed(spec["np_seed"])
    if kind == "fista":
        x, h = backend.fista(A, b, spec["reg_type"], a1, a2, return_history=True, **spec["kw"])
        out.update(x=x, hx=h["x"], hobj=h["obj"])
    elif kind == "fista_delta":
        x, h = backend.fista_delta(A, b, spec["reg_type"], a1, a2, spec["delta"],
                                   return_history=True, **spec["kw"])
        out.update(x=x, hx=h["x"], hobj=h["obj"])
    elif kind == "ista":
        L = backend.estimate_lipschitz(A)
        if a2 > 0:
            L += a2
        gfun, grad_g, prox_h = backend.ista_callables(A, b, a1, a2)
        x, h = backend.ista(cases.ista_start(spec, d), gfun, grad_g, prox_h, L, return_history=True, **spec["kw"])
        out.update(x=x, hx=h["x"], ht=h["t"], hdelta=h["delta"], L=L)
    elif kind == "lbfgs":
        s = backend.lbfgs_cls(spec["reg_type"], a1, a2, **spec["kw"])
        s.fit(A, b)
        out.update(x=s.x_, final_obj=s.final_obj_, hobj=list(s.history_),
                   norm_reg=np.array([s.alpha1, s.alpha2]), norm_kind=s.reg_type)
    out["ls_iters"] = backend.ls_iters()
    out["grad_num_calls"] = backend.grad_calls()
    return out, spec


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.linalg.norm(b.ravel())), 1e-300)
    return float(np.linalg.norm((a - b).ravel())) / den


def check_case(out, spec, name, key, rtol, *, lbfgs_trace_rtol=None, check_signs=True):
    """Compare a backend's outputs with the golden arrays of (name, key).

    rtol: relative l2 tolerance on every iterate and on the objective trace.
    Sign/sparsity pattern must be identical away from threshold ties
    (north_star): entries whose golden magnitude is above ``tie`` must have the
    same sign, entries that are exactly zero in the golden must be zero or tiny.
    """
    g = golden(name)
    ref = {k.rsplit("/", 1)[1]: v for k, v in g.items() if k.startswith(key + "/")}
    kind = spec["solver"]
    x_ref = ref["x"]
    if kind != "lbfgs":
        assert len(out["hx"]) == ref["hx"].shape[0], "history x length"
        for k_it, (xa, xb) in enumerate(zip(out["hx"], ref["hx"])):
            assert isinstance(xa, np.ndarray) and xa.dtype == np.float64
            scale = max(np.linalg.norm(xb), np.linalg.norm(x_ref), 1e-300)
            assert np.linalg.norm(xa - xb) <= rtol * scale, (
                f"{name}:{key} iterate {k_it}: rel {np.linalg.norm(xa - xb) / scale:.3e}")
        # Armijo shrink counts must match exactly -- except from the first line search
        # that needs > 20 halvings on: there the reference's rule (no ||diff||^2/2t term,
        # iterative_solvers.py:191) only passes once t*|grad.diff| has sunk below the
        # rounding noise of g(y), so the count is decided by the BLAS summation order (it
        # differs between two numpy builds too) while the iterates no longer move.
        ref_ls = list(ref["ls_iters"])
        cut = next((i for i, v in enumerate(ref_ls) if v > 20), len(ref_ls))
        assert len(out["ls_iters"]) == len(ref_ls), "number of line searches differs"
        assert list(out["ls_iters"])[:cut] == ref_ls[:cut], "Armijo shrink counts differ"
        assert out["grad_num_calls"] == int(ref["grad_num_calls"])
    if "hobj" in ref and kind != "lbfgs":
        assert len(out["hobj"]) == len(ref["hobj"])
        if len(ref["hobj"]):
            err = np.abs(np.asarray(out["hobj"]) - ref["hobj"]) / np.maximum(np.abs(ref["hobj"]), 1e-300)
            assert err.max() <= rtol, f"{name}:{key} objective trace rel {err.max():.3e}"
    if kind == "ista":
        assert len(out["ht"]) == len(ref["ht"]) and len(out["hdelta"]) == len(ref["hdelta"])
        # step sizes are comparable up to the first rounding-decided line search (see above)
        ref_ls = list(ref["ls_iters"])
        cut = next((i for i, v in enumerate(ref_ls) if v > 20), len(ref["ht"]) - 1)
        np.testing.assert_allclose(out["ht"][: cut + 1], ref["ht"][: cut + 1], rtol=rtol)
        # delta_k = ||x_{k+1} - x_k|| is a difference of iterates: absolute tolerance on the
        # scale of the iterates (a converged run has deltas at rounding level)
        np.testing.assert_allclose(out["hdelta"], ref["hdelta"], rtol=1e-7,
                                   atol=10 * rtol * max(np.linalg.norm(x_ref), 1e-300))
    if kind == "lbfgs":
        t = lbfgs_trace_rtol or rtol
        assert out["norm_kind"] == str(ref["norm_kind"])
        np.testing.assert_array_equal(out["norm_reg"], ref["norm_reg"])
        assert len(out["hobj"]) == len(ref["hobj"]), "L-BFGS iteration count"
        err = np.abs(np.asarray(out["hobj"]) - ref["hobj"]) / np.maximum(np.abs(ref["hobj"]), 1e-300)
        assert err.max() <= t, f"{name}:{key} L-BFGS objective trace rel {err.max():.3e}"
        assert abs(out["final_obj"] - float(ref["final_obj"])) <= t * abs(float(ref["final_obj"]))
        assert rel_err(out["x"], x_ref) <= max(t, 1e-8)
        return
    assert rel_err(out["x"], x_ref) <= rtol, f"{name}:{key} final x rel {rel_err(out['x'], x_ref):.3e}"
    if check_signs:
        tie = 1e3 * rtol * max(np.abs(x_ref).max(), 1e-300)
        xa = np.asarray(out["x"])
        big = np.abs(x_ref) > tie
        assert np.array_equal(np.sign(xa[big]), np.sign(x_ref[big])), "sign pattern differs"
        zero = x_ref == 0.0
        assert np.all(np.abs(xa[zero]) <= tie), "sparsity pattern differs"
