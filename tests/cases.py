"""Shared definition of the parity cases: designs and solver variants.

Imported by tests/golden/make_golden.py (which runs the reference on them) and by
the tests (which run the oracle / the CUDA path on the same inputs).  The variant
grid follows the figure legend of the reference's (missing) notebook:
{lasso, elasticnet} x {fixed-t1.0, armijo-t1.0, armijo-t2.0} for ISTA, FISTA and
FISTA-delta, plus L-BFGS.
"""
from __future__ import annotations

import os

import numpy as np

from fastoptsolver_b200.datagen import generate_correlated_design, standardize

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> how the design is built; ``store`` = arrays are saved in the fixture
DESIGNS = {
    "c1": dict(store=True),      # config 1: 1000 x 5 scenario s0_n0.5_r10.5_r20.7, z-scored
    "c1raw": dict(store=True),   # 1000 x 5, seed 42 defaults, NOT standardised (L ~ 9e7)
    "mid": dict(store=False),    # 3000 x 96 correlated columns  (generic kernel, d < 512)
    "mid32": dict(store=False),  # same, A stored as float32
    "odd": dict(store=False),    # 500 x 37  (odd d: padded leading dimension on device)
    "wide": dict(store=False),   # 1500 x 640 (streaming kernel, d >= 512)
    # 2000 x 64, option combinations off the figure-legend grid (restart thresholds, ratio stops with
    # backtracking, steep / shallow Armijo factors, non-zero ISTA start, capped L-BFGS).  Runs against
    # the oracle and the host layer on the CPU and against the kernels in the GPU suite.
    "midx": dict(store=False),
    # 2400 x 600, the same off-grid option combinations on a design wide and tall enough for the
    # streaming kernel AND the persistent solve kernel (16 rows per CTA): restarts, ratio / step /
    # gradient-norm stops, steep and shallow Armijo factors, non-zero ISTA start inside ONE launch
    "widex": dict(store=False),
}


def _elementwise_design(n, d, seed, noise):
    """Bit-stable across machines: only elementwise Generator/ufunc calls."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, d))
    A = z.copy()
    A[:, 1:] += 0.6 * z[:, :-1]          # neighbour correlation
    A[:, ::7] *= 3.0                     # uneven column scales
    x_true = np.where(np.arange(d) % 5 == 0, 1.5, 0.0) - np.where(np.arange(d) % 11 == 3, 0.7, 0.0)
    b = np.zeros(n)
    for j in range(d):                   # explicit axpy loop: no BLAS rounding differences
        if x_true[j] != 0.0:
            b += A[:, j] * x_true[j]
    b += noise * rng.standard_normal(n)
    return np.ascontiguousarray(A), b


def design(name):
    """Return (A, b) for a case name (loads stored arrays when the fixture has them)."""
    if DESIGNS[name].get("store"):
        path = os.path.join(GOLDEN_DIR, f"traces_{name}.npz")
        if os.path.exists(path):
            with np.load(path) as z:
                if "A" in z:
                    return z["A"], z["b"]
    if name == "c1":
        A, b, _ = generate_correlated_design(1000, 5, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
        return standardize(A, b)
    if name == "c1raw":
        A, b, _ = generate_correlated_design(1000, 5)
        return A, b
    if name == "mid":
        return _elementwise_design(3000, 96, 11, 0.5)
    if name == "mid32":
        A, b = _elementwise_design(3000, 96, 11, 0.5)
        return A.astype(np.float32), b
    if name == "odd":
        return _elementwise_design(500, 37, 12, 1.0)
    if name == "wide":
        return _elementwise_design(1500, 640, 13, 2.0)
    if name == "midx":
        return _elementwise_design(2000, 64, 21, 0.8)
    if name == "widex":
        return _elementwise_design(2400, 600, 23, 1.0)
    raise KeyError(name)


_STEP_RULES = {
    "fixed-t1.0": dict(backtracking=False, t_init_factor=1.0),
    "armijo-t1.0": dict(backtracking=True, t_init_factor=1.0),
    "armijo-t2.0": dict(backtracking=True, t_init_factor=2.0),
}


def solver_specs(name, A, b):
    """Ordered dict key -> spec.  alpha values depend on the data (fraction of
    lambda_max); the golden file stores the exact floats used and the tests read
    them back so both sides see identical scalars."""
    lam = float(np.max(np.abs(np.asarray(A, dtype=np.float64).T @ b)))
    a1 = 0.1 * lam
    if name in ("midx", "widex"):
        return _extra_specs(lam)
    regs = {"lasso": (a1, 0.0), "elasticnet": (a1, 0.5 * a1)}
    specs = {}
    full_grid = name in ("c1", "mid")
    iters = 50
    for reg, (x1, x2) in regs.items():
        for rule, kw in _STEP_RULES.items():
            if not full_grid and rule == "armijo-t1.0":
                continue
            if not full_grid and reg == "elasticnet" and rule != "fixed-t1.0":
                continue
            base = dict(max_iter=iters, tol=0.0, **kw)
            specs[f"fista/{reg}-{rule}"] = dict(solver="fista", reg_type=reg, alpha1=x1, alpha2=x2, kw=base)
            specs[f"fista_delta/{reg}-{rule}"] = dict(solver="fista_delta", reg_type=reg, alpha1=x1,
                                                      alpha2=x2, delta=3.0, kw=base)
            if full_grid or rule == "fixed-t1.0":
                specs[f"ista/{reg}-{rule}"] = dict(solver="ista", reg_type=reg, alpha1=x1, alpha2=x2, kw=base)
    if name != "c1raw":
        specs["lbfgs/lasso"] = dict(solver="lbfgs", reg_type="lasso", alpha1=a1, alpha2=0.3, kw=dict(max_iter=50))
        specs["lbfgs/ridge"] = dict(solver="lbfgs", reg_type="ridge", alpha1=a1, alpha2=0.05 * lam, kw=dict(max_iter=50))
        specs["lbfgs/elasticnet"] = dict(solver="lbfgs", reg_type="elasticnet", alpha1=a1, alpha2=0.05 * lam,
                                         kw=dict(max_iter=50, tol=1e-8))
    if name == "mid":
        specs["fista/lasso-tol"] = dict(solver="fista", reg_type="lasso", alpha1=a1, alpha2=0.0,
                                        kw=dict(max_iter=400, tol=1e-4))
        specs["fista/lasso-restart"] = dict(solver="fista", reg_type="lasso", alpha1=0.02 * lam, alpha2=0.0,
                                            kw=dict(max_iter=120, adaptive_restart=True, restart_threshold=1.0))
        specs["fista/lasso-tolratio"] = dict(solver="fista", reg_type="lasso", alpha1=a1, alpha2=0.0,
                                             kw=dict(max_iter=200, tol_ratio=0.3))
        specs["fista/ridge-fixed"] = dict(solver="fista", reg_type="ridge", alpha1=0.0, alpha2=0.05 * lam,
                                          kw=dict(max_iter=50))
        specs["fista/ols-armijo"] = dict(solver="fista", reg_type="lasso", alpha1=0.0, alpha2=0.0,
                                         kw=dict(max_iter=30, backtracking=True, t_init_factor=4.0, eta=0.7))
        # reg_type says lasso but alpha2 > 0: alpha2 enters the gradient, not the
        # recorded objective (iterative_solvers.py:293-294 vs :321)
        specs["fista_delta/lasso-with-a2"] = dict(solver="fista_delta", reg_type="lasso", alpha1=a1,
                                                  alpha2=0.5 * a1, delta=2.5, kw=dict(max_iter=40))
        specs["fista_delta/lasso-tol"] = dict(solver="fista_delta", reg_type="lasso", alpha1=a1, alpha2=0.0,
                                              delta=4.0, kw=dict(max_iter=400, tol=1e-4))
        specs["ista/lasso-tol"] = dict(solver="ista", reg_type="lasso", alpha1=a1, alpha2=0.0,
                                       kw=dict(max_iter=400, tol=1e-3))
    for s in specs.values():
        s.setdefault("np_seed", 0)
    return specs


def _extra_specs(lam):
    """Option combinations beyond the figure-legend grid (design "midx")."""
    a1, a2 = 0.08 * lam, 0.03 * lam
    S = {}

    def add(key, solver, reg, x1, x2, kw, **extra):
        S[key] = dict(solver=solver, reg_type=reg, alpha1=x1, alpha2=x2, kw=kw, **extra)

    add("fista/restart-thr0.5", "fista", "lasso", 0.02 * lam, 0.0,
        dict(max_iter=150, adaptive_restart=True, restart_threshold=0.5))
    add("fista/restart-armijo", "fista", "elasticnet", 0.02 * lam, a2,
        dict(max_iter=80, adaptive_restart=True, restart_threshold=1.2, backtracking=True, t_init_factor=3.0))
    add("fista/tolratio-armijo", "fista", "lasso", a1, 0.0,
        dict(max_iter=200, tol_ratio=0.25, backtracking=True, t_init_factor=2.0))
    add("fista/steep-armijo", "fista", "elasticnet", a1, a2,
        dict(max_iter=40, backtracking=True, t_init_factor=8.0, eta=0.3))
    add("fista/shallow-armijo", "fista", "lasso", a1, 0.0,
        dict(max_iter=40, backtracking=True, t_init_factor=5.0, eta=0.9))
    add("fista/half-step", "fista", "lasso", a1, 0.0, dict(max_iter=60, t_init_factor=0.5))
    add("fista/one-iteration", "fista", "elasticnet", a1, a2, dict(max_iter=1))
    add("fista/gradnorm-stop", "fista", "ridge", 0.0, a2, dict(max_iter=400, tol=1e-2))
    add("fista_delta/tolratio", "fista_delta", "lasso", a1, 0.0, dict(max_iter=300, tol_ratio=0.6), delta=3.5)
    add("fista_delta/delta10-armijo", "fista_delta", "elasticnet", a1, a2,
        dict(max_iter=60, backtracking=True, t_init_factor=2.0, eta=0.8), delta=10.0)
    # reg_type picks the recorded objective only: "ridge" with alpha1 > 0 still thresholds the iterates
    add("fista_delta/ridge-objective-with-l1-prox", "fista_delta", "ridge", a1, a2, dict(max_iter=40), delta=3.0)
    add("fista_delta/step-stop", "fista_delta", "elasticnet", a1, a2, dict(max_iter=500, tol=1e-5), delta=2.5)
    add("ista/steep-armijo", "ista", "elasticnet", a1, a2,
        dict(max_iter=60, backtracking=True, t_init_factor=6.0, eta=0.3))
    add("ista/tol-armijo", "ista", "lasso", a1, 0.0, dict(max_iter=400, tol=1e-4, backtracking=True, t_init_factor=2.0))
    add("ista/nonzero-start", "ista", "lasso", a1, 0.0, dict(max_iter=50), x0_seed=17)
    add("ista/nonzero-start-armijo", "ista", "elasticnet", a1, a2,
        dict(max_iter=50, backtracking=True, t_init_factor=2.0), x0_seed=18)
    add("lbfgs/capped", "lbfgs", "ridge", a1, a2, dict(max_iter=3))
    add("lbfgs/loose-tol", "lbfgs", "elasticnet", a1, a2, dict(max_iter=200, tol=1e-1))
    add("lbfgs/tiny-alpha1", "lbfgs", "elasticnet", 1e-9, a2, dict(max_iter=100))
    add("lbfgs/tiny-alpha2", "lbfgs", "elasticnet", a1, 1e-9, dict(max_iter=100))
    for s in S.values():
        s.setdefault("np_seed", 3)
    return S


def ista_start(spec, d):
    """Start point of an ISTA case: zeros unless the spec names a seed."""
    if "x0_seed" in spec:
        return np.random.default_rng(spec["x0_seed"]).standard_normal(d) * 0.3
    return np.zeros(d)


def prox_probe_vector():
    """Edge cases for the soft threshold: +-0, exact ties |v| == tau, NaN, +-inf,
    denormals (-0.0 for shrunk negatives, NaN propagates)."""
    rng = np.random.default_rng(5)
    special = np.array([0.0, -0.0, 0.75, -0.75, 0.5, -0.5, 1.0, -1.0, np.nan, np.inf, -np.inf,
                        5e-324, -5e-324, 0.7500000000000001, -0.7499999999999999, 3.25])
    return np.concatenate([special, rng.standard_normal(48) * 2.0])


def ista_callables_numpy(A, b, alpha1, alpha2, prox_l1):
    """The closures the notebook must have built for ``ista`` (iterative_solvers.py
    :65-77 takes g, grad_g, prox_h): smooth part with the ridge term, prox of the
    L1 term with alpha1 folded into the step."""
    def g(x):
        r = A @ x - b
        val = 0.5 * r.dot(r)
        if alpha2 > 0:
            val += 0.5 * alpha2 * x.dot(x)
        return val

    def grad_g(x):
        out = A.T @ (A @ x - b)
        if alpha2 > 0:
            out += alpha2 * x
        return out

    def prox_h(v, t):
        return prox_l1(v, t * alpha1) if alpha1 > 0 else v

    return g, grad_g, prox_h
