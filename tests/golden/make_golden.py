#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the dev container only (needs /root/reference, which is absent on the GPU
box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own,
so these files are the pins: the oracle (oracle/ref_numpy.py) is checked against
them on CPU, the CUDA path is checked against them on the GPU.

Inputs are either stored in the fixture (small cases) or rebuilt from a seed with
elementwise-only numpy Generator calls (bit-stable across machines); see
``tests/cases.py`` which both this script and the tests import.
"""
from __future__ import annotations

import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

import iterative_solvers as ref_is  # noqa: E402  (the reference)
import lbfgs as ref_lbfgs  # noqa: E402
import objective_functions as ref_obj  # noqa: E402
import prox_operators as ref_prox  # noqa: E402
import easy_boston_data as ref_data  # noqa: E402

import cases  # noqa: E402  (tests/cases.py)

assert ref_is.__file__.startswith("/root/reference"), ref_is.__file__


def _stack(xs, d):
    return np.stack(xs) if len(xs) else np.zeros((0, d))


def run_solver_case(A, b, spec):
    """Run one reference solver per ``spec`` (see cases.solver_specs) and return a
    flat dict of arrays."""
    kind = spec["solver"]
    a1, a2 = spec["alpha1"], spec["alpha2"]
    d = A.shape[1]
    out = {}
    np.random.seed(spec.get("np_seed", 0))
    if kind == "fista":
        x, h = ref_is.fista(A, b, spec["reg_type"], a1, a2, return_history=True, **spec["kw"])
        out.update(x=x, hx=_stack(h["x"], d), hobj=np.array(h["obj"], dtype=np.float64))
    elif kind == "fista_delta":
        x, h = ref_is.fista_delta(A, b, spec["reg_type"], a1, a2, spec["delta"],
                                  return_history=True, **spec["kw"])
        out.update(x=x, hx=_stack(h["x"], d), hobj=np.array(h["obj"], dtype=np.float64))
    elif kind == "ista":
        L = ref_is.estimate_lipschitz(A)
        if a2 > 0:
            L += a2
        g, grad_g, prox_h = cases.ista_callables_numpy(A, b, a1, a2, ref_prox.prox_l1)
        x, h = ref_is.ista(cases.ista_start(spec, d), g, grad_g, prox_h, L, return_history=True, **spec["kw"])
        out.update(x=x, hx=_stack(h["x"], d), ht=np.array(h["t"]), hdelta=np.array(h["delta"]),
                   L=np.float64(L))
    elif kind == "lbfgs":
        s = ref_lbfgs.LBFGSSolver(spec["reg_type"], a1, a2, **spec["kw"])
        s.fit(A, b)
        out.update(x=s.x_, final_obj=np.float64(s.final_obj_), hobj=np.array(s.history_),
                   norm_reg=np.array([s.alpha1, s.alpha2]),
                   norm_kind=np.array(s.reg_type))
    else:
        raise KeyError(kind)
    out["alpha"] = np.array([a1, a2], dtype=np.float64)
    m = ref_is.get_metrics()
    out["grad_num_calls"] = np.int64(m["grad_num_calls"])
    out["ls_num_calls"] = np.int64(m["ls_num_calls"])
    out["ls_iters"] = np.array(list(ref_is.ls_call_iters), dtype=np.int64)
    return out


def traces(names):
    for name in names:
        A, b = cases.design(name)
        blob = {}
        if cases.DESIGNS[name].get("store"):
            blob["A"] = A
            blob["b"] = b
        for key, spec in cases.solver_specs(name, A, b).items():
            res = run_solver_case(A, b, spec)
            for k, val in res.items():
                blob[f"{key}/{k}"] = val
            print(f"{name:10s} {key:40s} iters={len(res.get('hobj', res.get('hdelta', [])))} "
                  f"grad_calls={int(res['grad_num_calls'])} ls_total={int(res['ls_iters'].sum())}")
        np.savez_compressed(os.path.join(HERE, f"traces_{name}.npz"), **blob)


def main():
    os.makedirs(HERE, exist_ok=True)
    if len(sys.argv) > 1:
        # only the named designs' traces (leaves the other fixtures byte-identical)
        return traces(sys.argv[1:])

    # ---- data generator: reference arrays for the d == 5 check
    A, b, xt = ref_data.generate_correlated_boston_like_data()
    A2, b2, _ = ref_data.generate_correlated_boston_like_data(m=200, seed=3, noise_std=0.5,
                                                              rho1=0.5, rho2=0.7)
    np.savez_compressed(os.path.join(HERE, "datagen.npz"), A=A, b=b, x_true=xt, A2=A2, b2=b2)

    # ---- operators
    ops = {}
    v = cases.prox_probe_vector()
    ops["prox_v"] = v
    ops["prox_l1_out"] = ref_prox.prox_l1(v, 0.75)
    ops["prox_en_out"] = ref_prox.prox_elastic_net(v, 0.5, 1.5, 0.25)
    m2 = v[: 12].reshape(3, 4)
    ops["prox_l1_2d"] = ref_prox.prox_l1(m2, 0.3)
    Aop, bop = cases.design("mid")
    rng = np.random.default_rng(99)
    xop = rng.standard_normal(Aop.shape[1]) * (rng.random(Aop.shape[1]) < 0.5)
    ops["obj_x"] = xop
    for reg in ("lasso", "ridge", "elasticnet"):
        ops[f"obj_{reg}"] = np.float64(ref_obj.compute_objective(xop, Aop, bop, reg, 0.7, 0.3))
    for s in (0, 1, 7):
        np.random.seed(s)
        ops[f"lip_seed{s}"] = np.float64(ref_is.estimate_lipschitz(Aop))
    np.random.seed(0)
    ops["lip_n5"] = np.float64(ref_is.estimate_lipschitz(Aop, n_iter=5))
    r = Aop @ xop - bop
    ops["fg_loss"] = np.float64(0.5 * r.dot(r))
    ops["fg_grad"] = Aop.T @ r
    np.savez_compressed(os.path.join(HERE, "operators.npz"), **ops)

    # ---- solver traces
    traces(list(cases.DESIGNS))


if __name__ == "__main__":
    main()
