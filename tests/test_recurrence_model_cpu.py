"""CPU: the residual recurrence behind GM_QREC (oracle/qrec_model.py) against golden traces of the unmodified
reference: same iterates, and an objective trace that needs no product with x_k."""
import numpy as np
import pytest

import cases
import harness


def _fixed_step_cases():
    return list(harness.all_case_ids(solvers=("fista", "fista_delta"), include_cpu_only=True))


@pytest.mark.parametrize("name,key", _fixed_step_cases())
def test_objective_trace_from_the_recurrence_matches_the_reference(name, key):
    import oracle
    from oracle import qrec_model
    A, b = cases.design(name)
    spec = cases.solver_specs(name, A, b)[key]
    kw = dict(spec["kw"])
    if kw.get("backtracking") or kw.get("tol", 0.0) > 0 or kw.get("tol_ratio", 0.0) > 0:
        pytest.skip("the recurrence is used for fixed-step runs; stop rules are exercised on the device")
    g = harness.golden(name)
    a1, a2 = (float(v) for v in g[f"{key}/alpha"])
    np.random.seed(spec["np_seed"])
    L = oracle.estimate_lipschitz(np.asarray(A, dtype=np.float64))
    if a2 > 0:
        L += a2
    L /= kw.get("t_init_factor", 1.0)
    if spec["solver"] == "fista":
        terms = (1 if a1 > 0 else 0) | (2 if a2 > 0 else 0)
        x, h = qrec_model.fista_with_recurrence(A, b, a1, a2, L, kw["max_iter"], obj_terms=terms,
                                                adaptive_restart=kw.get("adaptive_restart", False),
                                                restart_threshold=kw.get("restart_threshold", 1.0))
        hx_ref = g[f"{key}/hx"]
    else:
        terms = {"lasso": 1, "ridge": 2, "elasticnet": 3}[spec["reg_type"]]
        x, h = qrec_model.fista_with_recurrence(A, b, a1, a2, L, kw["max_iter"], scheme="delta", delta=spec["delta"],
                                                obj_terms=terms)
        hx_ref = g[f"{key}/hx"][:]
        h["x"] = h["x"][1:]                     # fista_delta records no x_0 (iterative_solvers.py:320)
    obj_ref = g[f"{key}/hobj"]
    assert len(h["obj"]) == len(obj_ref)
    assert harness.rel_err(x, g[f"{key}/x"]) <= 1e-10
    assert len(h["x"]) == hx_ref.shape[0]
    err = np.abs(np.asarray(h["obj"]) - obj_ref) / np.abs(obj_ref)
    assert err.max() <= 1e-11, (name, key, err.max())
