"""CPU: the host-side logic of the drop-in layer against the golden traces of the reference, with a
numpy stand-in for the device design (the stand-in replaces ONLY the three device operators
gradient / objective / power iteration; everything under test is product code):

* LBFGSSolver with the scipy driver: alpha shortcuts, the terms of the recorded objective, history
  accumulation, final_obj_ (lbfgs.py:9-73 of the reference);
* ista() on arbitrary Python callables (the reference's loop semantics, iterative_solvers.py:65-125),
  including Armijo counts, step history and the tolerance stop;
* compute_objective's "residual first, then validate reg_type" order (objective_functions.py:13,28);
* the metric lists shared between the modules.
"""
import numpy as np
import pytest

import cases
import harness
import oracle


class _StandInDesign:
    """numpy in place of the GPU for the three one-shot operators of DeviceDesign."""

    def __init__(self, A, b):
        self.A = np.asarray(A, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)
        self.shape = self.A.shape
        self.device = 0
        self.calls = {"grad": 0, "objective": 0}

    def grad(self, x, alpha2=0.0):
        self.calls["grad"] += 1
        return oracle.smooth_value_and_grad(np.asarray(x, dtype=np.float64), self.A, self.b, alpha2)

    def objective(self, x, bits, alpha1, alpha2):
        self.calls["objective"] += 1
        x = np.asarray(x, dtype=np.float64)
        r = self.A @ x - self.b
        val = 0.5 * r.dot(r)
        if bits & 2:
            val += 0.5 * alpha2 * x.dot(x)
        if bits & 1:
            val += alpha1 * np.abs(x).sum()
        return float(val)


@pytest.fixture
def standin(monkeypatch):
    from fastoptsolver_b200 import lbfgs as LB
    from fastoptsolver_b200 import operators as OPS
    made = []

    def as_design(A, b=None, device=0):
        if isinstance(A, _StandInDesign):
            return A
        d = _StandInDesign(A, b)
        made.append(d)
        return d

    monkeypatch.setattr(LB, "as_design", as_design)
    monkeypatch.setattr(OPS, "as_design", as_design)
    return made


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("lbfgs",)))
def test_lbfgs_wrapper_against_golden(standin, name, key):
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import lbfgs as LB
    be = harness.Backend(name="product-host-logic", fista=None, fista_delta=None, ista=None, estimate_lipschitz=None,
                         lbfgs_cls=LB.LBFGSSolver, ista_callables=None,
                         ls_iters=lambda: list(S.ls_call_iters), grad_calls=lambda: len(S.grad_call_times))
    out, spec = harness.run_case(be, name, key)
    harness.check_case(out, spec, name, key, 1e-10, lbfgs_trace_rtol=1e-9)
    des = standin[-1]
    # one callback objective per iteration, one fused loss+gradient per evaluation, all timed
    assert des.calls["objective"] == len(out["hobj"])
    assert des.calls["grad"] == out["grad_num_calls"] >= len(out["hobj"])
    assert LB.grad_call_times is S.grad_call_times


def test_lbfgs_wrapper_quirks(standin):
    from fastoptsolver_b200 import lbfgs as LB
    A, b = cases.design("c1")
    with pytest.raises(ValueError, match="Unsupported reg_type='l0'"):
        LB.LBFGSSolver("l0", 1.0, 1.0)
    # elastic-net with a tiny alpha falls back to ridge / lasso (eps = 1e-8)
    s = LB.LBFGSSolver("elasticnet", 1e-9, 0.5)
    assert (s.reg_type, s.alpha1, s.alpha2) == ("ridge", 0.0, 0.5)
    s = LB.LBFGSSolver("elasticnet", 0.5, 1e-9)
    assert (s.reg_type, s.alpha1, s.alpha2) == ("lasso", 0.5, 0.0)
    s = LB.LBFGSSolver("lasso", 0.5, 7.0)
    assert s.alpha2 == 0.0
    # "lasso": the L1 term is in the recorded objective but never in what L-BFGS minimises
    s.fit(A, b)
    ols = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.linalg.norm(s.x_ - ols) <= 1e-5 * np.linalg.norm(ols)
    r = A @ s.x_ - b
    assert abs(s.final_obj_ - 0.5 * r.dot(r)) <= 1e-12 * s.final_obj_
    assert abs(s.history_[-1] - (0.5 * r.dot(r) + 0.5 * np.abs(s.x_).sum())) <= 1e-9 * s.history_[-1]
    n1 = len(s.history_)
    s.fit(A, b)
    assert len(s.history_) == 2 * n1                     # history_ is only reset in __init__
    with pytest.raises(ValueError, match="unknown L-BFGS driver"):
        LB.LBFGSSolver("ridge", 0.0, 1.0, driver="fortran").fit(A, b)


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("ista",)))
def test_ista_on_python_callables_against_golden(name, key):
    """ista() handed plain Python closures runs the reference's loop on the host."""
    from fastoptsolver_b200 import iterative_solvers as S
    be = harness.Backend(name="product-ista-callbacks", fista=None, fista_delta=None, ista=S.ista,
                         estimate_lipschitz=oracle.estimate_lipschitz, lbfgs_cls=None,
                         ista_callables=lambda A, b, a1, a2: cases.ista_callables_numpy(A, b, a1, a2, oracle.prox_l1),
                         ls_iters=lambda: list(S.ls_call_iters), grad_calls=lambda: len(S.grad_call_times))
    out, spec = harness.run_case(be, name, key)
    harness.check_case(out, spec, name, key, 1e-11)
    m = S.get_metrics()
    assert m["grad_num_calls"] == out["grad_num_calls"] and m["ls_iters_total"] == sum(out["ls_iters"])


def test_compute_objective_validates_after_the_residual_pass(standin):
    from fastoptsolver_b200 import operators as OPS
    A, b = cases.design("c1")
    x = np.arange(5, dtype=np.float64)
    for reg, (a1, a2) in {"lasso": (0.3, 9.0), "ridge": (9.0, 0.2), "elasticnet": (0.3, 0.2)}.items():
        assert abs(OPS.compute_objective(x, A, b, reg, a1, a2) - oracle.compute_objective(x, A, b, reg, a1, a2)) \
            <= 1e-12 * abs(oracle.compute_objective(x, A, b, reg, a1, a2))
    n_before = len(standin)
    with pytest.raises(ValueError, match="Unsupported reg_type='l0'"):
        OPS.compute_objective(x, A, b, "l0", 1.0, 1.0)
    assert len(standin) == n_before + 1 and standin[-1].calls["objective"] == 1   # the pass ran first
    assert isinstance(OPS.compute_objective(x, A, b, "lasso", 0.1, 0.0), np.float64)
