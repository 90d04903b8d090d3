"""CPU: the host-side logic of the drop-in layer against the golden traces of the reference, with a
numpy stand-in for the device design (the stand-in replaces ONLY the three device operators
gradient / objective / power iteration; everything under test is product code):

* LBFGSSolver with the scipy driver: alpha shortcuts, the terms of the recorded objective, history
  accumulation, final_obj_ (lbfgs.py:9-73 of the reference);
* ista() on arbitrary Python callables (the reference's loop semantics, iterative_solvers.py:65-125),
  including Armijo counts, step history and the tolerance stop;
* compute_objective's "residual first, then validate reg_type" order (objective_functions.py:13,28);
* the metric lists shared between the modules;
* fista / fista_delta / device-loop ista on top of a numpy model of the ``fos_prox_grad`` contract
  (oracle/pg_model.py) plugged in where the CUDA library sits: which terms each wrapper records,
  L + alpha2, history layouts, metrics, exceptions -- every golden trace.
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
        self.closed = False

    def close(self):
        self.closed = True

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


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("lbfgs",), include_cpu_only=True))
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


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("ista",), include_cpu_only=True))
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


# ------------------------------------------------------------------------------------------------
# fista / fista_delta / ista (device-loop flavour) with the engine replaced by its numpy model
# ------------------------------------------------------------------------------------------------
class _StandInDesignPG(_StandInDesign):
    handle = None

    def power_iter(self, v0, n_iter=100, tol=1e-6):
        v = np.asarray(v0, dtype=np.float64)
        prev, L, steps = 0.0, 0.0, 0
        for _ in range(n_iter):
            w = self.A.T @ (self.A @ v)
            L = np.linalg.norm(w)
            v = w / L
            steps += 1
            if abs(L - prev) < tol:
                break
            prev = L
        return float(L), steps, 0.0

    def upload_gram(self):
        return {"ptr": None, "state": 0, "copy_ms": 0.0, "tail_ms": 0.0}

    def comm_info(self):
        return 0, 1


@pytest.fixture
def engine_model(monkeypatch):
    """The product's Python layer with (a) a numpy design and (b) a fake library whose fos_prox_grad
    reads the real PGParams struct, runs oracle/pg_model.py and fills the real PGResult buffers."""
    import types

    from fastoptsolver_b200 import _lib
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import operators as OPS
    from oracle.pg_model import prox_grad_model

    def as_design(A, b=None, device=0):
        if isinstance(A, _StandInDesignPG):
            return A
        des = _StandInDesignPG(A, b)
        des.handle = des
        return des

    def fos_prox_grad(handle, p_ref, r_ref):
        p, r = p_ref._obj, r_ref._obj
        des = handle
        d = des.shape[1]
        x0 = np.ctypeslib.as_array(p.x0, shape=(d,)).copy() if p.x0 else None
        out = prox_grad_model(des.A, des.b, scheme=p.scheme, alpha1=p.alpha1, alpha2=p.alpha2, obj_terms=p.obj_terms,
                              delta=p.delta, backtracking=bool(p.backtracking), eta=p.eta, armijo_c=p.armijo_c,
                              step0=p.step0, max_iter=p.max_iter, tol=p.tol, tol_ratio=p.tol_ratio,
                              adaptive_restart=bool(p.adaptive_restart), restart_threshold=p.restart_threshold, x0=x0)
        it, K = out["n_iters"], p.max_iter
        np.ctypeslib.as_array(r.x, shape=(d,))[:] = out["x"]
        if p.want_history and r.x_hist:
            np.ctypeslib.as_array(r.x_hist, shape=(K + 1, d))[: it + 1] = np.array(out["x_hist"])
        if it:
            np.ctypeslib.as_array(r.obj_hist, shape=(max(K, 1),))[:it] = out["obj_hist"]
            np.ctypeslib.as_array(r.step_hist, shape=(max(K, 1),))[:it] = out["step_hist"]
            np.ctypeslib.as_array(r.ls_iters, shape=(max(K, 1),))[:it] = out["ls_iters"]
            np.ctypeslib.as_array(r.ls_ms, shape=(max(K, 1),))[:it] = 1e-3
        np.ctypeslib.as_array(r.t_hist, shape=(K + 1,))[: it + 1] = out["t_hist"]
        np.ctypeslib.as_array(r.grad_ms, shape=(K + 1,))[: out["n_grad_calls"]] = 1e-3
        r.n_iters, r.n_grad_calls, r.n_passes, r.stop_reason = it, out["n_grad_calls"], out["n_grad_calls"], out["stop_reason"]
        return 0

    fake = types.SimpleNamespace(fos_prox_grad=fos_prox_grad)
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(S, "as_design", as_design)
    monkeypatch.setattr(S, "find_by_matrix", lambda A, device=0: A if isinstance(A, _StandInDesignPG) else None)
    monkeypatch.setattr(OPS, "as_design", as_design)
    return harness.Backend(name="product-on-engine-model", fista=S.fista, fista_delta=S.fista_delta, ista=S.ista,
                           estimate_lipschitz=S.estimate_lipschitz, lbfgs_cls=None, ista_callables=OPS.ista_callables,
                           ls_iters=lambda: list(S.ls_call_iters), grad_calls=lambda: len(S.grad_call_times))


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("fista", "fista_delta", "ista"), include_cpu_only=True))
def test_python_layer_on_engine_model_against_golden(engine_model, name, key):
    out, spec = harness.run_case(engine_model, name, key)
    harness.check_case(out, spec, name, key, 1e-10)


def test_python_layer_quirks_on_engine_model(engine_model):
    be = engine_model
    A, b = cases.design("c1")
    np.random.seed(0)
    x1 = be.fista(A, b, "bogus", 1.0, 0.0, max_iter=5)           # reg_type accepted and ignored
    np.random.seed(0)
    x2 = be.fista(A, b, "lasso", 1.0, 0.0, max_iter=5)
    np.testing.assert_array_equal(x1, x2)
    with pytest.raises(AssertionError):
        be.fista_delta(A, b, "lasso", 1.0, 0.0, 2.0)
    np.random.seed(0)
    be.fista_delta(A, b, "bogus", 1.0, 0.0, 3.0, max_iter=3)     # unknown reg_type only matters with history
    with pytest.raises(ValueError, match="Unsupported reg_type='bogus'"):
        be.fista_delta(A, b, "bogus", 1.0, 0.0, 3.0, max_iter=3, return_history=True)
    np.random.seed(0)
    _, h = be.fista(A, b, "lasso", 1.0, 0.0, max_iter=7, return_history=True)
    assert (len(h["x"]), len(h["obj"])) == (8, 7)
    h["x"][0][0] = 123.0
    assert h["x"][1][0] != 123.0                                  # independent copies
    np.random.seed(0)
    _, h = be.fista_delta(A, b, "lasso", 1.0, 0.0, 3.0, max_iter=7, return_history=True)
    assert (len(h["x"]), len(h["obj"])) == (7, 7)
    np.random.seed(0)
    x0, h0 = be.fista(A, b, "lasso", 1.0, 0.0, max_iter=0, return_history=True)
    assert np.all(x0 == 0) and len(h0["x"]) == 1 and h0["obj"] == []
    # the reference draws randn(d) once per solver call: the global stream advances identically
    np.random.seed(5)
    be.fista(A, b, "lasso", 1.0, 0.0, max_iter=2)
    after = np.random.rand()
    np.random.seed(5)
    oracle.fista(A, b, "lasso", 1.0, 0.0, max_iter=2)
    assert after == np.random.rand()
