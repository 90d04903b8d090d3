"""CPU: the algorithm of the device-resident L-BFGS driver (csrc/lbfgs_kernels.cu), restated
evaluation by evaluation in numpy (oracle/lbfgs_model.py), against scipy's L-BFGS-B -- the driver
the reference uses (lbfgs.py:64-70).  Same iteration counts, same number of loss+gradient
evaluations, objective traces to 1e-10 on the golden designs; on the ill-conditioned raw design
(condition ~1e8) the traces drift like two BLAS builds do and the converged
values agree."""
import numpy as np
import pytest
from scipy.optimize import fmin_l_bfgs_b

import cases
import oracle
from oracle.lbfgs_model import lbfgs_device_model


def _both(A, b, a2, max_iter, pgtol):
    A = np.asarray(A, dtype=np.float64)
    d = A.shape[1]

    def fg(x):
        return oracle.smooth_value_and_grad(x, A, b, a2)

    tr_s, tr_m = [], []
    xs, fs, info = fmin_l_bfgs_b(func=fg, x0=np.zeros(d), maxiter=max_iter, pgtol=pgtol,
                                 callback=lambda x: tr_s.append(fg(x)[0]))
    r = lbfgs_device_model(fg, np.zeros(d), max_iter=max_iter, pgtol=pgtol, callback=lambda x: tr_m.append(fg(x)[0]))
    return xs, fs, info, np.array(tr_s), r, np.array(tr_m)


@pytest.mark.parametrize("name", ["c1", "mid", "odd", "wide"])
@pytest.mark.parametrize("a2_kind", ["none", "small", "ridge"])
def test_model_equals_scipy_on_golden_designs(name, a2_kind):
    A, b = cases.design(name)
    lam = float(np.max(np.abs(np.asarray(A, dtype=np.float64).T @ b)))
    a2 = {"none": 0.0, "small": 0.3, "ridge": 0.05 * lam}[a2_kind]
    xs, fs, info, tr_s, r, tr_m = _both(A, b, a2, 50, 1e-6)
    assert r["n_iters"] == info["nit"] and r["n_fg"] == info["funcalls"]
    # stop reason: scipy's Python wrapper tests the iteration cap BEFORE L-BFGS-B gets to its
    # convergence tests, the device driver after them; when both fire at the same iterate the
    # reasons differ (cap vs convergence) while iterates, counts and values are the same
    if info["warnflag"] == 0:
        assert r["stop"] in (1, 2)
    else:
        assert r["stop"] in (1, 2, 3) and r["n_iters"] == 50
    assert len(tr_m) == len(tr_s) == info["nit"]
    np.testing.assert_allclose(tr_m, tr_s, rtol=1e-10)
    assert abs(r["f"] - fs) <= 1e-12 * abs(fs)
    assert np.linalg.norm(r["x"] - xs) <= 1e-9 * np.linalg.norm(xs)


def test_model_on_ill_conditioned_design():
    A, b = cases.design("c1raw")
    xs, fs, info, tr_s, r, tr_m = _both(A, b, 0.0, 50, 1e-6)
    assert abs(r["n_iters"] - info["nit"]) <= 2 and r["stop"] == 2 and info["warnflag"] == 0
    np.testing.assert_allclose(tr_m[:3], tr_s[:3], rtol=1e-9)     # identical start
    assert abs(r["f"] - fs) <= 1e-9 * abs(fs)                     # same converged value


def test_model_stop_rules():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((600, 40))
    b = rng.standard_normal(600)

    def fg(x):
        return oracle.smooth_value_and_grad(x, A, b, 0.5)

    # projected-gradient stop at the start point
    x_star = np.linalg.solve(A.T @ A + 0.5 * np.eye(40), A.T @ b)
    r = lbfgs_device_model(fg, x_star, pgtol=1e-6)
    assert r["stop"] == 1 and r["n_iters"] == 0 and r["n_fg"] == 1
    # iteration cap, evaluation cap
    assert lbfgs_device_model(fg, np.zeros(40), max_iter=3, pgtol=0.0, factr=0.0)["stop"] == 3
    r = lbfgs_device_model(fg, np.zeros(40), maxfun=4, pgtol=0.0, factr=0.0)
    assert r["stop"] == 4 and r["n_fg"] == 5
    # tight tolerances: converges to the normal-equations solution
    r = lbfgs_device_model(fg, np.zeros(40), pgtol=1e-10, factr=10.0)
    assert np.linalg.norm(r["x"] - x_star) <= 1e-6 * np.linalg.norm(x_star)
    xs, fs, info = fmin_l_bfgs_b(func=fg, x0=np.zeros(40), pgtol=1e-10, factr=10.0)
    assert r["n_iters"] == info["nit"] and r["n_fg"] == info["funcalls"]
