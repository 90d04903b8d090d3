"""GPU: streaming multi-RHS mode (fos_mrhs_fista: batched FISTA without the Gram matrix, both
contractions on the fp64 tensor cores inside one cluster kernel) against the oracle's fista per
column, and against the Gram-mode path."""
import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


def _design(n, d, seed):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, d))
    A = z.copy()
    A[:, 1:] += 0.5 * z[:, :-1]
    A[:, ::5] *= 2.0
    x_true = np.where(np.arange(d) % 9 == 0, 1.0, 0.0)
    b = A @ x_true + 0.3 * rng.standard_normal(n)
    return A, b


@pytest.mark.parametrize("n,d,n_lambda,cluster", [(6000, 1024, 11, "4"), (3001, 2048, 8, "4"), (2500, 4096, 3, "4"),
                                                  (2500, 4096, 3, "8"), (9001, 4096, 9, "8"), (9001, 4096, 9, "4")])
def test_every_column_matches_reference_fista(n, d, n_lambda, cluster, monkeypatch):
    """Column l == fista(A, b, ..., alphas1[l], alpha2) of the reference (via the oracle) to 1e-10:
    iterates after a fixed iteration count and the objective; row counts that are not multiples of
    the 8-row tile or of the 37 row blocks, a padded last batch (11 = 8 + 3 penalties); clusters of 4
    (column quarters) and, at d = 4096, of 8 (column eighths, FOS_MRHS_CLUSTER)."""
    import oracle
    monkeypatch.setenv("FOS_MRHS_CLUSTER", cluster)
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    A, b = _design(n, d, 3)
    lam = float(np.max(np.abs(A.T @ b)))
    alphas = lam * np.logspace(-0.3, -1.5, n_lambda)
    a2 = 0.01 * lam
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A) + a2
    des = DeviceDesign.from_host(A, b)
    X, info = GM.fista_path_stream(des, None, alphas, alpha2=a2, max_iter=25, L=L)
    assert info["iters"] == 25 and info["batches"] == (n_lambda + 7) // 8
    for j, a1 in enumerate(alphas):
        np.random.seed(0)
        x_ref, h = oracle.fista(A, b, "elasticnet", a1, a2, max_iter=25, return_history=True)
        assert harness.rel_err(X[j], x_ref) <= 1e-10, j
        assert abs(info["obj"][j] - h["obj"][-1]) <= 1e-10 * abs(h["obj"][-1]), j
        tie = 1e-7 * np.abs(x_ref).max()
        big = np.abs(x_ref) > tie
        assert np.array_equal(np.sign(X[j][big]), np.sign(x_ref[big]))
    # bit-reproducible
    X2, _ = GM.fista_path_stream(des, None, alphas, alpha2=a2, max_iter=25, L=L)
    assert X2.tobytes() == X.tobytes()
    des.close()


def test_matches_gram_mode_warm_start_and_tolerance():
    """Same solutions as the Gram-mode path (1e-9: G y - c vs A^T(A y - b) round differently), the
    step-norm stop rule, and a warm start."""
    import oracle
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    A, b = _design(8000, 1024, 5)
    lam = float(np.max(np.abs(A.T @ b)))
    alphas = lam * np.logspace(-0.5, -1.2, 16)
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    des = DeviceDesign.from_host(A, b)
    Xg, ig = GM.fista_path(des, None, alphas, max_iter=60, L=L)
    Xs, is_ = GM.fista_path_stream(des, None, alphas, max_iter=60, L=L)
    assert harness.rel_err(Xs, Xg) <= 1e-9
    np.testing.assert_allclose(is_["obj"], ig["obj"], rtol=1e-9)
    Xt, it = GM.fista_path_stream(des, None, alphas, max_iter=5000, L=L, tol=1e-6, check_every=5)
    assert it["iters"] < 5000 and it["iters"] % 5 == 0 and it["last_max_step"] < 1e-6
    Xw, iw = GM.fista_path_stream(des, None, alphas, max_iter=5000, L=L, tol=1e-6, check_every=5, X0=Xt)
    assert iw["iters"] <= 10
    assert harness.rel_err(Xw, Xt) <= 1e-5
    des.close()


def test_unsupported_shapes_fail_loudly():
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    A, b = _design(900, 640, 1)
    des = DeviceDesign.from_host(A, b)
    with pytest.raises(Exception) as ei:
        GM.fista_path_stream(des, None, [1.0], max_iter=3, L=1.0)
    assert "multi-RHS" in str(ei.value)
    des.close()
