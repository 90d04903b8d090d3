"""GPU: the Gram matrix accumulated under the host->device upload (include/fos.h,
fos_design_upload_gram) and estimate_lipschitz / fista on top of it, against numpy and the oracle."""
import os

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


@pytest.fixture
def forced(monkeypatch):
    """Force the feature on for small designs; clear the design cache on both sides."""
    from fastoptsolver_b200 import design as D
    D.clear_cache()
    monkeypatch.setenv("FOS_UPLOAD_GRAM", "1")
    yield monkeypatch
    D.clear_cache()


@pytest.mark.parametrize("n,d,chunk", [(20000, 256, None), (20011, 128, 3000), (9001, 384, 1024), (5000, 128, 4999)])
def test_gram_under_upload_matches_numpy(forced, n, d, chunk):
    """G from the chunked upload == A^T A (several chunks, ragged last chunk and last stage), and
    GramDesign reuses it bit for bit."""
    from fastoptsolver_b200.design import DeviceDesign
    from fastoptsolver_b200.gram import GramDesign
    if chunk:
        forced.setenv("FOS_UPLOAD_GRAM_CHUNK_ROWS", str(chunk))
    A, b = _design(n, d, 3)
    des = DeviceDesign.from_host(A, b)
    info = des.upload_gram()
    assert info["state"] == 1 and info["ptr"]
    gram = GramDesign(des)
    G, c = gram.download()
    assert harness.rel_err(G, A.T @ A) <= 1e-13
    assert np.array_equal(G, G.T)
    assert harness.rel_err(c, A.T @ b) <= 1e-13
    # the uploaded matrix itself is intact
    A2, b2 = des.download()
    assert A2.tobytes() == A.tobytes() and b2.tobytes() == b.tobytes()
    # the chunked accumulation is deterministic
    des2 = DeviceDesign.from_host(A, b)
    G2, _ = GramDesign(des2).download()
    assert G2.tobytes() == G.tobytes()
    des.close()
    des2.close()


@pytest.mark.parametrize("tol,n_iter", [(1e-6, 100), (1e-3, 100), (1e-6, 7)])
def test_lipschitz_on_gram_matches_oracle(forced, tol, n_iter):
    """estimate_lipschitz on G == the reference's power iteration on A (value to 1e-12, the same
    number of steps, the same consumption of numpy's global random stream)."""
    import oracle
    from fastoptsolver_b200 import iterative_solvers as S
    A, b = _design(20000, 256, 4)
    np.random.seed(7)
    L_ref = oracle.estimate_lipschitz(A, n_iter=n_iter, tol=tol)
    after_ref = np.random.rand()
    np.random.seed(7)
    L = S.estimate_lipschitz(A, n_iter=n_iter, tol=tol)
    after = np.random.rand()
    assert S.last_run["lipschitz"]["via"] == "gram"
    assert isinstance(L, np.float64)
    assert abs(L - L_ref) <= 1e-12 * L_ref
    assert after == after_ref
    # same step count as the streaming power iteration on the same design
    it_gram = S.last_run["lipschitz"]["iters"]
    from fastoptsolver_b200 import design as D
    D.clear_cache()
    forced.setenv("FOS_UPLOAD_GRAM", "0")
    np.random.seed(7)
    L_stream = S.estimate_lipschitz(A, n_iter=n_iter, tol=tol)
    assert S.last_run["lipschitz"]["via"] == "stream"
    assert abs(L - L_stream) <= 1e-12 * L_ref
    assert S.last_run["lipschitz"]["iters"] == it_gram


def test_fista_traces_with_gram_lipschitz(forced):
    """Full fista / fista_delta traces with the Lipschitz estimate taken on G: every iterate and
    objective within 1e-10 of the oracle, identical sparsity pattern."""
    import oracle
    from fastoptsolver_b200 import iterative_solvers as S
    A, b = _design(12000, 128, 5)
    lam = float(np.max(np.abs(A.T @ b)))
    for a1, a2, kw in [(0.1 * lam, 0.0, {}), (0.05 * lam, 0.02 * lam, dict(backtracking=True, t_init_factor=2.0))]:
        np.random.seed(0)
        x_ref, h_ref = oracle.fista(A, b, "lasso", a1, a2, max_iter=40, return_history=True, **kw)
        np.random.seed(0)
        x, h = S.fista(A, b, "lasso", a1, a2, max_iter=40, return_history=True, **kw)
        assert S.last_run["lipschitz"]["via"] == "gram"
        assert len(h["x"]) == len(h_ref["x"]) and len(h["obj"]) == len(h_ref["obj"])
        scale = np.linalg.norm(x_ref)
        for xa, xb in zip(h["x"], h_ref["x"]):
            assert np.linalg.norm(xa - xb) <= 1e-10 * scale
        np.testing.assert_allclose(h["obj"], h_ref["obj"], rtol=1e-10)
        assert np.array_equal(x == 0.0, x_ref == 0.0)
        np.random.seed(0)
        xd_ref, hd_ref = oracle.fista_delta(A, b, "elasticnet", a1, a2, 3.0, max_iter=30, return_history=True, **kw)
        np.random.seed(0)
        xd, hd = S.fista_delta(A, b, "elasticnet", a1, a2, 3.0, max_iter=30, return_history=True, **kw)
        assert harness.rel_err(xd, xd_ref) <= 1e-10
        np.testing.assert_allclose(hd["obj"], hd_ref["obj"], rtol=1e-10)


def test_automatic_threshold_and_invalidation():
    """Auto mode: a tall >= 1 GB float64 design gets a Gram matrix under its upload, small /
    float32 / odd-width designs do not; z-scoring in place discards it."""
    from fastoptsolver_b200 import design as D
    from fastoptsolver_b200 import iterative_solvers as S
    os.environ.pop("FOS_UPLOAD_GRAM", None)
    D.clear_cache()
    A, b = _design(3000, 128, 6)
    des = D.DeviceDesign.from_host(A, b)
    assert des.upload_gram()["state"] == 0
    des.close()
    rng = np.random.default_rng(8)
    n, d = 131072, 1024                                   # 1.07 GB, n = 128 d
    A = rng.standard_normal((n, d))
    A[:, ::3] *= 1.5
    b = rng.standard_normal(n)
    des = D.DeviceDesign.from_host(A, b)
    info = des.upload_gram()
    assert info["state"] == 1 and info["copy_ms"] > 0
    np.random.seed(1)
    L = S.estimate_lipschitz(des)
    assert S.last_run["lipschitz"]["via"] == "gram"
    # reference power iteration (numpy) on the same matrix
    np.random.seed(1)
    v = np.random.randn(d)
    v /= np.linalg.norm(v)
    prev = 0.0
    for _ in range(100):
        w = A.T @ (A @ v)
        L_ref = np.linalg.norm(w)
        v = w / L_ref
        if abs(L_ref - prev) < 1e-6:
            break
        prev = L_ref
    assert abs(L - L_ref) <= 1e-12 * L_ref
    des.standardize()
    assert des.upload_gram()["state"] == 0
    np.random.seed(1)
    S.estimate_lipschitz(des)
    assert S.last_run["lipschitz"]["via"] == "stream"
    des.close()
    des32 = D.DeviceDesign.from_host(A[:70000].astype(np.float32), b[:70000])
    assert des32.upload_gram()["state"] == 0
    des32.close()


@pytest.mark.parametrize("n,d,dtype", [(70001, 37, np.float64), (40000, 640, np.float32), (3, 5, np.float64),
                                       (300000, 128, np.float64)])
def test_staged_upload_of_pageable_arrays(monkeypatch, n, d, dtype):
    """The threaded pinned-staging copy used for large pageable (plain numpy) sources delivers the
    same bytes as the direct copy: forced on for small shapes, ragged last piece, padded leading
    dimension (odd d goes through the strided path and must be unaffected), with and without the
    Gram accumulation riding on it."""
    from fastoptsolver_b200 import design as D
    D.clear_cache()
    rng = np.random.default_rng(11)
    A = rng.standard_normal((n, d)).astype(dtype)
    b = rng.standard_normal(n)
    monkeypatch.setenv("FOS_UPLOAD_STAGED", "1")
    for gram in ("0", "1"):
        monkeypatch.setenv("FOS_UPLOAD_GRAM", gram)
        des = D.DeviceDesign.from_host(A, b)
        A2, b2 = des.download()
        assert A2.tobytes() == A.tobytes() and b2.tobytes() == b.tobytes()
        if gram == "1" and dtype == np.float64 and d % 128 == 0:
            assert des.upload_gram()["state"] == 1
            from fastoptsolver_b200.gram import GramDesign
            G, _ = GramDesign(des).download()
            assert harness.rel_err(G, A.T @ A) <= 1e-13
        x = rng.standard_normal(d)
        loss, g = des.grad(x)
        r = A.astype(np.float64) @ x - b
        assert harness.rel_err(g, A.astype(np.float64).T @ r) <= 1e-12
        des.close()
