"""GPU: Gram-matrix mode (fp64 DMMA) and the batched multi-lambda path against numpy / the oracle."""
import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


def _design(n, d, seed):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, d))
    A = z.copy()
    A[:, 1:] += 0.5 * z[:, :-1]
    x_true = np.where(np.arange(d) % 9 == 0, 1.0, 0.0)
    b = A @ x_true + 0.3 * rng.standard_normal(n)
    return A, b


@pytest.mark.parametrize("n,d", [(3000, 128), (5003, 256), (777, 384)])
def test_gram_matches_numpy(n, d):
    from fastoptsolver_b200.design import DeviceDesign
    from fastoptsolver_b200.gram import GramDesign
    A, b = _design(n, d, 1)
    des = DeviceDesign.from_host(A, b)
    gram = GramDesign(des)
    G, c = gram.download()
    G_ref = A.T @ A
    assert harness.rel_err(G, G_ref) <= 1e-13
    assert np.array_equal(G, G.T), "G must be exactly symmetric"
    assert harness.rel_err(c, A.T @ b) <= 1e-13
    assert abs(gram.btb - b @ b) <= 1e-13 * (b @ b)
    # deterministic rebuild
    G2, _ = GramDesign(des).download()
    assert G2.tobytes() == G.tobytes()
    des.close()


def test_path_matches_reference_fista():
    """Every column of the batched path == the reference fista for that penalty."""
    import oracle
    from fastoptsolver_b200 import gram as GM
    A, b = _design(4000, 256, 2)
    lam = float(np.max(np.abs(A.T @ b)))
    alphas = 1.001 * lam * np.logspace(0, -3, 70)  # 70 penalties -> two column blocks, one padded
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    for a2 in (0.0, 0.05 * lam):
        X, info = GM.fista_path(A, b, alphas, alpha2=a2, max_iter=60, L=L + (a2 if a2 > 0 else 0.0))
        for j in (0, 1, 17, 63, 64, 69):
            # same L for both sides (estimate_lipschitz is pinned elsewhere)
            x_ref, h = _oracle_fista_fixed_L(oracle, A, b, alphas[j], a2, L + (a2 if a2 > 0 else 0.0), 60)
            assert harness.rel_err(X[j], x_ref) <= 1e-9 or np.linalg.norm(X[j] - x_ref) <= 1e-9 * np.linalg.norm(X[-1])
            assert abs(info["obj"][j] - h[-1]) <= 1e-9 * abs(h[-1])
            tie = 1e-6 * max(np.abs(x_ref).max(), 1e-300)
            big = np.abs(x_ref) > tie
            assert np.array_equal(np.sign(X[j][big]), np.sign(x_ref[big]))
    assert np.all(X[0] == 0.0)                      # alpha1 > lambda_max -> zero solution


def _oracle_fista_fixed_L(oracle, A, b, a1, a2, L, iters):
    """oracle.fista with the Lipschitz estimate replaced by a given value."""
    saved = oracle.ref_numpy.estimate_lipschitz
    oracle.ref_numpy.estimate_lipschitz = lambda A_: L - (a2 if a2 > 0 else 0.0)
    try:
        x, h = oracle.fista(A, b, "elasticnet", a1, a2, max_iter=iters, return_history=True)
    finally:
        oracle.ref_numpy.estimate_lipschitz = saved
    return x, h["obj"]


def test_path_small_lists_tolerance_and_warm_start():
    """Short penalty lists use the 32-row tile; tol stops early; a warm start from the solution
    stops at the first check."""
    import oracle
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    A, b = _design(3000, 256, 5)
    lam = float(np.max(np.abs(A.T @ b)))
    alphas = lam * np.array([0.5, 0.2, 0.05])
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    des = DeviceDesign.from_host(A, b)
    gram = GM.GramDesign(des)
    X, info = GM.fista_path(des, None, alphas, max_iter=40, L=L, gram=gram)
    assert info["tile_rows"] == 32 and info["iters"] == 40
    for j, a1 in enumerate(alphas):
        x_ref, h = _oracle_fista_fixed_L(oracle, A, b, a1, 0.0, L, 40)
        assert harness.rel_err(X[j], x_ref) <= 1e-9
        assert abs(info["obj"][j] - h[-1]) <= 1e-9 * abs(h[-1])
    # tolerance stop: the reference stops when ||x_{k+1}-x_k|| < tol; the batch stops when all columns do
    Xt, it = GM.fista_path(des, None, alphas, max_iter=5000, L=L, gram=gram, tol=1e-7, check_every=5)
    assert it["iters"] < 5000 and it["iters"] % 5 == 0 and 0 <= it["last_max_step"] < 1e-7
    iters_ref = []
    for a1 in alphas:
        saved = oracle.ref_numpy.estimate_lipschitz
        oracle.ref_numpy.estimate_lipschitz = lambda A_: L
        try:
            _, h = oracle.fista(A, b, "lasso", a1, 0.0, max_iter=5000, tol=1e-7, return_history=True)
        finally:
            oracle.ref_numpy.estimate_lipschitz = saved
        iters_ref.append(len(h["obj"]))
    # (FISTA step norms are not monotone and the batch needs every column below tol at a check,
    # so it may run a few checks past the slowest column's first crossing)
    assert max(iters_ref) <= it["iters"] <= max(iters_ref) + 25
    # warm start from the converged path: first check already passes
    Xw, iw = GM.fista_path(des, None, alphas, max_iter=5000, L=L, gram=gram, tol=1e-6, check_every=5, X0=Xt)
    assert iw["iters"] == 5
    assert harness.rel_err(Xw, Xt) <= 1e-6
    gram.close()
    des.close()


def test_warm_started_path_matches_cold_batch():
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    import oracle
    A, b = _design(3000, 256, 6)
    lam = float(np.max(np.abs(A.T @ b)))
    alphas = lam * np.logspace(-0.3, -2, 24)
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    des = DeviceDesign.from_host(A, b)
    gram = GM.GramDesign(des)
    Xc, ic = GM.fista_path(des, None, alphas, max_iter=20000, L=L, gram=gram, tol=1e-9, check_every=10)
    Xw, iw = GM.fista_path_warm(des, None, alphas, chunk=8, tol=1e-9, L=L, gram=gram)
    assert harness.rel_err(Xw, Xc) <= 1e-6
    np.testing.assert_allclose(iw["obj"], ic["obj"], rtol=1e-10)
    assert sum(iw["iters"]) > 0 and len(iw["iters"]) == 3
    gram.close()
    des.close()


def _staging_counts():
    import ctypes as C
    from fastoptsolver_b200 import _lib
    t, c = C.c_longlong(0), C.c_longlong(0)
    _lib.check(_lib.load().fos_debug_gram_staging(C.byref(t), C.byref(c)))
    return t.value, c.value


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("n,d", [(4099, 384), (40, 128), (20011, 1024)])
def test_tile_product_staging_variants(tma, n, d, monkeypatch):
    """The tile products staged by the TMA unit (tensor maps, the default) and by cp.async (FOS_GRAM_TMA=0): the
    variant asked for is the one that runs, both match numpy, ragged row counts are zero-filled."""
    from fastoptsolver_b200.design import DeviceDesign
    from fastoptsolver_b200.gram import GramDesign
    monkeypatch.setenv("FOS_GRAM_TMA", tma)
    A, b = _design(n, d, 3)
    t0, c0 = _staging_counts()
    des = DeviceDesign.from_host(A, b)
    gram = GramDesign(des)
    G, _ = gram.download()
    t1, c1 = _staging_counts()
    assert (t1 > t0, c1 > c0) == ((True, False) if tma == "1" else (False, True))
    assert harness.rel_err(G, A.T @ A) <= 1e-13
    assert np.array_equal(G, G.T)
    gram.close()
    des.close()


def _path_staging_counts():
    import ctypes as C
    from fastoptsolver_b200 import _lib
    t, c = C.c_longlong(0), C.c_longlong(0)
    _lib.check(_lib.load().fos_debug_path_staging(C.byref(t), C.byref(c)))
    return t.value, c.value


@pytest.mark.parametrize("d,n_lambda,sk", [(384, 64, "0"), (256, 200, "0"), (128, 16, "0"), (1024, 192, "1"), (1024, 256, "1"), (2048, 32, "1")])
def test_path_operand_staging_variants_give_the_same_bits(d, n_lambda, sk, monkeypatch):
    """The batched path iteration with its operands staged by the TMA unit (default) and by cp.async
    (FOS_PATH_TMA=0), on the tile schedule (32/64/128-row tiles) and the stream-K schedule (64- and 128-penalty
    tiles): the k order inside every product is the same, so X and the objectives agree bit for bit."""
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    A, b = _design(3000, d, 5)
    des = DeviceDesign.from_host(A, b)
    gram = GM.GramDesign(des)
    G, c = gram.download()
    L = float(np.linalg.eigvalsh(G)[-1]) * 1.0001
    alphas = float(np.max(np.abs(c))) * np.logspace(0, -2, n_lambda)
    monkeypatch.setenv("FOS_PATH_SK", sk)
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("FOS_PATH_TMA", tma)
        t0, c0 = _path_staging_counts()
        X, info = GM.fista_path(des, None, alphas, max_iter=25, L=L, gram=gram)
        t1, c1 = _path_staging_counts()
        assert (t1 - t0, c1 - c0) == ((1, 0) if tma == "1" else (0, 1))
        out[tma] = (X.copy(), np.asarray(info["obj"]).copy())
    assert out["1"][0].tobytes() == out["0"][0].tobytes()
    assert out["1"][1].tobytes() == out["0"][1].tobytes()
    gram.close()
    des.close()


def test_gram_and_path_at_config5_width():
    """d = 4096, 256 penalties -- the column count and penalty count of BASELINE config 5 (the
    7-split SYRK and the 128 x 64 path tiles at their real shape) on 20 000 rows: G, c, b.b against
    numpy; every one of the 256 columns against the numpy model of the Gram recurrence
    (oracle/gram_model.py, itself pinned to the reference formulation in tests/test_gram_model_cpu.py);
    three columns against the reference formulation (oracle.fista on A) directly."""
    import oracle
    from oracle import gram_model
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    n, d, n_lam, iters = 20000, 4096, 256, 40
    des = DeviceDesign.synthetic(n, d, seed=2, noise_std=0.5, rho1=0.5, rho2=0.7)
    des.standardize()
    A, b = des.download()
    gram = GM.GramDesign(des)
    G, c = gram.download()
    G_ref = A.T @ A
    assert harness.rel_err(G, G_ref) <= 1e-13
    assert np.array_equal(G, G.T)
    assert harness.rel_err(c, A.T @ b) <= 1e-13 and abs(gram.btb - b @ b) <= 1e-13 * (b @ b)
    lam = float(np.max(np.abs(c)))
    alphas = lam * np.logspace(0, -3, n_lam)
    L = float(np.linalg.eigvalsh(G_ref)[-1]) * 1.0001
    X, info = GM.fista_path(des, None, alphas, max_iter=iters, L=L, gram=gram)
    assert info["iters"] == iters and X.shape == (n_lam, d)
    X_ref, obj_ref = gram_model.fista_gram_batch(G_ref, A.T @ b, float(b @ b), alphas, 0.0, L, iters)
    scale = np.linalg.norm(X_ref[-1])
    for j in range(n_lam):
        assert np.linalg.norm(X[j] - X_ref[j]) <= 1e-9 * max(np.linalg.norm(X_ref[j]), 1e-3 * scale), j
        assert abs(info["obj"][j] - obj_ref[j]) <= 1e-9 * abs(obj_ref[j]), j
    assert np.array_equal(X == 0.0, X_ref == 0.0) or np.mean((X == 0.0) != (X_ref == 0.0)) < 1e-6
    for j in (3, 128, 255):
        x_ref, h = _oracle_fista_fixed_L(oracle, A, b, alphas[j], 0.0, L, iters)
        assert harness.rel_err(X[j], x_ref) <= 1e-9
        assert abs(info["obj"][j] - h[-1]) <= 1e-9 * abs(h[-1])
    gram.close()
    des.close()


def test_screened_path_matches_unscreened_on_device():
    """Strong-rule screening on the device path (fos_gram_subset / fos_gram_apply + the batched
    kernel on the kept features): every column equals the unscreened warm-started path to 1e-9, the
    rule discards most features for the large penalties, and an over-aggressive rule is repaired by
    the KKT re-check."""
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    import oracle
    A, b = _design(4000, 512, 7)
    lam = float(np.max(np.abs(A.T @ b)))
    # 64 penalties over 1.5 decades, 8 per chunk: the smallest penalty of a chunk is 0.76 x the previous
    # chunk's, so the rule's threshold 2 lam_min - lam_prev is positive (a chunk that spans more than a
    # factor 2 keeps every feature: the rule has nothing to say there)
    alphas = lam * np.logspace(-0.02, -1.5, 64)
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    des = DeviceDesign.from_host(A, b)
    gram = GM.GramDesign(des)
    # building blocks against numpy
    idx = np.array([0, 3, 4, 100, 257, 511])
    sub = gram.subset(idx)
    Gs, cs = sub.download()
    G, c = gram.download()
    assert sub.d == 128 and np.array_equal(Gs[:6, :6], G[np.ix_(idx, idx)]) and np.array_equal(cs[:6], c[idx])
    assert not Gs[6:].any() and not Gs[:, 6:].any() and not cs[6:].any() and sub.btb == gram.btb
    sub.close()
    Xp = np.random.default_rng(1).standard_normal((3, 512))
    assert harness.rel_err(gram.apply(Xp), Xp @ G - c) <= 1e-13
    with pytest.raises(ValueError):
        gram.subset(np.array([5, 5]))
    # the path
    Xw, iw = GM.fista_path_warm(des, None, alphas, chunk=8, tol=1e-11, L=L, gram=gram, max_iter=20000)
    Xs, isc = GM.fista_path_screened(des, None, alphas, chunk=8, tol=1e-11, L=L, gram=gram, max_iter=20000)
    scale = np.linalg.norm(Xw[-1])
    for j in range(len(alphas)):
        assert np.linalg.norm(Xs[j] - Xw[j]) <= 1e-9 * max(np.linalg.norm(Xw[j]), 1e-3 * scale), j
    np.testing.assert_allclose(isc["obj"], iw["obj"], rtol=1e-10)
    assert isc["kept"][0] < 512 // 4 and min(isc["kept"]) < 512 // 2
    Xa, ia = GM.fista_path_screened(des, None, alphas, chunk=8, tol=1e-11, L=L, gram=gram, max_iter=20000,
                                    rule_scale=8.0)
    assert sum(ia["violations"]) > 0 and max(ia["kkt_rounds"]) > 1
    for j in range(len(alphas)):
        assert np.linalg.norm(Xa[j] - Xw[j]) <= 1e-9 * max(np.linalg.norm(Xw[j]), 1e-3 * scale), j
    gram.close()
    des.close()


@pytest.mark.parametrize("d,n_lambda", [(1024, 70), (1024, 200), (2048, 33), (4096, 32)])
def test_stream_k_schedule_matches_tile_schedule(d, n_lambda, monkeypatch):
    """The stream-K schedule of the batched iteration (equal k-step ranges per SM, partial tiles combined by
    the last contributor in CTA order) against the one-tile-per-CTA schedule: same iterates to 1e-12 (the
    k-sums associate differently), same objectives, same stop iteration; bit-identical from run to run;
    and every column against the reference's fista (via the oracle) to 1e-9."""
    import oracle
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200.design import DeviceDesign
    A, b = _design(3 * d, d, 11)
    lam = float(np.max(np.abs(A.T @ b)))
    alphas = lam * np.logspace(-0.3, -1.8, n_lambda)
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    des = DeviceDesign.from_host(A, b)
    gram = GM.GramDesign(des)
    out = {}
    for sk in ("1", "0", "1"):
        monkeypatch.setenv("FOS_PATH_SK", sk)
        X, info = GM.fista_path(des, None, alphas, alpha2=0.01 * lam, max_iter=30, L=L + 0.01 * lam, gram=gram)
        Xt, it = GM.fista_path(des, None, alphas[:5], max_iter=4000, L=L, gram=gram, tol=1e-7, check_every=10)
        out.setdefault(sk, []).append((X, info["obj"], Xt, it["iters"]))
    (Xa, oa, Xta, ita), (Xb, ob, Xtb, itb) = out["1"][0], out["0"][0]
    assert harness.rel_err(Xa, Xb) <= 1e-12 and harness.rel_err(oa, ob) <= 1e-12
    assert ita == itb and harness.rel_err(Xta, Xtb) <= 1e-9
    assert out["1"][1][0].tobytes() == Xa.tobytes()
    for j in (0, n_lambda // 2, n_lambda - 1):
        np.random.seed(0)
        x_ref = oracle.fista(A, b, "elasticnet", alphas[j], 0.01 * lam, max_iter=30)
        assert harness.rel_err(Xa[j], x_ref) <= 1e-9, j
    gram.close()
    des.close()
