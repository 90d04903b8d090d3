"""CPU: the Gram-matrix formulation the product uses for estimate_lipschitz (after an upload that
accumulated A^T A) and for the batched path, restated in numpy (oracle/gram_model.py), against the
reference formulation (oracle.estimate_lipschitz / oracle.fista, pinned to the golden traces)."""
import numpy as np
import pytest

import cases
import oracle
from oracle import gram_model as GM


def _tall(n, d, seed):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, d))
    A = z.copy()
    A[:, 1:] += 0.5 * z[:, :-1]
    A[:, ::5] *= 2.0
    x_true = np.where(np.arange(d) % 9 == 0, 1.0, 0.0)
    b = A @ x_true + 0.3 * rng.standard_normal(n)
    return A, b


@pytest.mark.parametrize("tol,n_iter", [(1e-6, 100), (1e-3, 100), (1e-6, 7)])
@pytest.mark.parametrize("shape", [(20000, 256), (6000, 128), (3000, 96)])
def test_power_iteration_on_gram_equals_reference(shape, tol, n_iter):
    A, b = _tall(*shape, seed=4)
    G, _, _ = GM.gram(A, b)
    np.random.seed(7)
    L_ref = oracle.estimate_lipschitz(A, n_iter=n_iter, tol=tol)
    after_ref = np.random.rand()
    np.random.seed(7)
    L, steps = GM.estimate_lipschitz_gram(G, n_iter=n_iter, tol=tol)
    assert np.random.rand() == after_ref                     # same consumption of the global stream
    assert abs(L - L_ref) <= 1e-12 * L_ref
    # same number of steps as the reference loop
    np.random.seed(7)
    v = np.random.randn(A.shape[1])
    v /= np.linalg.norm(v)
    prev, ref_steps = 0.0, 0
    for _ in range(n_iter):
        w = A.T @ (A @ v)
        Lr = np.linalg.norm(w)
        v = w / Lr
        ref_steps += 1
        if abs(Lr - prev) < tol:
            break
        prev = Lr
    assert steps == ref_steps


@pytest.mark.parametrize("name", ["mid", "wide", "c1"])
def test_power_iteration_on_golden_designs(name):
    A, b = cases.design(name)
    A = np.asarray(A, dtype=np.float64)
    G, _, _ = GM.gram(A, b)
    np.random.seed(0)
    L_ref = oracle.estimate_lipschitz(A)
    np.random.seed(0)
    L, _ = GM.estimate_lipschitz_gram(G)
    assert abs(L - L_ref) <= 1e-12 * L_ref


@pytest.mark.parametrize("a2_frac", [0.0, 0.05])
def test_fista_on_gram_equals_reference(a2_frac):
    A, b = _tall(4000, 256, seed=2)
    G, c, btb = GM.gram(A, b)
    lam = float(np.max(np.abs(c)))
    np.random.seed(0)
    L = oracle.estimate_lipschitz(A)
    a2 = a2_frac * lam
    for a1 in (0.3 * lam, 0.05 * lam, 0.0):
        np.random.seed(0)
        x_ref, h = oracle.fista(A, b, "elasticnet", a1, a2, max_iter=60, return_history=True)
        x, objs = GM.fista_gram(G, c, btb, a1, a2, L + (a2 if a2 > 0 else 0.0), 60)
        scale = max(np.linalg.norm(x_ref), 1e-300)
        assert np.linalg.norm(x - x_ref) <= 1e-9 * scale
        np.testing.assert_allclose(objs, h["obj"], rtol=1e-9)
        tie = 1e-6 * max(np.abs(x_ref).max(), 1e-300)
        big = np.abs(x_ref) > tie
        assert np.array_equal(np.sign(x[big]), np.sign(x_ref[big]))


def test_batched_model_equals_per_column_model():
    """fista_gram_batch (used by the GPU test at the config-5 shape) == fista_gram column by column."""
    rng = np.random.default_rng(9)
    A = rng.standard_normal((400, 48))
    A[:, 1:] += 0.4 * A[:, :-1]
    b = A @ np.where(np.arange(48) % 6 == 0, 1.0, 0.0) + 0.2 * rng.standard_normal(400)
    G, c, btb = GM.gram(A, b)
    lam = float(np.max(np.abs(c)))
    alphas = lam * np.array([1.2, 0.5, 0.1, 0.01, 0.0])
    L = float(np.linalg.eigvalsh(G)[-1])
    for a2 in (0.0, 0.3):
        X, obj = GM.fista_gram_batch(G, c, btb, alphas, a2, L + a2, 40)
        for j, a1 in enumerate(alphas):
            x, objs = GM.fista_gram(G, c, btb, a1, a2, L + a2, 40)
            assert np.linalg.norm(X[j] - x) <= 1e-13 * max(np.linalg.norm(x), 1.0)
            assert abs(obj[j] - objs[-1]) <= 1e-12 * abs(objs[-1])
