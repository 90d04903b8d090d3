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


class _NumpySystem:
    """numpy stand-in for GramDesign (subset / apply / solve / d / close) so that the host logic of the
    screened path (fastoptsolver_b200/gram.py: screened_path) runs on the CPU; solve() is fista's
    fixed-step recurrence on (G, c) with the batch stop rule of fos_gram_path_fista."""

    def __init__(self, G, c, btb):
        self.G, self.c, self.btb, self.d = G, c, btb, G.shape[0]
        self.solves = []

    def subset(self, idx):
        s = _NumpySystem(self.G[np.ix_(idx, idx)], self.c[idx], self.btb)
        s.solves = self.solves
        return s

    def apply(self, X):
        return np.atleast_2d(X) @ self.G - self.c

    def close(self):
        pass

    def solve(self, alphas1, alpha2, step, max_iter, tol, check_every, X0=None):
        m, d = len(alphas1), self.d
        X = np.zeros((m, d)) if X0 is None else np.array(X0, dtype=np.float64)
        Y = X.copy()
        t_prev, it = 1.0, 0
        thr = step * np.asarray(alphas1)[:, None]
        for k in range(max_iter):
            grad = Y @ self.G - self.c + alpha2 * Y
            V = Y - step * grad
            Xn = np.sign(V) * np.maximum(np.abs(V) - thr, 0.0)
            t_cur = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t_prev * t_prev))
            Y = Xn + ((t_prev - 1.0) / t_cur) * (Xn - X)
            dx = np.linalg.norm(Xn - X, axis=1).max()
            X, t_prev, it = Xn, t_cur, k + 1
            if tol > 0 and (k + 1) % check_every == 0 and dx < tol:
                break
        obj = 0.5 * np.einsum("ij,ij->i", X @ self.G, X) - X @ self.c + 0.5 * self.btb
        obj += 0.5 * alpha2 * np.einsum("ij,ij->i", X, X) + np.asarray(alphas1) * np.abs(X).sum(axis=1)
        self.solves.append((d, m, it))
        return X, obj, it, 0.0


@pytest.mark.parametrize("alpha2_frac", [0.0, 0.02])
def test_screened_path_equals_unscreened(alpha2_frac):
    """Strong-rule screening + KKT repair returns the unscreened solutions (1e-9 per column), solves
    most chunks on a fraction of the features, and an over-aggressive rule is repaired by the KKT
    re-check rather than changing the answer."""
    from fastoptsolver_b200.gram import screened_path
    rng = np.random.default_rng(5)
    n, d = 3000, 160
    Z = rng.standard_normal((n, d))
    A = Z.copy()
    A[:, 1:] += 0.5 * Z[:, :-1]
    x_true = np.where(np.arange(d) % 13 == 0, 1.0, 0.0)
    b = A @ x_true + 0.5 * rng.standard_normal(n)
    G, c, btb = GM.gram(A, b)
    lam = float(np.max(np.abs(c)))
    alphas = lam * np.logspace(-0.02, -2.0, 48)
    a2 = alpha2_frac * lam
    L = float(np.linalg.eigvalsh(G)[-1]) + a2
    sys_full = _NumpySystem(G, c, btb)
    X_ref, _, _, _ = sys_full.solve(alphas, a2, 1.0 / L, 20000, 1e-11, 10)
    sys_scr = _NumpySystem(G, c, btb)
    X, log = screened_path(sys_scr, alphas, a2, 1.0 / L, chunk=6, tol=1e-11, max_iter=20000, check_every=10)
    scale = np.linalg.norm(X_ref[-1])
    for j in range(len(alphas)):
        assert np.linalg.norm(X[j] - X_ref[j]) <= 1e-9 * max(np.linalg.norm(X_ref[j]), 1e-3 * scale), j
    obj_ref = 0.5 * np.einsum("ij,ij->i", X_ref @ G, X_ref) - X_ref @ c + 0.5 * btb \
        + 0.5 * a2 * np.einsum("ij,ij->i", X_ref, X_ref) + alphas * np.abs(X_ref).sum(axis=1)
    np.testing.assert_allclose(log["obj"], obj_ref, rtol=1e-10)
    assert min(log["kept"]) < d // 2 and log["kept"][0] < d // 4      # the rule really discards
    assert sum(log["violations"]) == 0 or max(log["kkt_rounds"]) > 1
    # the contraction shrank: restricted solves ran on fewer than d features
    assert any(dd < d for dd, _, _ in sys_scr.solves)
    # a rule that discards far too much is repaired by the KKT re-check, not trusted
    sys_bad = _NumpySystem(G, c, btb)
    Xb, logb = screened_path(sys_bad, alphas, a2, 1.0 / L, chunk=6, tol=1e-11, max_iter=20000, check_every=10,
                             rule_scale=8.0)
    assert sum(logb["violations"]) > 0 and max(logb["kkt_rounds"]) > 1
    for j in range(len(alphas)):
        assert np.linalg.norm(Xb[j] - X_ref[j]) <= 1e-9 * max(np.linalg.norm(X_ref[j]), 1e-3 * scale), j
    # penalties given in arbitrary order come back in the caller's order
    perm = rng.permutation(len(alphas))
    Xp, _ = screened_path(_NumpySystem(G, c, btb), alphas[perm], a2, 1.0 / L, chunk=6, tol=1e-11, max_iter=20000)
    assert np.linalg.norm(Xp - X_ref[perm]) <= 1e-9 * np.linalg.norm(X_ref)
