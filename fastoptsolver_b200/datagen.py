"""Host-side synthetic designs (input fixtures for tests and benches).

``generate_correlated_design`` generalises the reference's 5-column
"Boston-like" generator (easy_boston_data.py:7-45) to d columns; with d == 5 it
consumes the numpy Generator stream in the same order (block1, block2, distance,
noise) and therefore reproduces the reference arrays bit for bit
(tests/test_oracle_golden.py::test_datagen_matches_reference).

Large designs (n >= ~500k) are generated directly in HBM by the Philox kernel
behind ``fos_design_create_synthetic`` (csrc/datagen.cu); this module is only for
sizes the host can hold.
"""
from __future__ import annotations

import numpy as np

# easy_boston_data.py:40 -- coefficient pattern of one 5-column group
_GROUP_COEF = np.array([5.0, 0.0, -0.02, -0.05, 1.5])


def generate_correlated_design(n, d=5, seed=42, noise_std=2.0, rho1=0.8, rho2=0.9):
    """Return (A, b, x_true): A is n x d float64 C-order.

    Columns come in groups of five -- a correlated pair N([6, .2], .25[[1,r1],[r1,1]]),
    a correlated pair N([300, 60], 100[[1,r2],[r2,1]]) and one N(4, 1) column
    (easy_boston_data.py:25-34); the d mod 5 left-over columns are N(4, 1) with
    zero true coefficient.  b = A x_true + N(0, noise_std^2).
    """
    rng = np.random.default_rng(seed)
    groups, rest = divmod(int(d), 5)
    pair1 = 0.25 * np.array([[1.0, rho1], [rho1, 1.0]])
    pair2 = 100 * np.array([[1.0, rho2], [rho2, 1.0]])
    cols = []
    for _ in range(groups):
        cols.append(rng.multivariate_normal([6, 0.2], pair1, size=n))
        cols.append(rng.multivariate_normal([300, 60], pair2, size=n))
        cols.append(rng.normal(4, 1.0, size=(n, 1)))
    if rest:
        cols.append(rng.normal(4, 1.0, size=(n, rest)))
    A = np.hstack(cols)
    x_true = np.concatenate([np.tile(_GROUP_COEF, groups), np.zeros(rest)])
    b = A @ x_true + rng.normal(0, noise_std, size=n)
    return A, b, x_true


def generate_correlated_boston_like_data(m=1000, seed=42, noise_std=2.0, rho1=0.8, rho2=0.9):
    """Drop-in for easy_boston_data.generate_correlated_boston_like_data (same
    signature and defaults, easy_boston_data.py:7-13)."""
    return generate_correlated_design(m, 5, seed, noise_std, rho1, rho2)


def standardize(A, b):
    """z-score the columns of A and centre b -- what the (missing) notebook
    evidently did before calling the solvers."""
    mu = A.mean(axis=0)
    sd = A.std(axis=0)
    sd = np.where(sd > 0, sd, 1.0)
    return np.ascontiguousarray((A - mu) / sd), b - b.mean()


def lambda_max(A, b):
    """Smallest alpha1 for which the Lasso solution is identically zero."""
    return float(np.max(np.abs(A.T @ b)))
