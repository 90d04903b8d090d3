#!/usr/bin/env python
"""A/B on one box: one row per stage (256,16,1) vs two rows per stage (256,16,2, FOS_ROWS_X2=1) for the
gradient + residual-recurrence pass at 1M x 4096; alternating solves on two designs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import _lib, iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

designs = {}
for x2 in ("0", "1"):
    os.environ["FOS_ROWS_X2"] = x2
    designs[x2] = DeviceDesign.synthetic(1_000_000, 4096, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
a1 = 0.1 * designs["0"].lambda_max()
v = np.random.default_rng(0).standard_normal(4096)
os.environ["FOS_ROWS_X2"] = "0"
L, _, _ = designs["0"].power_iter(v / np.linalg.norm(v), 30, 0.0)


def run(x2, K=40):
    os.environ["FOS_ROWS_X2"] = x2
    x, it, xh, oh, _, _ = S._run(designs[x2], scheme=_lib.SCHEME_NESTEROV, alpha1=a1, alpha2=0.0, obj_terms=1, delta=0.0,
                                 backtracking=False, eta=0.5, step0=1.0 / L, max_iter=K, tol=0.0, tol_ratio=0.0,
                                 adaptive_restart=False, restart_threshold=1.0, want_history=True)
    i = S.last_run["solver"]
    return i["loop_ms"] / i["passes"], float(oh[it - 1])


run("0", 10), run("1", 10)
for rep in range(4):
    a, b, c, d = run("0"), run("1"), run("0"), run("1")
    print("ms/pass  1 row/stage %.4f %.4f | 2 rows/stage %.4f %.4f   objective equal: %s" % (a[0], c[0], b[0], d[0], a[1] == b[1]),
          flush=True)
