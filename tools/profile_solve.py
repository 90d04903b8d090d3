#!/usr/bin/env python
"""One short FISTA solve on the bench design through the persistent solve kernel (for ncu).

    ncu --set full -k regex:solve_stream -s 1 -c 1 python tools/profile_solve.py --iters 3
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import _lib, iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--solves", type=int, default=2)
a = ap.parse_args()
des = DeviceDesign.synthetic(a.rows, a.cols, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
alpha1 = 0.1 * des.lambda_max()
v = np.random.default_rng(0).standard_normal(a.cols)
L, _, _ = des.power_iter(v / np.linalg.norm(v), 5, 0.0)      # a few steps are enough for a usable step size
for i in range(a.solves):
    x, it, xh, oh, _, _ = S._run(des, scheme=_lib.SCHEME_NESTEROV, alpha1=alpha1, alpha2=0.0, obj_terms=1, delta=0.0,
                                 backtracking=False, eta=0.5, step0=1.0 / (2.0 * L), max_iter=a.iters, tol=0.0, tol_ratio=0.0,
                                 adaptive_restart=False, restart_threshold=1.0, want_history=True)
    info = S.last_run["solver"]
    print(f"solve {i}: {it} iterations, {info['passes']} passes in {info['kernel_launches']} launch(es), "
          f"{info['loop_ms']:.3f} ms, objective {oh[it - 1]:.6e}")
des.close()
