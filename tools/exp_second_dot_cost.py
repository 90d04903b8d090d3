#!/usr/bin/env python
"""Alternating A/B on one box: ms per pass of a fixed-step FISTA solve with the recorded objective coming
from (a) the residual recurrence (GM_QREC), (b) a second dot product per pass (FOS_QREC=0), (c) no objective at
all (history off: gradient-only passes).

    python tools/exp_second_dot_cost.py [--rows 1000000] [--cols 4096] [--iters 40]
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
ap.add_argument("--iters", type=int, default=40)
a = ap.parse_args()
des = DeviceDesign.synthetic(a.rows, a.cols, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
a1 = 0.1 * des.lambda_max()
v = np.random.default_rng(0).standard_normal(a.cols)
L, _, _ = des.power_iter(v / np.linalg.norm(v), 30, 0.0)


def run(kind, K):
    os.environ["FOS_QREC"] = "0" if kind == "dot2" else "1"
    S._run(des, scheme=_lib.SCHEME_NESTEROV, alpha1=a1, alpha2=0.0, obj_terms=1, delta=0.0, backtracking=False, eta=0.5,
           step0=1.0 / L, max_iter=K, tol=0.0, tol_ratio=0.0, adaptive_restart=False, restart_threshold=1.0,
           want_history=(kind != "none"))
    i = S.last_run["solver"]
    return i["loop_ms"] / i["passes"]


run("qrec", 10)
for rep in range(4):
    r = {k: [] for k in ("qrec", "dot2", "none")}
    for _ in range(2):
        for k in r:
            r[k].append(run(k, a.iters))
    print("ms/pass  recurrence %.4f %.4f | second dot %.4f %.4f | no objective %.4f %.4f" %
          (*r["qrec"], *r["dot2"], *r["none"]), flush=True)
