#!/usr/bin/env python
"""Streaming multi-RHS mode: passes/s, effective HBM GB/s and fp64 tensor TFLOP/s of the cluster kernel.

    python tools/bench_mrhs.py [--rows 500000] [--cols 4096] [--lambdas 8] [--iters 20]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import gram as GM  # noqa: E402
from fastoptsolver_b200 import iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=500_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--lambdas", type=int, default=8)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
des = DeviceDesign.synthetic(a.rows, a.cols, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
des.standardize()
lam = des.lambda_max()
alphas = lam * np.logspace(-0.3, -2.0, a.lambdas)
np.random.seed(0)
L = S.estimate_lipschitz(des)
for rep in range(3):
    X, info = GM.fista_path_stream(des, None, alphas, max_iter=a.iters, L=L)
    passes = (a.iters + 1) * info["batches"]
    ms = info["loop_ms"] / passes
    nbytes = a.rows * a.cols * 8
    print(json.dumps({"rows": a.rows, "cols": a.cols, "lambdas": a.lambdas, "iters": a.iters, "loop_ms": info["loop_ms"],
                      "ms_per_pass": ms, "GBps": nbytes / ms / 1e6, "dmma_TFLOPs": 4.0 * a.rows * a.cols * 8 / ms / 1e9,
                      "column_iterations_per_s": a.lambdas * a.iters / (info["loop_ms"] * 1e-3),
                      "nnz": [int(np.count_nonzero(x)) for x in X[:4]]}), flush=True)
des.close()
