#!/usr/bin/env python
"""Config 5: regularisation path, 256 lambdas batched on 500k x 4096 via Gram mode.
Reports the Gram build (fp64 DMMA SYRK) and the batched path iteration against the fp64
tensor peak, and the streaming alternative for comparison."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import gram as GM  # noqa: E402
from fastoptsolver_b200 import iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=500_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--lambdas", type=int, default=256)
ap.add_argument("--iters", type=int, default=100)
args = ap.parse_args()
n, d, Lm = args.rows, args.cols, args.lambdas
des = DeviceDesign.synthetic(n, d, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
t0 = time.perf_counter()
gram = GM.GramDesign(des)
wall_build = time.perf_counter() - t0
lam = des.lambda_max()
alphas = lam * np.logspace(0, -3, Lm)
np.random.seed(0)
L = S.estimate_lipschitz(des)
X, info = GM.fista_path(des, None, alphas, max_iter=args.iters, L=L, gram=gram)
X, info = GM.fista_path(des, None, alphas, max_iter=args.iters, L=L, gram=gram)
syrk_flop = 2.0 * n * d * d / 2 * (1 + 1.0 / (d // 128))   # upper tile triangle incl. diagonal tiles
it_flop = 2.0 * d * d * ((Lm + 63) // 64 * 64)
out = {
    "n": n, "d": d, "lambdas": Lm, "nsplit": gram.nsplit,
    "gram_build_ms": gram.build_ms, "gram_build_wall_s": wall_build,
    "gram_tflops": syrk_flop / (gram.build_ms * 1e-3) / 1e12,
    "path_iter_ms": info["loop_ms"] / args.iters, "path_iter_tflops": it_flop / (info["loop_ms"] / args.iters * 1e-3) / 1e12,
    "path_lambda_iters_per_s": Lm * args.iters / (info["loop_ms"] * 1e-3),
    "nnz_first_last": [int(np.count_nonzero(X[0])), int(np.count_nonzero(X[-1]))],
    "fp64_tensor_peak_tflops_nominal": 40.0,
}
print(json.dumps(out))
