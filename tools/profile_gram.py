#!/usr/bin/env python
"""Gram build + a few batched path iterations (for ncu)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import gram as GM  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=500_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--lambdas", type=int, default=256)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
des = DeviceDesign.synthetic(args.rows, args.cols, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
gram = GM.GramDesign(des)
lam = des.lambda_max()
X, info = GM.fista_path(des, None, lam * np.logspace(0, -3, args.lambdas), max_iter=args.iters, L=1e6, gram=gram)
print("build ms", gram.build_ms, "iter ms", info["loop_ms"] / args.iters)
