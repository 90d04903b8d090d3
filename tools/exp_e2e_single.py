#!/usr/bin/env python
"""Where the wall time of fista(A_host, b_host, ...) goes on one GPU: five consecutive calls on the
same pinned host arrays with the library's upload / teardown timing printed (FOS_UPLOAD_DEBUG=1).

    FOS_UPLOAD_DEBUG=1 python tools/exp_e2e_single.py [--rows 1000000] [--cols 4096] [--iters 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fastoptsolver_b200 import iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--calls", type=int, default=5)
ap.add_argument("--keep-synthetic", action="store_true", help="keep the generator's design alive (as bench.py does)")
ap.add_argument("--pageable", action="store_true", help="plain numpy arrays (threaded pinned-staging upload)")
a = ap.parse_args()
des = DeviceDesign.synthetic(a.rows, a.cols, np.float64, **bench.SCENARIO)
alpha1 = 0.1 * des.lambda_max()
A_h, b_h = bench.host_copy_of(des, not a.pageable)
if not a.keep_synthetic:
    des.close()
for i in range(a.calls):
    np.random.seed(0)
    t0 = time.perf_counter()
    x, h = S.fista(A_h, b_h, "lasso", alpha1, 0.0, max_iter=a.iters, return_history=True)
    wall = time.perf_counter() - t0
    print(json.dumps({"call": i, "wall_s": round(wall, 4), "host_s": {k: round(v, 4) for k, v in S.last_run["host_s"].items()},
                      "upload_gram": S.last_run.get("upload_gram"), "loop_ms": S.last_run["solver"]["loop_ms"]}), flush=True)
