#!/usr/bin/env python
"""Persistent solve kernel vs the two-launch path, iterate by iterate (debugging aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
K = int(sys.argv[3]) if len(sys.argv) > 3 else 200
out = {}
for fused in ("1", "0"):
    os.environ["FOS_FUSED"] = fused
    des = DeviceDesign.synthetic(n, d, seed=3, noise_std=0.5, rho1=0.5, rho2=0.7)
    a1 = 0.05 * des.lambda_max()
    np.random.seed(0)
    x, h = S.fista(des, None, "elasticnet", a1, 0.01 * a1, max_iter=K, return_history=True)
    info = dict(S.last_run["solver"])
    X = np.stack(h["x"])
    steps = np.linalg.norm(np.diff(X, axis=0), axis=1)
    out[fused] = (X, np.array(h["obj"]), steps, info)
    print("fused", fused, "launches", info["kernel_launches"], "iters", info["iters"], "loop_ms", info["loop_ms"])
    des.close()
Xa, oa, sa, _ = out["1"]
Xb, ob, sb, _ = out["0"]
m = min(len(sa), len(sb))
ra, rb = sa[1:m] / sa[:m - 1], sb[1:m] / sb[:m - 1]
dx = np.linalg.norm(Xa[:m + 1] - Xb[:m + 1], axis=1) / np.maximum(np.linalg.norm(Xb[:m + 1], axis=1), 1e-300)
print("max rel iterate diff", dx.max(), "at", int(dx.argmax()))
print("max rel obj diff", np.max(np.abs(oa[:m] - ob[:m]) / np.abs(ob[:m])))
print("min step ratio fused %.6f at %d, unfused %.6f at %d" % (ra.min(), ra.argmin() + 1, rb.min(), rb.argmin() + 1))
for k in range(max(0, int(ra.argmin()) - 3), min(m - 1, int(ra.argmin()) + 4)):
    print(k + 1, "ratio fused %.9f unfused %.9f  step %.6e %.6e  dx %.2e" % (ra[k], rb[k], sa[k + 1], sb[k + 1], dx[k + 1]))

# phase profile of the persistent kernel on a fresh design
import ctypes as C
from fastoptsolver_b200 import _lib
os.environ["FOS_FUSED"] = "1"
des = DeviceDesign.synthetic(n, d, seed=3, noise_std=0.5, rho1=0.5, rho2=0.7)
a1 = 0.05 * des.lambda_max()
np.random.seed(0)
L = S.estimate_lipschitz(des)
for rep in range(3):
    buf = (C.c_ulonglong * 9)()
    _lib.check(_lib.load().fos_debug_solve_profile(des.handle, buf, 1))
    np.random.seed(0)
    S.fista(des, None, "lasso", a1, 0.0, max_iter=K, return_history=True)
    _lib.check(_lib.load().fos_debug_solve_profile(des.handle, buf, 1))
    p = max(int(buf[8]), 1)
    names = ["stream", "bar1", "slice", "xchg", "elem1", "bar2", "decide+elem2", "bar3"]
    print("rep", rep, "loop_ms/pass %.4f" % (S.last_run["solver"]["loop_ms"] / p), " ".join(f"{nm}={buf[i] / p / 1e3:.2f}us" for i, nm in enumerate(names)))
des.close()
