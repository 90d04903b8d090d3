#!/usr/bin/env python
"""Phase profile of the persistent solve kernel (CTA 0's %globaltimer stamps, fos_debug_solve_profile): us per
pass spent in the streaming loop, at each of the three grid barriers, in the slice sums, the peer exchange, the
elementwise steps and the decision, for the three kinds of pass -- objective from the residual recurrence
(qrec), from a second dot product (dot2), gradient only (none).

    python tools/exp_solve_phases.py [rows=125000]        # FOS_BALANCE=1: rate-weighted row blocks
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import _lib, iterative_solvers as S
from fastoptsolver_b200.design import DeviceDesign
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
des = DeviceDesign.synthetic(rows, 4096, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
a1 = 0.1 * des.lambda_max()
v = np.random.default_rng(0).standard_normal(4096)
L, _, _ = des.power_iter(v / np.linalg.norm(v), 30, 0.0)
names = ["stream", "bar1", "slice", "xchg", "elem1", "bar2", "decide+elem2", "bar3"]
def run(kind, K=20):
    os.environ["FOS_QREC"] = "0" if kind == "dot2" else "1"
    buf = (C.c_ulonglong * 9)()
    _lib.check(_lib.load().fos_debug_solve_profile(des.handle, buf, 1))
    S._run(des, scheme=_lib.SCHEME_NESTEROV, alpha1=a1, alpha2=0.0, obj_terms=1, delta=0.0, backtracking=False, eta=0.5,
           step0=1.0 / L, max_iter=K, tol=0.0, tol_ratio=0.0, adaptive_restart=False, restart_threshold=1.0,
           want_history=(kind != "none"))
    _lib.check(_lib.load().fos_debug_solve_profile(des.handle, buf, 1))
    p = max(int(buf[8]), 1)
    i = S.last_run["solver"]
    return "%-5s %.4f ms/pass | " % (kind, i["loop_ms"] / i["passes"]) + " ".join(f"{nm}={buf[j] / p / 1e3:.1f}" for j, nm in enumerate(names))
run("qrec", 10)
for rep in range(3):
    for k in ("qrec", "dot2", "none"):
        print(run(k), flush=True)
