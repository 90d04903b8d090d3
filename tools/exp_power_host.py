#!/usr/bin/env python
"""Host wall clock vs device time of the streaming power iteration (estimate_lipschitz)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FOS_UPLOAD_GRAM"] = "0"
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

import torch  # noqa: E402


def run(des, tag):
    v = np.random.randn(des.shape[1])
    v /= np.linalg.norm(v)
    for rep in range(3):
        t0 = time.perf_counter()
        L, it, ms = des.power_iter(v, 100, 1e-6)
        w = time.perf_counter() - t0
        print(f"{tag} rep{rep}: host {w * 1e3:.1f} ms, device {ms:.1f} ms, iters {it}", flush=True)


des = DeviceDesign.synthetic(1000000, 4096, np.float64)
run(des, "synthetic 1Mx4096")
des.close()
n, d = 200000, 4096
A = torch.empty((n, d), dtype=torch.float64, pin_memory=True).numpy()
A[:] = np.random.default_rng(0).standard_normal((1000, d))[np.arange(n) % 1000]
b = np.random.default_rng(1).standard_normal(n)
t0 = time.perf_counter()
des = DeviceDesign.from_host(A, b)
print(f"upload {time.perf_counter() - t0:.3f} s")
run(des, "uploaded 200kx4096")
