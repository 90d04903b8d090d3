#!/usr/bin/env python
"""Host->device bandwidth of the design upload: pinned source, pageable source through the
driver's own staging (FOS_UPLOAD_STAGED=0) and through the threaded pinned-staging copy."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

os.environ["FOS_UPLOAD_GRAM"] = os.environ.get("FOS_UPLOAD_GRAM", "0")
n, d = 250000, 4096          # 8.2 GB
gb = n * d * 8 / 1e9
b = np.zeros(n)
A_pin = torch.empty((n, d), dtype=torch.float64, pin_memory=True).numpy()
A_pin[:] = 1.0
A_page = np.ones((n, d))


def upload(A, tag, **env):
    for k, v in env.items():
        os.environ[k] = v
    for rep in range(2):
        t0 = time.perf_counter()
        des = DeviceDesign.from_host(A, b)
        dt = time.perf_counter() - t0
        print(f"{tag} rep{rep}: {gb / dt:.1f} GB/s ({dt * 1e3:.0f} ms, copy {des.upload_gram()['copy_ms']:.0f} ms)", flush=True)
        des.close()
    for k in env:
        os.environ.pop(k)


upload(A_pin, "pinned source")
upload(A_page, "pageable, driver staging", FOS_UPLOAD_STAGED="0")
for t in ("2", "4", "8"):
    upload(A_page, f"pageable, {t} staging threads", FOS_UPLOAD_THREADS=t)
os.environ["FOS_UPLOAD_GRAM"] = "1"
upload(A_page, "pageable, 8 threads + Gram under the copy")
upload(A_pin, "pinned + Gram under the copy")
