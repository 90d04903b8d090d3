#!/usr/bin/env python
"""Per-CTA duration spread of one gradient-kernel launch (static row partition)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import _lib  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--mode", type=int, default=3)
args = ap.parse_args()
des = DeviceDesign.synthetic(args.rows, args.cols, seed=0)
lib = _lib.load()
by_sm = []
for rep in range(4):
    buf = (C.c_longlong * 2048)()
    n = C.c_int()
    _lib.check(lib.fos_debug_cta_times(des.handle, args.mode, buf, 2048, C.byref(n)))
    raw = np.array(buf[: 2 * n.value], dtype=np.int64).reshape(-1, 2)
    smid = raw[:, 0] >> 40
    start = (raw[:, 0] & ((1 << 40) - 1)) * 1e-3
    end = raw[:, 1] * 1e-3
    dur = end - start
    print(f"rep {rep}: CTAs {n.value} distinct SMs {len(set(smid.tolist()))} start spread {start.max():.0f} us | end min/median/max "
          f"{end.min():.0f}/{np.median(end):.0f}/{end.max():.0f} us | idle tail share {(end.max() - end.mean()) / end.max() * 100:.2f}%"
          f" | blockIdx==smid for {int(np.sum(smid == np.arange(n.value)))} CTAs")
    d_sm = np.full(256, np.nan)
    d_sm[smid] = dur
    by_sm.append(d_sm)
ok = ~np.isnan(by_sm[0])
for i in range(1, len(by_sm)):
    a, b = by_sm[0][ok], by_sm[i][ok]
    print(f"corr(duration by SM id, rep 0 vs rep {i}) = {np.corrcoef(a, b)[0, 1]:.3f}")
order = np.argsort(np.nan_to_num(by_sm[0], nan=0))
print("slowest SMs:", order[-10:].tolist(), "fastest SMs:", [int(s) for s in np.argsort(np.nan_to_num(by_sm[0], nan=1e18))[:10]])
