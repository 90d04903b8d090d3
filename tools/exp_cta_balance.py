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
for rep in range(3):
    buf = (C.c_longlong * 2048)()
    n = C.c_int()
    _lib.check(lib.fos_debug_cta_times(des.handle, args.mode, buf, 2048, C.byref(n)))
    t = np.array(buf[: 2 * n.value], dtype=np.float64).reshape(-1, 2) * 1e-3  # us
    start, end = t[:, 0], t[:, 1]
    dur = end - start
    print(f"rep {rep}: CTAs {n.value} start spread {start.max():.0f} us | end min/median/max {end.min():.0f}/{np.median(end):.0f}/{end.max():.0f} us"
          f" | dur min/median/max {dur.min():.0f}/{np.median(dur):.0f}/{dur.max():.0f} us | idle tail share {(end.max() - end.mean()) / end.max() * 100:.2f}%")
    order = np.argsort(end)
    print("   slowest CTAs:", order[-6:].tolist(), "fastest:", order[:6].tolist())
