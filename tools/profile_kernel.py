#!/usr/bin/env python
"""Launch the gradient kernel a few times in one mode on the bench design (for ncu).

    ncu --set full -k regex:grad_stream -s 1 -c 1 python tools/profile_kernel.py --mode 3
"""
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
ap.add_argument("--dtype", default="f64")
ap.add_argument("--mode", type=int, default=3)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
des = DeviceDesign.synthetic(args.rows, args.cols, np.float64 if args.dtype == "f64" else np.float32, seed=0,
                             noise_std=0.5, rho1=0.5, rho2=0.7)
ms = C.c_float()
_lib.check(_lib.load().fos_time_grad_kernel(des.handle, args.mode, args.reps, C.byref(ms)))
print(f"mode {args.mode}: {ms.value:.4f} ms/launch")
des.close()
