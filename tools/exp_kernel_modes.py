#!/usr/bin/env python
"""Sustained vs burst duration of the gradient kernel in each mode (and of the streaming
probe) on one design, with nvidia-smi clocks/power sampled during each run.

    python tools/exp_kernel_modes.py [--rows 1000000] [--cols 4096] [--dtype f64]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import _lib  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

Q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"


class Smi:
    def __init__(self):
        self.rows = []

    def __enter__(self):
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={Q}", "--format=csv,noheader,nounits",
                                   "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=lambda: [self.rows.append(l.strip()) for l in self.p.stdout], daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        time.sleep(0.05)
        self.p.terminate()

    def summary(self):
        sm, mem, pw, tmp, cap = [], [], [], [], 0
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mem.append(float(f[1])); pw.append(float(f[2])); tmp.append(float(f[3]))
            except Exception:
                continue
            cap += f[4].lower().startswith("active")
        if not sm:
            return {}
        return {"n": len(sm), "sm_med": float(np.median(sm)), "sm_min": min(sm), "mem_med": float(np.median(mem)),
                "mem_min": min(mem), "power_max": max(pw), "power_med": float(np.median(pw)), "temp_max": max(tmp),
                "power_cap_samples": cap}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--cols", type=int, default=4096)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=60)
    ap.add_argument("--modes", default="probe,grad,grad+dot2,dot2")
    args = ap.parse_args()
    dt = np.float64 if args.dtype == "f64" else np.float32
    des = DeviceDesign.synthetic(args.rows, args.cols, dt, seed=0)
    lib = _lib.load()
    nbytes = args.rows * args.cols * np.dtype(dt).itemsize + args.rows * 8
    out = {"rows": args.rows, "cols": args.cols, "dtype": args.dtype, "bytes": nbytes, "runs": []}
    for name, mode in (("probe", 8), ("grad", 1), ("grad+dot2", 3), ("dot2", 2)):
        if name not in args.modes.split(","):
            continue
        for reps in (2, args.reps):
            ms = C.c_float()
            time.sleep(1.0)
            with Smi() as s:
                _lib.check(lib.fos_time_grad_kernel(des.handle, mode, reps, C.byref(ms)))
            rec = {"mode": name, "reps": reps, "ms": ms.value, "GBps": nbytes / ms.value / 1e6, **s.summary()}
            out["runs"].append(rec)
            print(json.dumps(rec), flush=True)
    des.close()


if __name__ == "__main__":
    main()
