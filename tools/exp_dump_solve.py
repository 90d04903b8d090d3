#!/usr/bin/env python
"""Run one FISTA solve (fixed step and Armijo, with history) plus a power iteration on a synthetic
design and dump every output to an .npz -- for A/B comparisons of library builds / switches
(FOS_FUSED, FOS_LIB_PATH, ...) that must leave the results bit-identical.

    FOS_FUSED=0 python tools/exp_dump_solve.py out0.npz ; FOS_FUSED=1 python tools/exp_dump_solve.py out1.npz
    python tools/exp_dump_solve.py --compare out0.npz out1.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    if sys.argv[1] == "--compare":
        a, b = (dict(np.load(p)) for p in sys.argv[2:4])
        bad = [k for k in a if a[k].tobytes() != b[k].tobytes()]
        print("keys:", len(a), "differing:", bad or "none (bit-identical)")
        sys.exit(1 if bad else 0)
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    out = {}
    for n, d, dt in ((60000, 4096, np.float64), (50000, 2048, np.float64), (30011, 1000, np.float64),
                     (40000, 4096, np.float32), (20000, 8192, np.float32), (9001, 640, np.float64)):
        des = DeviceDesign.synthetic(n, d, dt, seed=1, noise_std=0.5, rho1=0.5, rho2=0.7)
        lam = des.lambda_max()
        tag = f"{n}x{d}{np.dtype(dt).char}"
        for name, kw in (("fixed", {}), ("armijo", dict(backtracking=True, t_init_factor=2.0))):
            np.random.seed(0)
            x, h = S.fista(des, None, "elasticnet", 0.1 * lam, 0.01 * lam, max_iter=12, return_history=True, **kw)
            out[f"{tag}/{name}/x"] = x
            out[f"{tag}/{name}/hx"] = np.array(h["x"])
            out[f"{tag}/{name}/obj"] = np.array(h["obj"])
        np.random.seed(0)
        out[f"{tag}/nohist"] = S.fista(des, None, "lasso", 0.1 * lam, 0.0, max_iter=7)
        loss, g = des.grad(np.linspace(-1, 1, d), 0.5)
        out[f"{tag}/grad"] = np.concatenate([[loss], g])
        des.close()
    np.savez(sys.argv[1], **out)
    print("wrote", sys.argv[1], len(out), "arrays")


if __name__ == "__main__":
    main()
