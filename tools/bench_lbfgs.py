#!/usr/bin/env python
"""Config 4 (per-GPU shard): elastic-net L-BFGS (m=10) on fp32 storage, fg evaluations/s with the
scipy driver (the reference's) and with the device driver; achieved HBM GB/s per evaluation."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402
from fastoptsolver_b200.lbfgs import LBFGSSolver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=500_000)
ap.add_argument("--cols", type=int, default=8192)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--sharded", action="store_true", help="run under torchrun: --rows is the TOTAL row count")
args = ap.parse_args()
dt = np.float32 if args.dtype == "f32" else np.float64
rank, world = 0, 1
if args.sharded:
    import torch
    import torch.distributed as dist
    from fastoptsolver_b200 import multigpu
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    des = multigpu.sharded_synthetic(args.rows, args.cols, dist, device=local, dtype=dt, seed=0, noise_std=0.5,
                                     rho1=0.5, rho2=0.7)
else:
    des = DeviceDesign.synthetic(args.rows, args.cols, dt, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
lam = des.lambda_max()
nbytes = args.rows * args.cols * np.dtype(dt).itemsize + args.rows * 8
out = {"rows": args.rows, "cols": args.cols, "dtype": args.dtype, "bytes_per_fg": nbytes, "world": world}
for driver in ("scipy", "device", "scipy", "device"):
    sol = LBFGSSolver("elasticnet", 0.1 * lam, 0.1 * lam, max_iter=args.iters, driver=driver)
    if args.sharded:
        dist.barrier()
    t0 = time.perf_counter()
    sol.fit(des)
    wall = time.perf_counter() - t0
    info = S.last_run["lbfgs"]
    nfg = info["fg_calls"]
    out[driver] = {"iters": len(sol.history_), "fg_calls": nfg, "wall_s": wall, "fg_per_s": nfg / wall,
                   "GBps": nbytes * nfg / wall / 1e9, "final_obj": float(sol.history_[-1]),
                   "loop_ms": info.get("loop_ms")}
if rank == 0:
    print(json.dumps(out))
if args.sharded:
    dist.destroy_process_group()
