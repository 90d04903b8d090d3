#!/usr/bin/env python
"""Config 5: regularisation path, 256 lambdas batched on 500k x 4096 via Gram mode.
Reports the Gram build (fp64 DMMA SYRK) and the batched path iteration against the fp64
tensor peak, and the streaming alternative for comparison."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import gram as GM  # noqa: E402
from fastoptsolver_b200 import iterative_solvers as S  # noqa: E402
from fastoptsolver_b200.design import DeviceDesign  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=500_000)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--lambdas", type=int, default=256)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--sharded", action="store_true", help="run under torchrun: rows sharded for the build, lambdas for the path")
args = ap.parse_args()
n, d, Lm = args.rows, args.cols, args.lambdas
rank, world, dist = 0, 1, None
if args.sharded:
    import torch
    import torch.distributed as dist
    from fastoptsolver_b200 import multigpu
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    des = multigpu.sharded_synthetic(n, d, dist, device=local, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
    dist.barrier()
else:
    des = DeviceDesign.synthetic(n, d, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
t0 = time.perf_counter()
gram = GM.GramDesign(des)
if dist is not None:
    gram.allreduce(dist)
wall_build = time.perf_counter() - t0
lam = des.lambda_max()
alphas_all = lam * np.logspace(0, -3, Lm)
alphas = alphas_all[rank::world] if world > 1 else alphas_all      # each rank solves its share of the path
Lm_total, Lm = Lm, len(alphas)
np.random.seed(0)
L = S.estimate_lipschitz(des)
X, info = GM.fista_path(des, None, alphas, max_iter=args.iters, L=L, gram=gram)
X, info = GM.fista_path(des, None, alphas, max_iter=args.iters, L=L, gram=gram)
rows_local = des.shape[0]
syrk_flop = 2.0 * rows_local * d * d / 2 * (1 + 1.0 / (d // 128))   # upper tile triangle incl. diagonal tiles
it_flop = 2.0 * d * d * Lm    # algorithmic: the penalties asked for, not the padded tile width
out = {
    "n": n, "d": d, "lambdas": Lm_total, "lambdas_per_rank": Lm, "world": world, "nsplit": gram.nsplit,
    "gram_build_ms": gram.build_ms, "gram_build_wall_s": wall_build,
    "gram_tflops": syrk_flop / (gram.build_ms * 1e-3) / 1e12,
    "path_iter_ms": info["loop_ms"] / args.iters, "path_iter_tflops": it_flop / (info["loop_ms"] / args.iters * 1e-3) / 1e12,
    "path_lambda_iters_per_s": Lm_total * args.iters / (info["loop_ms"] * 1e-3),
    "nnz_first_last": [int(np.count_nonzero(X[0])), int(np.count_nonzero(X[-1]))],
    "fp64_tensor_peak_tflops_nominal": 40.0,
}
if rank == 0:
    print(json.dumps(out))
if dist is not None:
    dist.destroy_process_group()
