#!/usr/bin/env python
"""Multi-GPU end-to-end upload breakdown (run under torchrun): with a peer-attached design
already alive in every process (as in bench.py), time sharded_from_host of a pinned row block.
FOS_UPLOAD_DEBUG=1 prints the C-side stage timings."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastoptsolver_b200 import multigpu  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, d = 1000000, 4096
rows = n // world
main = multigpu.sharded_synthetic(n, d, dist, device=lr, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
A = torch.empty((rows, d), dtype=torch.float64, pin_memory=True).numpy()
A[:] = np.random.default_rng(rank).standard_normal((1000, d))[np.arange(rows) % 1000]
b = np.zeros(rows)
for rep in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh = multigpu.sharded_from_host(A, b, dist, device=lr)
    dt = time.perf_counter() - t0
    info = sh.upload_gram()
    if rank == 0:
        print(f"rep{rep}: sharded_from_host {dt * 1e3:.0f} ms, copy {info['copy_ms']:.0f} ms, tail {info['tail_ms']:.0f} ms, "
              f"state {info['state']}", flush=True)
    multigpu.close(sh, dist)
multigpu.close(main, dist)
dist.destroy_process_group()
