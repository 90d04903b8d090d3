#!/usr/bin/env python
"""bench.py -- FISTA iterations/s + achieved HBM GB/s, Lasso 1M x 4096 fp64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one FISTA-Lasso iteration with history on (objective recorded every iteration):
one fused pass over the row-sharded A plus the d-length epilogue.  `value` is measured
with A resident in HBM (CUDA events on the solver stream, max over ranks); `e2e` is the
same metric through the drop-in `fista(A, b, ...)` call on HOST numpy buffers, with the
upload of A, the Lipschitz power iteration and the history download inside the timed
region; `cpu_baseline` / `--impl reference` time the numpy oracle (the reference's
algorithm, reference numpy/OpenBLAS code path) on a row sample of the same design.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, N_COLS = 1_000_000, 4096
SCENARIO = dict(seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)   # s0_n0.5_r10.5_r20.7 of the figure grid
ALPHA_FRAC = 0.1                                            # alpha1 = 0.1 * lambda_max
METRIC = "fista_lasso_iters_per_s"
UNIT = "it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="override (testing only)")
    ap.add_argument("--cols", type=int, default=N_COLS, help="override (testing only)")
    ap.add_argument("--cpu-sample-rows", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, flag in zip(names, f[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(A, b, alpha1, steps, scale):
    """Time the oracle (numpy restatement of the reference's fista, same OpenBLAS code path
    the reference takes) on a row sample; it/s scaled to the full row count."""
    import oracle
    np.random.seed(0)
    t0 = time.perf_counter()
    x, hist = oracle.fista(A, b, "lasso", alpha1, 0.0, max_iter=steps, return_history=True)
    wall = time.perf_counter() - t0
    grad_t = float(np.sum(oracle.METRICS["grad_times"]))
    return {"call_it_s": steps / wall / scale, "wall_s": wall, "grad_it_s": steps / grad_t / scale,
            "x": x, "obj": hist["obj"]}


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def pick_sample_rows(n, d, requested):
    if requested:
        return min(n, requested)
    # ~4.3 GB of A: 260 passes (200 for the Lipschitz estimate + 3 per iteration x 20) are
    # 8-15 s of work for the box's 16 host cores (measured: 3.9-6.4 s at half this size)
    return int(min(n, max(1024, (4 << 30) // (8 * d))))


# ------------------------------------------------------------------------------- main
_JSON_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries that write to file descriptor 1 themselves
    (NCCL prints its version banner there) are sent to stderr for the rest of the run."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    args = parse()
    _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, d = args.rows, args.cols
    K, W = args.steps, max(args.warmup, 3)
    workload = f"lasso_fista_{n}x{d}_fp64_dense_rowsharded"
    config = {"workload": workload, "n": n, "d": d, "alpha1": f"{ALPHA_FRAC}*lambda_max", "history": True,
              "scenario": "s0_n0.5_r10.5_r20.7", "l2_policy": "inputs larger than L2 (A >> 126 MB, evict-first)",
              "parallelism": f"rows/{args.gpus}"}

    if args.impl == "reference":
        return reference_arm(args, world, rank, local_rank, config, K, W)

    from fastoptsolver_b200 import _lib, iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    import ctypes as C

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from fastoptsolver_b200 import multigpu
    device = local_rank
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world

    t_gen = time.perf_counter()
    if world > 1:
        des = multigpu.sharded_synthetic(n, d, dist, device=device, **SCENARIO)
    else:
        des = DeviceDesign.synthetic(n, d, np.float64, device=device, **SCENARIO)
    gen_s = time.perf_counter() - t_gen
    lam = des.lambda_max()        # one fused pass; global across ranks (exchange inside the kernel)
    alpha1 = ALPHA_FRAC * lam

    # Lipschitz estimate exactly as fista() does it (<= 100 passes), timed separately
    np.random.seed(0)
    t_lip = time.perf_counter()
    L = S.estimate_lipschitz(des)
    lip = dict(S.last_run["lipschitz"])
    lip["host_ms"] = (time.perf_counter() - t_lip) * 1e3

    lib = _lib.load()

    def solve(iters, profile):
        lib.fos_design_set_profile(des.handle, 1 if profile else 0)
        S.reset_metrics()
        x, it, xh, oh, _, _ = S._run(
            des, scheme=_lib.SCHEME_NESTEROV, alpha1=alpha1, alpha2=0.0, obj_terms=1, delta=0.0,
            backtracking=False, eta=0.5, step0=1.0 / L, max_iter=iters, tol=0.0, tol_ratio=0.0,
            adaptive_restart=False, restart_threshold=1.0, want_history=True)
        return x, oh[:it], dict(S.last_run["solver"])

    import torch

    def barrier():
        if dist is not None:
            dist.barrier()
        # fos_prox_grad synchronises its own stream before returning; this covers torch's streams
        torch.cuda.synchronize(device)

    solve(W, False)                                   # warm-up steps (untimed)
    barrier()
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    x, obj, info = solve(K, False)                    # EXACTLY K timed steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # Same K steps once more with a CUDA-event pair around every gradient-kernel launch (on the
    # solver stream) for the roofline.  Kept out of the region above because an event between two
    # launches switches off their programmatic-dependent-launch overlap.
    _, _, pinfo = solve(K, True)
    barrier()
    loop_ms = info["loop_ms"]
    per_rank = None
    if dist is not None:
        import torch
        t = torch.tensor([loop_ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        loop_ms = float(t.item())
        mine = torch.tensor([pinfo["grad_kernel_ms"] / max(pinfo["grad_kernel_launches"], 1),
                             info["epilogue_ms"] / K, info["exchange_ms"] / K, info["loop_ms"]],
                            device=f"cuda:{local_rank}", dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"grad_kernel_ms_avg": [float(v[0]) for v in allr], "epilogue_ms_per_step": [float(v[1]) for v in allr],
                    "exchange_wait_ms_per_step": [float(v[2]) for v in allr], "loop_ms": [float(v[3]) for v in allr]}
    value = K / (loop_ms * 1e-3)

    # ---- roofline of the dominant kernel (the fused gradient pass)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650"
    lda = d + (d % 2)
    rows_local = hi - lo
    alg_bytes = rows_local * lda * 8 + rows_local * 8
    k_launch = max(pinfo["grad_kernel_launches"], 1)
    k_ms = pinfo["grad_kernel_ms"] / k_launch
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "grad_kernel_traffic.json"))).get(
            f"{rows_local}x{d}")
    except Exception:
        pass
    roofline = {"kernel": "grad_stream_kernel<double,256,16,1>", "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": k_ms, "launches_timed": k_launch,
                "kernel_share_of_step": pinfo["grad_kernel_ms"] / pinfo["loop_ms"] if pinfo["loop_ms"] else None,
                "ms_per_step_with_events": pinfo["loop_ms"] / K,
                "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": loop_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (Philox correlated-column design generated in HBM)",
        "config": config, "gpu_launches": int(info["kernel_launches"]),
        "passes_over_A": int(info["passes"]), "roofline": roofline,
        "hbm_gbs_whole_step": world * alg_bytes * info["passes"] / (loop_ms * 1e-3) / 1e9,
        "lipschitz": {"L": float(L), "power_iters": lip["iters"], "gpu_ms": lip["gpu_ms"], "host_ms": lip["host_ms"],
                      "via": lip.get("via")},
        "final_objective": float(obj[-1]) if len(obj) else None, "nnz": int(np.count_nonzero(x)),
        "gen_s": gen_s,
    }
    out["epilogue_ms_per_step"] = info["epilogue_ms"] / K
    out["exchange_wait_ms_per_step"] = info["exchange_ms"] / K
    if per_rank is not None:
        out["per_rank"] = per_rank
    if clocks is not None:
        out["clocks"] = clocks

    # ---- end-to-end through the public drop-in API on host buffers (rank-local shard)
    if not args.no_e2e:
        out["e2e"] = e2e_run(des, alpha1, K, n, d, dist, local_rank)
        if world == 1:
            # the same call on PAGEABLE arrays (what a numpy caller of the reference passes): the upload
            # then goes through the threaded pinned-staging copy instead of a direct DMA.  Extra
            # information only: a host too small for a second copy of A must not cost the line.
            try:
                pg = e2e_run(des, alpha1, K, n, d, dist, local_rank, pinned=False)
                out["e2e_pageable"] = {k: pg[k] for k in ("value", "unit", "wall_s", "upload_s", "h2d_GBps",
                                                           "lipschitz_via")}
            except (MemoryError, RuntimeError) as e:
                out["e2e_pageable"] = {"skipped": f"{type(e).__name__}: {e}"[:200]}

    # ---- CPU baseline on rank 0, N == 1 only
    if rank == 0 and world == 1 and not args.no_cpu:
        rows_s = pick_sample_rows(n, d, args.cpu_sample_rows)
        A_s, b_s = des.download(0, rows_s)
        steps_cpu = min(K, 20)
        res = cpu_reference_run(A_s, b_s, alpha1 * rows_s / n, steps_cpu, n / rows_s)
        out["cpu_baseline"] = {
            "value": res["call_it_s"], "unit": UNIT, "cores": host_threads(), "kind": "port",
            "sample": f"oracle.fista on rows [0,{rows_s}) of the same design ({rows_s}x{d} fp64, "
                      f"{rows_s * d * 8 / 1e9:.2f} GB), {steps_cpu} iterations incl. Lipschitz estimate, "
                      f"wall {res['wall_s']:.1f} s, it/s divided by {n / rows_s:.1f} (rows ratio)",
            "gradient_only_it_s": res["grad_it_s"],
        }
    if rank == 0:
        emit(out)
    if dist is not None:
        dist.barrier()
    des.close()
    if dist is not None:
        dist.destroy_process_group()


def e2e_run(des, alpha1, K, n, d, dist, local_rank, pinned=True):
    """fista(A_host, b_host, ...) through the public API: H2D of this rank's rows of A and b from
    pinned host memory, (multi-GPU: exchange-window wiring,) Lipschitz estimate, K iterations with
    history, D2H of the iterates -- all inside the timed region; wall clock, max over ranks."""
    import ctypes as C

    import torch
    from fastoptsolver_b200 import _lib, multigpu
    from fastoptsolver_b200 import design as D
    from fastoptsolver_b200 import iterative_solvers as S
    rows = des.shape[0]
    # stage the same numbers in pinned host memory (outside the timed region)
    if pinned:
        A_pin = torch.empty((rows, d), dtype=torch.float64, pin_memory=True)
        b_pin = torch.empty((rows,), dtype=torch.float64, pin_memory=True)
        A_h, b_h = A_pin.numpy(), b_pin.numpy()
    else:
        A_h, b_h = np.empty((rows, d)), np.empty(rows)
    _lib.check(_lib.load().fos_design_download(des.handle, 0, rows, C.c_void_p(A_h.ctypes.data),
                                               C.c_void_p(b_h.ctypes.data)))
    D.clear_cache()
    if dist is not None:
        dist.barrier()
    np.random.seed(0)
    t0 = time.perf_counter()
    upload_s = None
    if dist is not None:
        shard = multigpu.sharded_from_host(A_h, b_h, dist, device=local_rank)
        upload_s = time.perf_counter() - t0
        x, hist = S.fista(shard, None, "lasso", alpha1, 0.0, max_iter=K, return_history=True)
    else:
        shard = None
        D.as_design(A_h, b_h, device=local_rank)      # the upload fista() would do itself, timed apart
        upload_s = time.perf_counter() - t0
        x, hist = S.fista(A_h, b_h, "lasso", alpha1, 0.0, max_iter=K, return_history=True)
    wall = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([wall], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    info = dict(S.last_run["solver"])
    lip = dict(S.last_run["lipschitz"])
    ug = (shard if shard is not None else D.as_design(A_h, b_h, device=local_rank)).upload_gram()
    gram_info = {"state": ug["state"], "copy_ms": ug["copy_ms"], "tail_ms": ug["tail_ms"]}
    if shard is not None:
        shard.close()
    D.clear_cache()
    h2d = rows * d * 8 + rows * 8
    d2h = (K + 1) * d * 8 + K * 8
    return {"value": K / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
            "wall_s": wall, "upload_s": upload_s, "h2d_GBps": (h2d / upload_s / 1e9) if upload_s else None,
            "loop_ms": info["loop_ms"], "lipschitz_ms": lip["gpu_ms"],
            "lipschitz_iters": lip["iters"], "lipschitz_via": lip.get("via"),
            "upload_gram": gram_info, "host_s": dict(S.last_run.get("host_s", {})),
            "solve_host_ms": info.get("host_ms"), "bytes_are": "per rank",
            "what": "fista(A, b, 'lasso', a1, 0, max_iter=K, return_history=True) on pinned host numpy "
                    "arrays (each rank its row block): upload of A+b (with G = A^T A accumulated under the "
                    "copy on the tensor cores when lipschitz_via == 'gram'), <=100-step Lipschitz estimate, K "
                    "streaming iterations, history download"}


def reference_arm(args, world, rank, local_rank, config, K, W):
    """--impl reference: the reference's CPU implementation (oracle port; the Python reference
    tree itself is not on the GPU box) on the host cores, same metric/config."""
    if rank != 0:
        return
    n, d = args.rows, args.cols
    rows_s = pick_sample_rows(n, d, args.cpu_sample_rows)
    from fastoptsolver_b200.design import DeviceDesign
    des = DeviceDesign.synthetic(rows_s, d, np.float64, row0=0, device=local_rank, **SCENARIO)
    full = None
    A_s, b_s = des.download(0, rows_s)
    lam_s = float(np.max(np.abs(A_s.T @ b_s)))
    des.close()
    alpha1 = ALPHA_FRAC * lam_s
    steps = max(1, min(K, 20))
    cpu_reference_run(A_s[: max(64, rows_s // 16)], b_s[: max(64, rows_s // 16)], alpha1, min(W, 3), 1.0)  # warm-up
    res = cpu_reference_run(A_s, b_s, alpha1, steps, n / rows_s)
    val = res["call_it_s"]
    sample = (f"oracle.fista (numpy/OpenBLAS, the reference's code path) on a {rows_s}x{d} fp64 row sample "
              f"of the same synthetic design, {steps} iterations incl. Lipschitz estimate, wall "
              f"{res['wall_s']:.1f} s, it/s divided by {n / rows_s:.1f} (rows ratio)")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": steps,
           "warmup": W, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample,
                            "gradient_only_it_s": res["grad_it_s"]},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


if __name__ == "__main__":
    main()
