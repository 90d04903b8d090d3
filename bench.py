#!/usr/bin/env python
"""bench.py -- FISTA iterations/s + achieved HBM GB/s, Lasso 1M x 4096 fp64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c3|c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default (--config c3, BASELINE configs[2], the configuration the metric is quoted on): a step is
one FISTA-Lasso iteration with history on (objective recorded every iteration) = one fused pass
over the row-sharded A plus the d-length epilogue.  `value` is measured with A resident in HBM
(CUDA events on the solver stream, max over ranks); `e2e` is the same metric through the drop-in
`fista(A, b, ...)` call on HOST numpy buffers, with the upload of A, the Lipschitz power iteration
and the history download inside the timed region; `parity` runs one golden trace of the unmodified
reference (tests/golden) through the same row-sharded path at this N; `cpu_baseline` /
`--impl reference` time the numpy oracle (the reference's algorithm on the reference's
numpy/OpenBLAS code path) on a row sample of the same design, all host threads.

--config c2 | c4 | c5 measure BASELINE configs[1], [3] and [4] with the same JSON shape:
  c2  Lasso FISTA 100 000 x 2048 fp64 (one scenario of the figure grid; the 80-scenario sweep
      is fastoptsolver_b200/sweep.py)
  c4  elastic-net L-BFGS (m = 10), fp32 storage, 500 000 x 8192 rows PER GPU (4M x 8192 at N = 8);
      a step is one LBFGSSolver.fit; metric = loss+gradient evaluations / s
  c5  regularisation path: 256 penalties batched on 500 000 x 4096 through the Gram matrix;
      a step is one batched FISTA iteration; roofline = fp64 tensor pipe
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

SCENARIO = dict(seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)   # s0_n0.5_r10.5_r20.7 of the figure grid
ALPHA_FRAC = 0.1                                            # alpha1 = 0.1 * lambda_max
UNIT = "it/s"
SHAPES = {"c3": (1_000_000, 4096), "c2": (100_000, 2048), "c4": (500_000, 8192), "c5": (500_000, 4096)}
FP64_TENSOR_PEAK_TFLOPS = 40.0      # B200 nominal dense fp64 (DMMA); no measured figure in MEASURED_PEAKS.json
PARITY_CASE = ("wide", "fista/lasso-armijo-t2.0")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(SHAPES))
    ap.add_argument("--rows", type=int, default=0, help="override (testing only)")
    ap.add_argument("--cols", type=int, default=0, help="override (testing only)")
    ap.add_argument("--lambdas", type=int, default=256, help="c5: number of penalties")
    ap.add_argument("--cpu-sample-rows", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = {"c3": 100, "c2": 200, "c4": 20, "c5": 100}[a.config]
    n, d = SHAPES[a.config]
    a.rows = a.rows or n
    a.cols = a.cols or d
    return a


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled at 10 Hz DURING the timed regions.  The timed
    K steps alone can be shorter than one sample period, so the sampler stays on across the
    event-timed repeat and a soak of further identical steps until it has seen >= MIN_WINDOW_S of
    load (`covers` in the record says what ran underneath)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    MIN_WINDOW_S = 1.5

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = None
        self.nvml = None
        self._stop = False
        self.source = None

    def start(self):
        # in-process NVML at 10 Hz (four light queries per sample); the nvidia-smi -lms subprocess is the
        # fallback.  Either way the first sample is awaited before anything is timed (wait_first_sample).
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                     "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi"
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None
        self.t0 = time.perf_counter()

    def _poll_nvml(self):
        nv, h = self.nvml
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = int(get_reasons(h))
                self.rows.append(", ".join([str(sm), str(mx), f"{pw:.2f}"] +
                                           [("Active" if rs & b else "Not Active") for _, b in bits]))
            except Exception:
                pass
            time.sleep(0.1)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first_sample(self, timeout=3.0):
        """Return once nvidia-smi has delivered its first row: its start-up (NVML initialisation, a
        fresh process attaching to the GPU) takes 0.1-0.3 s and measurably slows kernels that run
        meanwhile (r2: 197 it/s in a timed region that overlapped it, 213-224 it/s for the identical
        solves that followed), so it must be over before the timed region begins."""
        t = time.perf_counter()
        while (self.proc is not None or self.nvml is not None) and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)
        self.t0 = time.perf_counter()

    def elapsed(self):
        return time.perf_counter() - self.t0 if self.t0 else 0.0

    def stop(self, covers=()):
        if self.proc is None and self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"]}
        window = self.elapsed()
        time.sleep(0.12)
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, pw = [], [], {}, []
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
                    reasons[nm] = reasons.get(nm, 0) + 1
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "power_w_median": float(np.median(pw)) if pw else None, "samples": len(sm),
                "window_s": window, "period_ms": 100, "source": self.source, "covers": list(covers), "reasons": sorted(reasons),
                "reason_samples": reasons}


# ------------------------------------------------------------------------------- host threads
def affinity_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; numpy's OpenBLAS has read it by
    now.  Raise the BLAS pool to every core this process may run on and report what was obtained."""
    want = affinity_cores()
    got = None
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want, user_api="blas")
        n = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        got = int(max(n)) if n else None
    except Exception:
        pass
    return got or 1, want


# ------------------------------------------------------------------------------- CPU arm
def cpu_fista_run(A, b, alpha1, steps, scale):
    """The oracle (numpy restatement of the reference's fista; every product goes through the same
    OpenBLAS dgemv calls the reference makes) on a row sample.  The Lipschitz estimate
    (iterative_solvers.py:45-60: up to 100 x 2 passes over A, more than the 20 iterations that
    follow) is timed apart, as BASELINE.md section 4 prescribes: `loop_it_s` is the loop proper --
    what the GPU arm's `value` measures -- `call_it_s` the whole call, what its `e2e` measures
    (less the upload, which the CPU does not have).  Rates are divided by `scale` = full rows /
    sample rows (every pass is linear in the row count)."""
    import oracle
    from oracle import ref_numpy
    lip = {"s": 0.0, "calls": 0}
    orig = ref_numpy.estimate_lipschitz

    def timed(*a, **k):
        t = time.perf_counter()
        out = orig(*a, **k)
        lip["s"] += time.perf_counter() - t
        lip["calls"] += 1
        return out

    ref_numpy.estimate_lipschitz = timed
    try:
        np.random.seed(0)
        t0 = time.perf_counter()
        x, hist = oracle.fista(A, b, "lasso", alpha1, 0.0, max_iter=steps, return_history=True)
        wall = time.perf_counter() - t0
    finally:
        ref_numpy.estimate_lipschitz = orig
    grad_t = float(np.sum(oracle.METRICS["grad_times"]))
    loop_s = wall - lip["s"]
    return {"loop_it_s": steps / loop_s / scale, "call_it_s": steps / wall / scale, "wall_s": wall,
            "lipschitz_s": lip["s"], "loop_s": loop_s, "grad_it_s": steps / grad_t / scale,
            "x": x, "obj": hist["obj"]}


def pick_sample_rows(n, d, requested, elem=8):
    if requested:
        return min(n, requested)
    # ~4.3 GB of A: ~260 passes (200 for the Lipschitz estimate + 3 per iteration x 20) are 8-15 s of
    # work for 16 host cores
    return int(min(n, max(1024, (4 << 30) // (elem * d))))


# ------------------------------------------------------------------------------- output
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


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def hbm_peak():
    peaks = measured_peaks()
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


class Ctx:
    """Process-group plumbing shared by the configurations."""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = self.local_rank
        self.dist = None

    def init_dist(self):
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        # the library synchronises its own stream before returning; this covers torch's streams
        torch.cuda.synchronize(self.device)

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([float(v)], device=f"cuda:{self.local_rank}", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(self, vals):
        import torch
        mine = torch.tensor([float(v) for v in vals], device=f"cuda:{self.local_rank}", dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allr, mine)
        return [[float(x) for x in v] for v in allr]

    def finish(self, des=None):
        if self.dist is not None:
            self.dist.barrier()
        if des is not None:
            des.close()
        if self.dist is not None:
            self.dist.destroy_process_group()


def soak_rounds(ctx, sampler, per_round_s, cap=400):
    """How many more identical rounds the clock sampler needs to reach MIN_WINDOW_S of load -- the same
    number on every rank (rank 0 owns the sampler; the others contribute 0 to the max)."""
    elapsed = ctx.max_over_ranks(sampler.elapsed() if ctx.rank == 0 else 0.0)
    per = max(ctx.max_over_ranks(per_round_s), 1e-4)
    return int(min(cap, np.ceil(max(0.0, ClockSampler.MIN_WINDOW_S - elapsed) / per)))


def pinned_host_pair(rows, d, dtype, pinned=True):
    import torch
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    if pinned:
        A = torch.empty((rows, d), dtype=tdt, pin_memory=True).numpy()
        b = torch.empty((rows,), dtype=torch.float64, pin_memory=True).numpy()
    else:
        A, b = np.empty((rows, d), dtype=dtype), np.empty(rows)
    return A, b


def host_copy_of(des, pinned=True):
    """This rank's rows of a device design in (pinned) host memory -- staged OUTSIDE the timed region."""
    import ctypes as C
    from fastoptsolver_b200 import _lib
    rows, d = des.shape
    A_h, b_h = pinned_host_pair(rows, d, des.dtype, pinned)
    _lib.check(_lib.load().fos_design_download(des.handle, 0, rows, C.c_void_p(A_h.ctypes.data),
                                               C.c_void_p(b_h.ctypes.data)))
    return A_h, b_h


def h2d_ceiling(ctx, A_h):
    """What the box gives a plain pinned -> HBM copy of the same host buffer, all ranks copying at the
    same moment (best of 2): the ceiling the upload inside the end-to-end call is measured against."""
    import torch
    src = torch.from_numpy(A_h)
    dev = torch.device("cuda", ctx.device)
    dst = torch.empty(src.shape, dtype=src.dtype, device=dev)
    best = None
    for _ in range(2):
        ctx.barrier()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    del dst
    torch.cuda.empty_cache()
    mine = src.numel() * src.element_size() / best / 1e9
    out = {"GBps_this_rank": mine, "what": "dst.copy_(pinned_src) of this rank's rows, all ranks at once, best of 2"}
    if ctx.dist is not None:
        per = [v[0] for v in ctx.gather_floats([mine])]
        out.update({"GBps_per_rank_min": min(per), "GBps_per_rank_max": max(per), "GBps_aggregate": sum(per)})
    return out


# =============================================================================== c3 / c2
def parity_record(ctx):
    """One golden trace of the UNMODIFIED reference (tests/golden/traces_wide.npz, written by
    tests/golden/make_golden.py from /root/reference) through the same row-sharded path the bench
    just timed, at this N: every iterate and the objective trace against the reference's, the Armijo
    shrink counts, and whether all ranks hold bit-identical iterates."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import multigpu
    from fastoptsolver_b200.design import DeviceDesign
    name, key = PARITY_CASE
    A, b = cases.design(name)
    spec = cases.solver_specs(name, A, b)[key]
    with np.load(os.path.join(cases.GOLDEN_DIR, f"traces_{name}.npz")) as z:
        ref = {k.rsplit("/", 1)[1]: z[k] for k in z.files if k.startswith(key + "/")}
    a1, a2 = (float(v) for v in ref["alpha"])
    lo, hi = multigpu.shard_bounds(A.shape[0], ctx.rank, ctx.world)
    if ctx.dist is not None:
        des = multigpu.sharded_from_host(np.ascontiguousarray(A[lo:hi]), b[lo:hi], ctx.dist, device=ctx.device)
    else:
        des = DeviceDesign.from_host(A, b, device=ctx.device)
    np.random.seed(spec["np_seed"])
    x, h = S.fista(des, None, spec["reg_type"], a1, a2, return_history=True, **spec["kw"])
    ls = list(S.ls_call_iters)
    nx = max(float(np.linalg.norm(ref["x"])), 1e-300)
    err_x = float(np.linalg.norm(x - ref["x"])) / nx
    err_hist = max(float(np.linalg.norm(xa - xb)) / nx for xa, xb in zip(h["x"], ref["hx"])) \
        if len(h["x"]) == ref["hx"].shape[0] else float("inf")
    err_obj = float(np.max(np.abs(np.asarray(h["obj"]) - ref["hobj"]) / np.abs(ref["hobj"]))) \
        if len(h["obj"]) == len(ref["hobj"]) else float("inf")
    ref_ls = [int(v) for v in ref["ls_iters"]]
    cut = next((i for i, v in enumerate(ref_ls) if v > 20), len(ref_ls))
    identical = True
    if ctx.dist is not None:
        import torch
        t = torch.from_numpy(np.concatenate([x] + [np.asarray(v) for v in h["x"]])).to(f"cuda:{ctx.local_rank}")
        lst = [torch.empty_like(t) for _ in range(ctx.world)]
        ctx.dist.all_gather(lst, t)
        identical = all(bool(torch.equal(lst[0], v)) for v in lst)
        ctx.dist.barrier()
    des.close()
    return {"case": f"{name}:{key}", "source": "tests/golden (unmodified reference, make_golden.py)",
            "shape": list(A.shape), "ranks": ctx.world, "iterations": len(h["obj"]),
            "rel_err_x": err_x, "rel_err_iterates_max": err_hist, "rel_err_obj": err_obj,
            "armijo_counts_equal": ls[:cut] == ref_ls[:cut] and len(ls) == len(ref_ls),
            "sign_pattern_equal": _same_pattern(x, ref["x"]),
            "ranks_bit_identical": identical, "tolerance": 1e-10,
            "ok": bool(err_x <= 1e-10 and err_hist <= 1e-10 and err_obj <= 1e-10 and identical)}


def _same_pattern(x, x_ref, rtol=1e-10):
    """Identical sign / sparsity pattern away from threshold ties (north_star)."""
    tie = 1e3 * rtol * max(float(np.abs(x_ref).max()), 1e-300)
    big = np.abs(x_ref) > tie
    return bool(np.array_equal(np.sign(x[big]), np.sign(x_ref[big])) and np.all(np.abs(x[x_ref == 0.0]) <= tie))


def bench_fista(ctx, cfg_name):
    args = ctx.args
    from fastoptsolver_b200 import _lib, iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    ctx.init_dist()
    world, rank, dist, device = ctx.world, ctx.rank, ctx.dist, ctx.device
    n, d = args.rows, args.cols
    K, W = args.steps, max(args.warmup, 3)
    rows_s = pick_sample_rows(n, d, args.cpu_sample_rows)
    config = fista_config(n, d, args.gpus, rows_s)
    if world > 1:
        from fastoptsolver_b200 import multigpu
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world

    t_gen = time.perf_counter()
    if world > 1:
        des = multigpu.sharded_synthetic(n, d, dist, device=device, **SCENARIO)
    else:
        des = DeviceDesign.synthetic(n, d, np.float64, device=device, **SCENARIO)
    gen_s = time.perf_counter() - t_gen
    lam = des.lambda_max()        # one fused pass; global across ranks (exchange inside the kernel)
    alpha1 = ALPHA_FRAC * lam

    # The clock sampler starts here, BEFORE the Lipschitz estimate: nvidia-smi's start-up idles the GPU
    # for 0.1-0.3 s, and an idle gap right before the W warm-up steps leaves the part in a low power
    # state that 25 ms of warm-up do not undo (r2: 198 it/s for K = 20 timed right after the gap,
    # 209 it/s for the identical K steps 0.1 s later).  The estimate is what fista() itself runs
    # before its loop (iterative_solvers.py:149 of the reference), so the order below is the order of
    # a real call: power iteration, then the iterations.
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    ctx.barrier()

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

    ctx.barrier()
    solve(W, False)                                   # warm-up steps (untimed)
    ctx.barrier()
    x, obj, info = solve(K, False)                    # EXACTLY K timed steps
    step_ms = np.asarray(S.grad_call_times, dtype=np.float64) * 1e3   # device time of every pass of the timed region
    ctx.barrier()
    # Same K steps once more with a CUDA-event pair around every gradient-kernel launch (on the
    # solver stream) for the roofline.  Kept out of the region above because an event between two
    # launches switches off their programmatic-dependent-launch overlap.
    _, _, pinfo = solve(K, True)
    ctx.barrier()
    # clock soak: the same steps until the sampler has seen >= 1.5 s of this load (all ranks take the
    # same number of rounds: the count is derived from the measured step time, not from a local clock)
    n_soak = soak_rounds(ctx, sampler, 1e-3 * info["loop_ms"] + 2e-3)
    soak_ms = []
    for _ in range(n_soak):
        soak_ms.append(solve(K, False)[2]["loop_ms"])
    ctx.barrier()
    clocks = sampler.stop(covers=["Lipschitz power iteration", "warm-up steps", "timed K steps", "event-timed repeat",
                                  f"{n_soak} x K soak steps"]) if rank == 0 else None
    loop_ms = ctx.max_over_ranks(info["loop_ms"])
    per_rank = None
    if dist is not None:
        g = ctx.gather_floats([pinfo["grad_kernel_ms"] / max(pinfo["grad_kernel_launches"], 1),
                               info["epilogue_ms"] / K, info["exchange_ms"] / K, info["loop_ms"]])
        per_rank = {"grad_kernel_ms_avg": [v[0] for v in g], "epilogue_ms_per_step": [v[1] for v in g],
                    "exchange_wait_ms_per_step": [v[2] for v in g], "loop_ms": [v[3] for v in g]}
    value = K / (loop_ms * 1e-3)

    # ---- roofline of the dominant kernel (the fused gradient pass)
    fused = int(info["kernel_launches"]) == 1      # the whole K-step solve was ONE launch of the persistent kernel
    peak, peak_src = hbm_peak()
    lda = d + (d % 2)
    rows_local = hi - lo
    alg_bytes = rows_local * lda * 8 + rows_local * 8
    if fused:
        # the dominant kernel IS the timed region: bytes per launch = passes x algorithmic bytes of a pass,
        # duration = the CUDA-event pair around the launch on the solver stream (the `value` measurement)
        k_launch = 1
        k_ms = info["loop_ms"]
        alg_bytes_launch = alg_bytes * int(info["passes"])
    else:
        k_launch = max(pinfo["grad_kernel_launches"], 1)
        k_ms = pinfo["grad_kernel_ms"] / k_launch
        alg_bytes_launch = alg_bytes
    achieved = alg_bytes_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "grad_kernel_traffic.json")))
        if fused:   # ncu measured one launch of 4 passes; per launch here = per-pass figure x this launch's passes
            per_pass = tj.get(f"solve:{rows_local}x{d}:per_pass")
            traffic = per_pass * int(info["passes"]) if per_pass else None
        else:
            traffic = tj.get(f"{rows_local}x{d}")
    except Exception:
        pass
    kname = {4096: "<double,256,16,1>", 2048: "<double,256,8,2>"}.get(d, "<double,...>")
    kname = ("solve_stream_kernel" if fused else "grad_stream_kernel") + kname
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "kernel_ms_avg": k_ms, "launches_timed": k_launch,
                "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None}
    if fused:
        passes = max(int(info["passes"]), 1)
        roofline.update({
            "passes_per_launch": passes,
            "what": "one persistent launch = every pass over A of the K timed steps + the in-kernel epilogues "
                    "(+ peer exchange); timed by the CUDA-event pair around the launch",
            # %globaltimer stamps taken inside the kernel: pass start -> all CTAs' partials published (the
            # streaming phase), and the rest of a pass (cross-CTA sums, exchange, prox, momentum, barriers)
            "gradient_phase_ms_avg": info["grad_kernel_ms"] / passes,
            "tail_ms_avg": info["epilogue_ms"] / passes,
            "gradient_phase_GBps": alg_bytes / (info["grad_kernel_ms"] / passes * 1e-3) / 1e9 if info["grad_kernel_ms"] else None})
    else:
        roofline.update({
            "kernel_share_of_step": pinfo["grad_kernel_ms"] / pinfo["loop_ms"] if pinfo["loop_ms"] else None,
            "ms_per_step_with_events": pinfo["loop_ms"] / K})

    out = {
        "metric": "fista_lasso_iters_per_s", "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
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
    if step_ms.size:
        out["pass_ms_in_timed_region"] = {"first": [round(float(v), 4) for v in step_ms[:3]],
                                          "median": float(np.median(step_ms)), "min": float(step_ms.min()),
                                          "max": float(step_ms.max()),
                                          "what": "%globaltimer span of every pass (gradient kernel start -> its epilogue) on rank 0"}
    out["epilogue_ms_per_step"] = info["epilogue_ms"] / K
    out["exchange_wait_ms_per_step"] = info["exchange_ms"] / K
    if soak_ms:
        out["soak"] = {"rounds": n_soak, "it_s_median": K / (float(np.median(soak_ms)) * 1e-3),
                       "it_s_min": K / (max(soak_ms) * 1e-3), "it_s_max": K / (min(soak_ms) * 1e-3),
                       "note": "rank-0 device times of further identical K-step solves run under the clock sampler"}
    if per_rank is not None:
        out["per_rank"] = per_rank
    if clocks is not None:
        out["clocks"] = clocks

    # ---- parity of the same (row-sharded) path against a golden trace of the reference, at this N
    if not args.no_parity:
        try:
            out["parity"] = parity_record(ctx)
        except Exception as e:      # a missing fixture must not cost the line; a mismatch is reported, not raised
            out["parity"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- end-to-end through the public drop-in API on host buffers (rank-local shard)
    if not args.no_e2e:
        try:
            out["e2e"] = e2e_fista(ctx, des, alpha1, K)
        except (MemoryError, RuntimeError) as e:     # e.g. no room for the pinned host copy: keep the line, say why
            out["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                          "error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and out["e2e"].get("value") is not None:
            # the same call on PAGEABLE arrays (what a numpy caller of the reference passes): the upload
            # then goes through the threaded pinned-staging copy instead of a direct DMA.  Extra
            # information only: a host too small for a second copy of A must not cost the line.
            try:
                pg = e2e_fista(ctx, des, alpha1, K, pinned=False, reps=1)
                out["e2e_pageable"] = {k: pg[k] for k in ("value", "unit", "wall_s", "upload_s", "h2d_GBps",
                                                           "lipschitz_via")}
            except (MemoryError, RuntimeError) as e:
                out["e2e_pageable"] = {"skipped": f"{type(e).__name__}: {e}"[:200]}

    # ---- CPU baseline on rank 0, N == 1 only
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            out["cpu_baseline"] = cpu_fista_record(des, alpha1, rows_s, n, d, K)
        except Exception as e:            # the CPU sample must not cost the GPU line
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                   "error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        emit(out)
    ctx.finish(des)


def cpu_fista_record(des, alpha1, rows_s, n, d, K):
    """cpu_baseline of the FISTA configs: the oracle on rows [0, rows_s) of the same design, all host threads."""
    cores, want = use_all_host_threads()
    A_s, b_s = des.download(0, rows_s)
    steps_cpu = min(K, 20)
    res = cpu_fista_run(A_s, b_s, alpha1 * rows_s / n, steps_cpu, n / rows_s)
    return {
        "value": res["loop_it_s"], "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"oracle.fista on rows [0,{rows_s}) of the same design ({rows_s}x{d} fp64, "
                  f"{rows_s * d * 8 / 1e9:.2f} GB), {steps_cpu} iterations with history; loop proper "
                  f"{res['loop_s']:.2f} s (value), Lipschitz estimate {res['lipschitz_s']:.2f} s apart, "
                  f"whole call {res['wall_s']:.2f} s; rates divided by {n / rows_s:.2f} (rows ratio)",
        "whole_call_it_s": res["call_it_s"], "gradient_only_it_s": res["grad_it_s"],
        "lipschitz_s": res["lipschitz_s"], "affinity_cores": want,
    }


def fista_config(n, d, gpus, rows_s):
    return {"workload": f"lasso_fista_{n}x{d}_fp64_dense_rowsharded", "n": n, "d": d,
            "alpha1": f"{ALPHA_FRAC}*lambda_max", "history": True, "scenario": "s0_n0.5_r10.5_r20.7",
            "l2_policy": "inputs larger than L2 (A >> 126 MB, evict-first)", "parallelism": f"rows/{gpus}",
            # the CPU arms (cpu_baseline / --impl reference) run on a row sample and scale by the row ratio
            "cpu_arm_sample_rows": rows_s, "cpu_arm_scaled": rows_s < n, "cpu_arm_scale": n / rows_s}


def e2e_fista(ctx, des, alpha1, K, pinned=True, reps=3):
    """fista(A_host, b_host, ...) through the public API: H2D of this rank's rows of A and b from
    pinned host memory, (multi-GPU: exchange-window wiring,) Lipschitz estimate, K iterations with
    history, D2H of the iterates, release of the device copy -- all inside the timed region; wall
    clock, max over ranks.  The call is made `reps` times (each one uploads again: the reference
    re-reads A on every call); `value` is K / the MEDIAN wall time, every wall time is listed."""
    from fastoptsolver_b200 import multigpu
    from fastoptsolver_b200 import iterative_solvers as S
    rows, d = des.shape
    dist = ctx.dist
    A_h, b_h = host_copy_of(des, pinned)       # staged outside the timed region
    runs = []
    for _ in range(reps):
        if dist is not None:
            dist.barrier()
        np.random.seed(0)
        t0 = time.perf_counter()
        if dist is not None:
            shard = multigpu.sharded_from_host(A_h, b_h, dist, device=ctx.device)
            upload_s = time.perf_counter() - t0
            x, hist = S.fista(shard, None, "lasso", alpha1, 0.0, max_iter=K, return_history=True)
            gram_info = shard.upload_gram()
        else:
            shard = None
            x, hist = S.fista(A_h, b_h, "lasso", alpha1, 0.0, max_iter=K, return_history=True)
            upload_s = S.last_run["host_s"]["design"]
            gram_info = S.last_run.get("upload_gram", {})
        wall = ctx.max_over_ranks(time.perf_counter() - t0)
        runs.append({"wall": wall, "upload_s": upload_s, "gram_info": gram_info, "info": dict(S.last_run["solver"]),
                     "lip": dict(S.last_run["lipschitz"]), "host_s": dict(S.last_run.get("host_s", {}))})
        if shard is not None:
            dist.barrier()
            shard.close()
    ceiling = None
    if pinned:
        try:
            ceiling = h2d_ceiling(ctx, A_h)
        except Exception as e:          # extra information only
            ceiling = {"error": f"{type(e).__name__}: {e}"[:200]}
    walls = [r["wall"] for r in runs]
    r = runs[int(np.argsort(walls)[len(walls) // 2])]          # the median run, reported in full
    wall, upload_s, gram_info, info, lip = r["wall"], r["upload_s"], r["gram_info"], r["info"], r["lip"]
    h2d = rows * d * 8 + rows * 8
    d2h = (K + 1) * d * 8 + K * 8
    return {"value": K / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
            "wall_s": wall, "walls_s": walls, "value_is": f"K / median wall time of {reps} calls",
            "upload_s": upload_s, "h2d_GBps": (h2d / upload_s / 1e9) if upload_s else None,
            "h2d_ceiling": ceiling,
            "upload_frac_of_ceiling": ((h2d / upload_s / 1e9) / ceiling["GBps_this_rank"]) if (
                upload_s and ceiling and ceiling.get("GBps_this_rank")) else None,
            "loop_ms": info["loop_ms"], "lipschitz_ms": lip["gpu_ms"],
            "lipschitz_iters": lip["iters"], "lipschitz_via": lip.get("via"),
            "upload_gram": {k: gram_info.get(k) for k in ("state", "copy_ms", "tail_ms")},
            "host_s": r["host_s"],
            "solve_host_ms": info.get("host_ms"), "bytes_are": "per rank",
            "what": "fista(A, b, 'lasso', a1, 0, max_iter=K, return_history=True) on pinned host numpy "
                    "arrays (each rank its row block): upload of A+b (with G = A^T A accumulated under the "
                    "copy on the tensor cores when lipschitz_via == 'gram'), <=100-step Lipschitz estimate, K "
                    "streaming iterations, history download, release of the device copy"}


# =============================================================================== c4
def bench_lbfgs(ctx):
    """Config 4: elastic-net L-BFGS (m = 10) on fp32 storage, 500 000 x 8192 rows per GPU (weak
    scaling: N = 8 is the 4M x 8192 design of BASELINE configs[3]), device driver.  A step is one
    LBFGSSolver.fit from x = 0 with the reference's defaults (max_iter 500, tol 1e-6); the metric
    counts loss+gradient evaluations (one fused pass over A each)."""
    args = ctx.args
    from fastoptsolver_b200 import _lib, iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    from fastoptsolver_b200.lbfgs import LBFGSSolver
    import ctypes as C
    ctx.init_dist()
    world, rank, dist, device = ctx.world, ctx.rank, ctx.dist, ctx.device
    rows_gpu, d = args.rows, args.cols
    n = rows_gpu * world
    K, W = args.steps, max(args.warmup, 3)
    if world > 1:
        from fastoptsolver_b200 import multigpu
        des = multigpu.sharded_synthetic(n, d, dist, device=device, dtype=np.float32, **SCENARIO)
    else:
        des = DeviceDesign.synthetic(n, d, np.float32, device=device, **SCENARIO)
    lam = des.lambda_max()
    a1 = a2 = ALPHA_FRAC * lam

    def fit():
        sol = LBFGSSolver("elasticnet", a1, a2, driver="device")
        sol.fit(des)
        return sol, dict(S.last_run["lbfgs"])

    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    ctx.barrier()
    for _ in range(W):
        fit()
    ctx.barrier()
    fg = 0
    loop_ms = 0.0
    iters = 0
    launches = 0
    for _ in range(K):                                # EXACTLY K timed steps (fits)
        sol, info = fit()
        fg += info["fg_calls"]
        iters += info["iters"]
        loop_ms += info["loop_ms"]
        launches += int(info["kernel_launches"])
    ctx.barrier()
    # the gradient kernel alone (gradient-only build, the mode every L-BFGS evaluation runs), event-timed
    ms = C.c_float()
    reps = 40
    _lib.check(_lib.load().fos_time_grad_kernel(des.handle, 1, reps, C.byref(ms)))
    soak = soak_rounds(ctx, sampler, 1e-3 * loop_ms / K + 2e-3, cap=200)
    for _ in range(soak):
        fit()
    ctx.barrier()
    clocks = sampler.stop(covers=["warm-up fits", "timed K fits", f"{reps} event-timed gradient launches", f"{soak} soak fits"]) \
        if rank == 0 else None
    loop_ms = ctx.max_over_ranks(loop_ms)
    value = fg / (loop_ms * 1e-3)
    peak, peak_src = hbm_peak()
    alg_bytes = rows_gpu * d * 4 + rows_gpu * 8
    achieved = alg_bytes / (ms.value * 1e-3) / 1e9
    out = {"metric": "lbfgs_fg_evals_per_s", "value": value, "unit": "fg/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": loop_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (Philox correlated-column design generated in HBM, float32 storage)",
           "config": {"workload": f"elasticnet_lbfgs_m10_{n}x{d}_fp32_rowsharded", "n": n, "d": d,
                      "rows_per_gpu": rows_gpu, "alpha1": f"{ALPHA_FRAC}*lambda_max", "alpha2": f"{ALPHA_FRAC}*lambda_max",
                      "driver": "device", "step": "one LBFGSSolver.fit (max_iter 500, tol 1e-6, m 10)",
                      "l2_policy": "inputs larger than L2 (16.4 GB per GPU, evict-first)", "parallelism": f"rows/{world}"},
           "gpu_launches": launches, "fg_evaluations": fg, "lbfgs_iterations": iters,
           "fg_per_fit": fg / K, "iters_per_s": iters / (loop_ms * 1e-3),
           "hbm_gbs_whole_step": world * alg_bytes * fg / (loop_ms * 1e-3) / 1e9,
           "final_objective": float(sol.history_[-1]),
           "roofline": {"kernel": "grad_stream_kernel<float,256,32,1,LITE>", "bound": "hbm", "achieved": achieved,
                        "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": ms.value, "launches_timed": reps,
                        "kernel_share_of_step": ms.value * fg / loop_ms,
                        "frac_of_nominal_8TBs": achieved / 8000.0,
                        "timed": "back-to-back launches of the gradient kernel alone between one CUDA-event pair"}}
    if clocks is not None:
        out["clocks"] = clocks
    if not args.no_e2e:
        out["e2e"] = e2e_lbfgs(ctx, des, a1, a2, K)
    if rank == 0 and world == 1 and not args.no_cpu:
        rows_s = pick_sample_rows(rows_gpu, d, args.cpu_sample_rows) // 8     # numpy up-casts A per product: keep it small
        A_s, b_s = des.download(0, rows_s)
        out["cpu_baseline"] = cpu_lbfgs_sample(A_s, b_s, a1 * rows_s / rows_gpu, a2 * rows_s / rows_gpu, rows_gpu / rows_s)
    if rank == 0:
        emit(out)
    ctx.finish(des)


def mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return None


def e2e_lbfgs(ctx, des, a1, a2, K):
    """LBFGSSolver.fit(A_host, b_host) on pinned host float32 arrays: upload, fit, read-back."""
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import multigpu
    from fastoptsolver_b200.lbfgs import LBFGSSolver
    rows, d = des.shape
    need = rows * d * 4 * ctx.world
    avail = mem_available_bytes()
    if avail is not None and need > 0.6 * avail:
        return {"value": None, "unit": "fg/s", "skipped": f"pinned host copies need {need / 1e9:.0f} GB, "
                f"MemAvailable is {avail / 1e9:.0f} GB", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}
    A_h, b_h = host_copy_of(des, True)
    dist = ctx.dist
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    fg = 0
    fits = min(K, 3)
    for _ in range(fits):
        sol = LBFGSSolver("elasticnet", a1, a2, driver="device")
        if dist is not None:
            shard = multigpu.sharded_from_host(A_h, b_h, dist, device=ctx.device)
            sol.fit(shard)
            dist.barrier()
            shard.close()
        else:
            sol.fit(A_h, b_h)
        fg += S.last_run["lbfgs"]["fg_calls"]
    wall = ctx.max_over_ranks(time.perf_counter() - t0)
    h2d = rows * d * 4 + rows * 8
    return {"value": fg / wall, "unit": "fg/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d * 8 + 8 * len(sol.history_),
            "wall_s": wall, "fits": fits, "fg_evaluations": fg, "bytes_are": "per rank, per fit",
            "what": "LBFGSSolver('elasticnet', a1, a2, driver='device').fit(A, b) on pinned host float32 arrays "
                    "(each rank its row block): upload of A+b, the fit, download of x and the objective history; "
                    "every fit uploads again (the reference re-reads A on every call)"}


def cpu_lbfgs_sample(A_s, b_s, a1_s, a2_s, scale):
    """oracle.LBFGSSolver on a row sample (penalties already scaled to the sample); rate divided by
    `scale` = rows of one GPU's share / sample rows."""
    import oracle
    cores, want = use_all_host_threads()
    rows_s, d = A_s.shape
    sol = oracle.LBFGSSolver("elasticnet", a1_s, a2_s)
    t0 = time.perf_counter()
    sol.fit(A_s, b_s)
    wall = time.perf_counter() - t0
    nfg = len(oracle.METRICS["grad_times"])
    return {"value": nfg / wall / scale, "unit": "fg/s", "cores": cores, "kind": "port", "affinity_cores": want,
            "sample": f"oracle.LBFGSSolver (scipy L-BFGS-B, the reference's driver) on rows [0,{rows_s}) of the design "
                      f"({rows_s}x{d} float32 storage), one fit = {nfg} evaluations + {len(sol.history_)} callback "
                      f"objectives, wall {wall:.1f} s; rate divided by {scale:.1f} (rows ratio, one GPU's share)"}


# =============================================================================== c5
def bench_path(ctx):
    """Config 5: the regularisation path, `--lambdas` penalties log-spaced from lambda_max down to
    1e-3 lambda_max, batched through the Gram matrix (fp64 DMMA).  A step is one batched FISTA
    iteration for all penalties; the Gram build is reported beside it.  N > 1: rows sharded for the
    build, penalties split across ranks for the iterations."""
    args = ctx.args
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    ctx.init_dist()
    world, rank, dist, device = ctx.world, ctx.rank, ctx.dist, ctx.device
    n, d, Lm_total = args.rows, args.cols, args.lambdas
    K, W = args.steps, max(args.warmup, 3)
    if world > 1:
        from fastoptsolver_b200 import multigpu
        des = multigpu.sharded_synthetic(n, d, dist, device=device, **SCENARIO)
    else:
        des = DeviceDesign.synthetic(n, d, np.float64, device=device, **SCENARIO)
    ctx.barrier()
    t0 = time.perf_counter()
    gram = GM.GramDesign(des)
    if dist is not None:
        gram.allreduce(dist)
    build_wall = ctx.max_over_ranks(time.perf_counter() - t0)
    build_ms = ctx.max_over_ranks(gram.build_ms)
    lam = des.lambda_max()
    alphas_all = lam * np.logspace(0, -3, Lm_total)
    alphas = np.ascontiguousarray(alphas_all[rank::world]) if world > 1 else alphas_all
    Lm = len(alphas)
    np.random.seed(0)
    L = S.estimate_lipschitz(des)
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    ctx.barrier()
    GM.fista_path(des, None, alphas, max_iter=W, L=L, gram=gram)        # warm-up steps
    ctx.barrier()
    X, info = GM.fista_path(des, None, alphas, max_iter=K, L=L, gram=gram)   # EXACTLY K timed steps
    ctx.barrier()
    soak = soak_rounds(ctx, sampler, 1e-3 * info["loop_ms"] + 5e-3, cap=200)
    for _ in range(soak):
        GM.fista_path(des, None, alphas, max_iter=K, L=L, gram=gram)
    ctx.barrier()
    clocks = sampler.stop(covers=["warm-up iterations", "timed K path iterations", f"{soak} x K soak iterations"]) if rank == 0 else None
    loop_ms = ctx.max_over_ranks(info["loop_ms"])
    rows_local = des.shape[0]
    tiles = d // 128 if d % 128 == 0 else (d + 127) // 128
    syrk_flop = 2.0 * rows_local * d * d / 2 * (1 + 1.0 / tiles)        # upper tile triangle incl. diagonal tiles
    it_flop = 2.0 * d * d * Lm    # algorithmic: the penalties asked for, not the padded tile width
    it_ms = loop_ms / K
    achieved = it_flop / (it_ms * 1e-3) / 1e12
    out = {"metric": "path_batched_fista_iters_per_s", "value": K / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": it_ms, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (Philox correlated-column design generated in HBM)",
           "config": {"workload": f"lasso_path_{Lm_total}lambdas_{n}x{d}_fp64_gram", "n": n, "d": d,
                      "lambdas": Lm_total, "lambdas_per_rank": Lm, "alphas": "lambda_max * logspace(0, -3)",
                      "l2_policy": "G (134 MB) + Y are re-read every iteration by design; between timed runs the "
                                   "warm-up leaves them in L2 as a running path would",
                      "parallelism": f"rows/{world} for the build, lambdas/{world} for the iterations"},
           "gpu_launches": int(info["launches"]), "lambda_iters_per_s": Lm_total * K / (loop_ms * 1e-3),
           "gram_build": {"ms": build_ms, "wall_s": build_wall, "tflops": syrk_flop / (build_ms * 1e-3) / 1e12,
                          "nsplit": gram.nsplit, "frac_of_nominal": syrk_flop / (build_ms * 1e-3) / 1e12 / FP64_TENSOR_PEAK_TFLOPS},
           "nnz_first_last": [int(np.count_nonzero(X[0])), int(np.count_nonzero(X[-1]))],
           "roofline": {"kernel": "path_step_sk_kernel / path_step_kernel (stream-K or tile schedule, TMA-staged operands "
                                  "unless FOS_PATH_TMA=0)", "bound": "tensor", "achieved": achieved,
                        "peak": FP64_TENSOR_PEAK_TFLOPS, "peak_source": "nominal B200 dense fp64 (DMMA) 40 TFLOP/s; "
                        "MEASURED_PEAKS.json holds no fp64 figure (ncu: 36 TFLOP/s with the pipe 97.6 % busy in the Gram build, "
                        "profiles/r2_gram_tma_syrk_path_sk_500kx4096_L256.txt)", "unit": "TFLOP/s", "frac": achieved / FP64_TENSOR_PEAK_TFLOPS,
                        "traffic": None, "algorithmic_flops_per_launch": it_flop, "kernel_ms_avg": it_ms,
                        "launches_timed": K}}
    if clocks is not None:
        out["clocks"] = clocks
    if not args.no_e2e:
        out["e2e"] = e2e_path(ctx, des, alphas, K)
    if rank == 0 and world == 1 and not args.no_cpu:
        rows_s = pick_sample_rows(n, d, args.cpu_sample_rows) // 2
        A_s, b_s = des.download(0, rows_s)
        out["cpu_baseline"] = cpu_path_sample(A_s, b_s, alphas_all * rows_s / n, n / rows_s, K)
    if rank == 0:
        emit(out)
    gram.close()
    ctx.finish(des)


def e2e_path(ctx, des, alphas, K):
    from fastoptsolver_b200 import gram as GM
    from fastoptsolver_b200 import multigpu
    rows, d = des.shape
    A_h, b_h = host_copy_of(des, True)
    dist = ctx.dist
    if dist is not None:
        dist.barrier()
    np.random.seed(0)
    t0 = time.perf_counter()
    if dist is not None:
        shard = multigpu.sharded_from_host(A_h, b_h, dist, device=ctx.device)
        gram = GM.GramDesign(shard)
        gram.allreduce(dist)
        X, info = GM.fista_path(shard, None, alphas, max_iter=K, gram=gram)
        gram.close()
        dist.barrier()
        shard.close()
    else:
        X, info = GM.fista_path(A_h, b_h, alphas, max_iter=K)
    wall = ctx.max_over_ranks(time.perf_counter() - t0)
    return {"value": K / wall, "unit": UNIT, "h2d_bytes_per_step": (rows * d * 8 + rows * 8) / K,
            "d2h_bytes_per_step": X.nbytes / K, "wall_s": wall, "loop_ms": info["loop_ms"], "bytes_are": "per rank",
            "what": "fista_path(A, b, alphas, max_iter=K) on pinned host arrays: upload (G accumulated under the copy "
                    "when eligible, else built after it), Lipschitz estimate, K batched iterations, download of X"}


def cpu_path_sample(A_s, b_s, alphas_s, scale, K):
    """The reference has no batched mode: its path is one fista call per penalty.  Two of the penalties
    (already scaled to the sample) are timed; a batched step covers all of them."""
    cores, want = use_all_host_threads()
    rows_s, d = A_s.shape
    picks = [len(alphas_s) // 4, len(alphas_s) - 1]
    steps = min(K, 20)
    rates = []
    for j in picks:
        res = cpu_fista_run(A_s, b_s, alphas_s[j], steps, scale)
        rates.append(res["loop_it_s"])
    per_lambda = float(np.mean(rates))
    return {"value": per_lambda / len(alphas_s), "unit": UNIT, "cores": cores, "kind": "port", "affinity_cores": want,
            "sample": f"oracle.fista (loop proper, Lipschitz estimate apart) for 2 of the {len(alphas_s)} penalties on rows "
                      f"[0,{rows_s}) ({rows_s}x{d} fp64), {steps} iterations each: {per_lambda:.3f} it/s per penalty after "
                      f"dividing by {scale:.1f} (rows ratio); a batched step covers all {len(alphas_s)} penalties, so "
                      f"value = that / {len(alphas_s)}"}


# =============================================================================== reference arm
def reference_arm(ctx):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the Python
    reference tree is not on the GPU box) on the host cores, same metric / config / unit.  Nothing
    of the product is imported: the row sample of the design comes from the numpy model of the
    device generator (oracle/datagen_model.py).  Under torchrun rank 0 alone runs."""
    args = ctx.args
    if ctx.rank != 0:
        return
    from oracle import datagen_model
    cores, want = use_all_host_threads()
    cfg = args.config
    n, d = args.rows, args.cols
    K, W = args.steps, max(args.warmup, 3)
    if cfg in ("c3", "c2"):
        rows_s = pick_sample_rows(n, d, args.cpu_sample_rows)
        t0 = time.perf_counter()
        A_s, b_s = datagen_model.synth_rows(rows_s, d, **SCENARIO)
        gen_s = time.perf_counter() - t0
        lam_s = float(np.max(np.abs(A_s.T @ b_s)))
        alpha1 = ALPHA_FRAC * lam_s             # = 0.1 lambda_max of the sample (lambda_max scales with the rows)
        steps = max(1, min(K, 100))
        m = max(64, rows_s // 16)
        cpu_fista_run(A_s[:m], b_s[:m], alpha1 * m / rows_s, min(W, 3), 1.0)     # warm-up
        res = cpu_fista_run(A_s, b_s, alpha1, steps, n / rows_s)
        val, e2e_val = res["loop_it_s"], res["call_it_s"]
        metric, unit, scaling = "fista_lasso_iters_per_s", UNIT, "strong"
        config = fista_config(n, d, args.gpus, rows_s)
        sample = (f"oracle.fista (numpy/OpenBLAS, the reference's code path) on a {rows_s}x{d} fp64 row sample of the same "
                  f"synthetic design (numpy model of the device generator, {gen_s:.1f} s), {steps} iterations with history: "
                  f"loop proper {res['loop_s']:.2f} s (value), Lipschitz estimate {res['lipschitz_s']:.2f} s, whole call "
                  f"{res['wall_s']:.2f} s (e2e); rates divided by {n / rows_s:.2f} (rows ratio)")
        extra = {"gradient_only_it_s": res["grad_it_s"], "whole_call_it_s": res["call_it_s"],
                 "lipschitz_s": res["lipschitz_s"]}
    elif cfg == "c4":
        rows_s = pick_sample_rows(n, d, args.cpu_sample_rows) // 8
        A_s, b_s = datagen_model.synth_rows(rows_s, d, dtype=np.float32, **SCENARIO)
        lam_s = float(np.max(np.abs(A_s.astype(np.float64).T @ b_s)))
        base = cpu_lbfgs_sample(A_s, b_s, ALPHA_FRAC * lam_s, ALPHA_FRAC * lam_s, n / rows_s)
        # n rows are ONE GPU's share; the N-GPU job has N times the rows (weak scaling)
        val = e2e_val = base["value"] / max(args.gpus, 1)
        steps = 1
        metric, unit, scaling, sample, extra = "lbfgs_fg_evals_per_s", "fg/s", "weak", base["sample"], {}
        config = {"workload": f"elasticnet_lbfgs_m10_{n * args.gpus}x{d}_fp32_rowsharded", "n": n * args.gpus, "d": d,
                  "rows_per_gpu": n, "cpu_arm_sample_rows": rows_s, "cpu_arm_scaled": True}
    else:
        rows_s = pick_sample_rows(n, d, args.cpu_sample_rows) // 2
        A_s, b_s = datagen_model.synth_rows(rows_s, d, **SCENARIO)
        lam_s = float(np.max(np.abs(A_s.T @ b_s)))
        base = cpu_path_sample(A_s, b_s, lam_s * np.logspace(0, -3, args.lambdas), n / rows_s, K)
        val = e2e_val = base["value"]
        steps = min(K, 20)
        metric, unit, scaling, sample, extra = "path_batched_fista_iters_per_s", UNIT, "strong", base["sample"], {}
        config = {"workload": f"lasso_path_{args.lambdas}lambdas_{n}x{d}_fp64_gram", "n": n, "d": d,
                  "lambdas": args.lambdas, "cpu_arm_sample_rows": rows_s, "cpu_arm_scaled": True}
    out = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": ctx.world, "steps": steps,
           "warmup": W, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": scaling,
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "cpu_baseline": dict({"value": val, "unit": unit, "cores": cores, "affinity_cores": want, "kind": "port",
                                 "sample": sample}, **extra),
           # value = the loop proper (what the GPU arm's `value` measures); e2e = the whole call including the
           # Lipschitz estimate (what the GPU arm's e2e includes), no copies on a CPU
           "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "includes": "estimate_lipschitz (<= 100 power iterations) + the iterations"},
           "gpu_launches": 0, "imports_product": "fastoptsolver_b200" in sys.modules}
    emit(out)


def main():
    args = parse()
    _claim_stdout()
    ctx = Ctx(args)
    if args.impl == "reference":
        return reference_arm(ctx)
    if args.config in ("c3", "c2"):
        return bench_fista(ctx, args.config)
    if args.config == "c4":
        return bench_lbfgs(ctx)
    return bench_path(ctx)


if __name__ == "__main__":
    main()
