"""CPU: bench.py's reference arm (the one leg that runs without a GPU) -- it must not import the
product, must use every host core even under torchrun's OMP_NUM_THREADS=1, must report the loop
proper as `value` with the Lipschitz estimate apart, and must say that it ran on a scaled row sample."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra],
                         capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_contract():
    cores = len(os.sched_getaffinity(0))
    r = _run("--cpu-sample-rows", "4096", "--steps", "6", "--warmup", "3", env={"OMP_NUM_THREADS": "1"})
    assert r["impl"] == "reference" and r["metric"] == "fista_lasso_iters_per_s" and r["unit"] == "it/s"
    assert r["imports_product"] is False and r["gpu_launches"] == 0
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == cores and cb["value"] == r["value"]
    cfg = r["config"]
    assert cfg["workload"] == "lasso_fista_1000000x4096_fp64_dense_rowsharded"
    assert cfg["cpu_arm_sample_rows"] == 4096 and cfg["cpu_arm_scaled"] is True
    # value = loop proper, e2e = whole call including the Lipschitz estimate: e2e is the slower one
    assert r["e2e"]["value"] < r["value"] and r["e2e"]["h2d_bytes_per_step"] == 0
    assert cb["whole_call_it_s"] == r["e2e"]["value"] and cb["lipschitz_s"] > 0


def test_reference_arm_under_torchrun_other_ranks_stay_silent():
    e = {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "OMP_NUM_THREADS": "1"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, **e), cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--rows", "2000", "--cols", "640"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("a GPU is present")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_generator_model_statistics():
    """The numpy model of the device generator: unit-variance columns with the scenario's
    correlations inside each group of five, b = A x_true + noise."""
    sys.path.insert(0, ROOT)
    from oracle import datagen_model
    A, b = datagen_model.synth_rows(20000, 15, seed=3, noise_std=0.5, rho1=0.5, rho2=0.7)
    assert np.all(np.abs(A.mean(axis=0)) < 0.05) and np.all(np.abs(A.std(axis=0) - 1.0) < 0.03)
    c = np.corrcoef(A.T)
    for g in range(3):
        assert abs(c[5 * g, 5 * g + 1] - 0.5) < 0.03 and abs(c[5 * g + 2, 5 * g + 3] - 0.7) < 0.03
        assert abs(c[5 * g, 5 * g + 2]) < 0.03
    x_true = np.tile([5.0, 0.0, -0.02, -0.05, 1.5], 3)
    assert abs(np.std(b - A @ x_true) - 0.5) < 0.02
    # deterministic and independent of the thread count / block size
    A2, b2 = datagen_model.synth_rows(20000, 15, seed=3, noise_std=0.5, rho1=0.5, rho2=0.7, threads=1, block=777)
    assert np.array_equal(A, A2) and np.array_equal(b, b2)
