"""The C ABI from plain C (examples/c_client.c): builds with gcc against include/fos.h and
libfos_b200.so; fails loudly without a GPU; on a GPU it reproduces the Python binding's numbers."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fastoptsolver_b200")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    from fastoptsolver_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    _lib.load()
    exe = str(tmp_path_factory.mktemp("c_client") / "c_client")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_client.c"), "-o", exe, "-L", PKG, "-lfos_b200",
                    f"-Wl,-rpath,{PKG}", "-lm"], check=True)
    return exe


@pytest.mark.skipif(_has_cuda(), reason="CPU-only check")
def test_c_client_fails_loudly_without_gpu(client):
    p = subprocess.run([client], capture_output=True, text=True)
    assert p.returncode == 3
    assert "no CUDA device visible" in p.stderr and "no CPU fallback" in p.stderr
    assert "abi 1, 0 CUDA device(s)" in p.stdout


@pytest.mark.gpu
def test_c_client_matches_python_binding(client):
    n, d, iters = 20000, 1024, 15
    p = subprocess.run([client, str(n), str(d), str(iters)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    obj_c = [float(v) for v in re.findall(r"^obj\[\d+\] (\S+)$", p.stdout, re.M)]
    L_c = float(re.search(r" L (\S+) after (\d+) power steps", p.stdout).group(1))
    final_c = float(re.search(r"^objective\(x\) (\S+)$", p.stdout, re.M).group(1))
    assert len(obj_c) == iters
    from fastoptsolver_b200 import _lib
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    des = DeviceDesign.synthetic(n, d, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
    v0 = np.sin(0.37 * np.arange(1, d + 1))
    nrm = 0.0
    for v in v0:                      # same left-to-right sum as the C loop
        nrm += v * v
    v0 = v0 / np.sqrt(nrm)
    L, _, _ = des.power_iter(v0, 100, 1e-6)
    assert abs(L - L_c) <= 1e-13 * L          # libm sin may differ in the last bit between C and numpy
    a1 = 0.1 * des.lambda_max()
    x, it, xh, oh, _, _ = S._run(des, scheme=_lib.SCHEME_NESTEROV, alpha1=a1, alpha2=0.0, obj_terms=1, delta=0.0,
                                 backtracking=False, eta=0.5, step0=1.0 / L_c, max_iter=iters, tol=0.0, tol_ratio=0.0,
                                 adaptive_restart=False, restart_threshold=1.0, want_history=True)
    np.testing.assert_allclose(obj_c, oh[:it], rtol=1e-13)
    assert abs(final_c - obj_c[-1]) <= 1e-12 * abs(final_c)
    des.close()
