import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from fastoptsolver_b200 import _lib, iterative_solvers as S
from fastoptsolver_b200.design import DeviceDesign
des = DeviceDesign.synthetic(1_000_000, 4096, np.float64, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
a1 = 0.1 * des.lambda_max()
v = np.random.default_rng(0).standard_normal(4096)
L, _, _ = des.power_iter(v / np.linalg.norm(v), 30, 0.0)
def run(hist, K=40):
    S._run(des, scheme=_lib.SCHEME_NESTEROV, alpha1=a1, alpha2=0.0, obj_terms=1, delta=0.0, backtracking=False, eta=0.5,
           step0=1.0 / L, max_iter=K, tol=0.0, tol_ratio=0.0, adaptive_restart=False, restart_threshold=1.0, want_history=hist)
    i = S.last_run["solver"]
    return i["loop_ms"] / i["passes"], i["passes"]
run(True, 10)
for rep in range(4):
    a = run(True); b = run(False); c = run(True); d = run(False)
    print("ms/pass with 2nd dot %.4f %.4f | gradient only %.4f %.4f" % (a[0], c[0], b[0], d[0]), flush=True)
