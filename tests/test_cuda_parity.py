"""GPU: the CUDA path (through the C ABI / the drop-in modules) against the golden traces
of the unmodified reference, and against the oracle on seeded inputs.

Tolerances are north_star's: fp64 iterates and objectives within 1e-10 relative after a
fixed iteration count, fp32 *storage* within 1e-5 (arithmetic is fp64 in both, so it is in
fact as tight as fp64), identical sign / sparsity pattern away from threshold ties.
L-BFGS traces follow SURVEY.md section 4: summation-order noise is amplified by the
quasi-Newton recursion, so the trace tolerance is looser on the long runs.
"""
import os

import numpy as np
import pytest

import cases
import harness

pytestmark = pytest.mark.gpu

RTOL_F64 = 1e-10


def _ops():
    with np.load(os.path.join(cases.GOLDEN_DIR, "operators.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def be():
    return harness.cuda_backend()


def test_library_sees_b200():
    from fastoptsolver_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    assert lib.fos_device_count() >= 1
    sm = C.c_int()
    tot = C.c_size_t()
    free = C.c_size_t()
    _lib.check(lib.fos_device_info(0, C.byref(sm), C.byref(tot), C.byref(free)))
    assert sm.value >= 100 and tot.value > 100e9


def test_prox_bitexact():
    from fastoptsolver_b200.operators import prox_elastic_net, prox_l1
    z = _ops()
    v = z["prox_v"]
    out = prox_l1(v, 0.75)
    np.testing.assert_array_equal(out, z["prox_l1_out"])
    np.testing.assert_array_equal(np.signbit(out), np.signbit(z["prox_l1_out"]))
    np.testing.assert_array_equal(prox_elastic_net(v, 0.5, 1.5, 0.25), z["prox_en_out"])
    np.testing.assert_array_equal(prox_l1(v[:12].reshape(3, 4), 0.3), z["prox_l1_2d"])
    assert prox_l1(np.zeros((0,)), 1.0).shape == (0,)


def test_objective_and_gradient():
    from fastoptsolver_b200.design import DeviceDesign
    from fastoptsolver_b200.operators import compute_objective
    z = _ops()
    A, b = cases.design("mid")
    x = z["obj_x"]
    for reg in ("lasso", "ridge", "elasticnet"):
        got = compute_objective(x, A, b, reg, 0.7, 0.3)
        assert abs(got - float(z[f"obj_{reg}"])) <= 1e-12 * abs(float(z[f"obj_{reg}"]))
    with pytest.raises(ValueError, match="Unsupported reg_type='bogus'"):
        compute_objective(x, A, b, "bogus", 0.7, 0.3)
    des = DeviceDesign.from_host(A, b)
    loss, g = des.grad(x)
    assert abs(loss - float(z["fg_loss"])) <= 1e-12 * float(z["fg_loss"])
    assert harness.rel_err(g, z["fg_grad"]) <= 1e-12
    # deterministic: bitwise identical on a rerun
    loss2, g2 = des.grad(x)
    assert loss2 == loss and np.array_equal(g, g2)


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_layouts_and_dtypes(order, dtype):
    """C / Fortran order, odd d (padded leading dimension), fp32 storage."""
    import oracle
    from fastoptsolver_b200.design import DeviceDesign
    rng = np.random.default_rng(3)
    for n, d in ((257, 37), (300, 640), (64, 1)):
        A = np.asarray(rng.standard_normal((n, d)), dtype=dtype, order=order)
        b = rng.standard_normal(n)
        x = rng.standard_normal(d)
        des = DeviceDesign.from_host(A, b)
        loss, g = des.grad(x, 0.25)
        loss_ref, g_ref = oracle.smooth_value_and_grad(x, A.astype(np.float64), b, 0.25)
        assert abs(loss - loss_ref) <= 1e-12 * abs(loss_ref)
        assert harness.rel_err(g, g_ref) <= 1e-12
        A_back, b_back = des.download()
        np.testing.assert_array_equal(A_back, A)
        np.testing.assert_array_equal(b_back, b)
        des.close()


def test_estimate_lipschitz(be):
    z = _ops()
    A, _ = cases.design("mid")
    for s in (0, 1, 7):
        np.random.seed(s)
        L = be.estimate_lipschitz(A)
        assert isinstance(L, np.float64)
        assert abs(L - float(z[f"lip_seed{s}"])) <= 1e-11 * L
    np.random.seed(0)
    assert abs(be.estimate_lipschitz(A, n_iter=5) - float(z["lip_n5"])) <= 1e-11 * float(z["lip_n5"])
    np.random.seed(3)
    be.estimate_lipschitz(A, n_iter=1)
    after = np.random.randn()
    np.random.seed(3)
    np.random.randn(A.shape[1])
    assert after == np.random.randn()


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("fista", "fista_delta", "ista")))
def test_prox_gradient_traces(be, name, key):
    rtol = RTOL_F64
    out, spec = harness.run_case(be, name, key)
    harness.check_case(out, spec, name, key, rtol)


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("fista", "fista_delta", "ista"),
                                                            designs=("wide", "widex")))
def test_persistent_solve_kernel_and_two_launch_path(be, name, key, fused, monkeypatch):
    """The streaming designs run every solve as ONE launch of the persistent kernel (gradient pass,
    in-kernel cross-CTA reduction, prox / momentum / stop rules on column slices, three grid barriers
    per pass); FOS_FUSED=0 keeps the (gradient, epilogue) launch pairs.  Both against the same golden
    traces of the reference, and the launch count says which one ran."""
    from fastoptsolver_b200 import iterative_solvers as S
    monkeypatch.setenv("FOS_FUSED", fused)
    # the persistent-kernel runs also force the residual recurrence on (short solves default to the second dot):
    # every fixed-step golden trace, objective trace included, through GM_QREC
    monkeypatch.setenv("FOS_QREC", "1" if fused == "1" else "0")
    out, spec = harness.run_case(be, name, key)
    harness.check_case(out, spec, name, key, RTOL_F64)
    info = S.last_run["solver"]
    if info["passes"] > 0:
        assert (info["kernel_launches"] == 1) == (fused == "1"), info


def test_persistent_solve_kernel_at_scale(monkeypatch):
    """200 000 x 2048 (1351 rows per CTA, ring of 6 stages wrapping from pass to pass): the persistent
    kernel and the two-launch path give the same iterates (1e-12: the cross-CTA sums associate
    differently), the same objective traces and Armijo counts; two runs of the persistent kernel are
    bit-identical."""
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    des = DeviceDesign.synthetic(200_000, 2048, seed=3, noise_std=0.5, rho1=0.5, rho2=0.7)
    a1 = 0.05 * des.lambda_max()
    np.random.seed(0)
    L = S.estimate_lipschitz(des)
    variants = [
        ("fista", dict(max_iter=30)),
        ("fista", dict(max_iter=25, backtracking=True, t_init_factor=3.0)),
        ("fista", dict(max_iter=60, adaptive_restart=True, restart_threshold=0.9)),
        ("fista", dict(max_iter=400, tol_ratio=0.3)),        # fires at iteration 6 (ratio 0.2505)
        ("fista_delta", dict(max_iter=25, backtracking=True, t_init_factor=2.0, eta=0.7)),
    ]
    for kind, kw in variants:
        res = {}
        for fused in ("1", "0", "1"):
            monkeypatch.setenv("FOS_FUSED", fused)
            d2 = DeviceDesign.from_device_pointers(*des_pointers(des), des.shape[0], des.shape[1], des.dtype, _lda(des),
                                                   keepalive=des)
            np.random.seed(0)
            if kind == "fista":
                x, h = S.fista(d2, None, "elasticnet", a1, 0.01 * a1, return_history=True, **kw)
            else:
                x, h = S.fista_delta(d2, None, "elasticnet", a1, 0.01 * a1, 3.0, return_history=True, **kw)
            info = dict(S.last_run["solver"])
            assert (info["kernel_launches"] == 1) == (fused == "1")
            res.setdefault(fused, []).append((x, h, list(S.ls_call_iters), info["iters"]))
            d2.close()
        (xa, ha, lsa, ita), (xb, hb, lsb, itb) = res["1"][0], res["0"][0]
        assert ita == itb and lsa == lsb, (kind, kw)
        assert harness.rel_err(xa, xb) <= 1e-12
        assert harness.rel_err(ha["obj"], hb["obj"]) <= 1e-12
        for u, v in zip(ha["x"], hb["x"]):
            assert np.linalg.norm(u - v) <= 1e-12 * max(np.linalg.norm(v), 1e-300) + 1e-300
        xc, hc, lsc, itc = res["1"][1]
        assert xc.tobytes() == xa.tobytes() and np.asarray(hc["obj"]).tobytes() == np.asarray(ha["obj"]).tobytes()
    des.close()


@pytest.mark.parametrize("n,d,dtype", [(30000, 1500, np.float32), (20000, 4100, np.float64), (20000, 8192, np.float32),
                                       (50000, 641, np.float64), (9000, 2048, np.float64)])
def test_persistent_solve_kernel_shapes(n, d, dtype, monkeypatch):
    """The persistent kernel on the other streaming builds: fp32 storage, odd d (padded leading dimension,
    a last column pair that is half padding), d > 4096 (512-thread build, 28 column pairs per CTA slice),
    short row blocks (ring shallower than the shared-memory budget).  Against the two-launch path (1e-12) for
    fista with Armijo and fista_delta with a fixed step, and against the oracle on a downloaded row block."""
    import oracle
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    des = DeviceDesign.synthetic(n, d, dtype, seed=4, noise_std=0.5, rho1=0.5, rho2=0.7)
    des.standardize()
    a1 = 0.05 * des.lambda_max()
    runs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("FOS_FUSED", fused)
        d2 = DeviceDesign.from_device_pointers(*des_pointers(des), n, d, des.dtype, _lda(des), keepalive=des)
        np.random.seed(0)
        xa, ha = S.fista(d2, None, "elasticnet", a1, 0.02 * a1, max_iter=12, backtracking=True, t_init_factor=2.0,
                         return_history=True)
        la = list(S.ls_call_iters)
        launches = S.last_run["solver"]["kernel_launches"]
        np.random.seed(0)
        xb, hb = S.fista_delta(d2, None, "lasso", a1, 0.0, 3.0, max_iter=12, return_history=True)
        assert (launches == 1) == (fused == "1"), (fused, launches)
        runs[fused] = (xa, ha, la, xb, hb)
        d2.close()
    (xa, ha, la, xb, hb), (ya, ga, ma, yb, gb) = runs["1"], runs["0"]
    assert la == ma
    assert harness.rel_err(xa, ya) <= 1e-12 and harness.rel_err(ha["obj"], ga["obj"]) <= 1e-12
    assert harness.rel_err(xb, yb) <= 1e-12 and harness.rel_err(hb["obj"], gb["obj"]) <= 1e-12
    assert not np.any(np.isnan(xa)) and xa.shape == (d,)
    # a row block small enough for the oracle, solved by the persistent kernel again
    rows = 3000
    A_blk, b_blk = des.download(0, rows)
    monkeypatch.setenv("FOS_FUSED", "1")
    blk = DeviceDesign.from_host(A_blk, b_blk)
    a1b = 0.05 * blk.lambda_max()
    np.random.seed(0)
    x_dev, h_dev = S.fista(blk, None, "lasso", a1b, 0.0, max_iter=15, return_history=True)
    np.random.seed(0)
    x_ref, h_ref = oracle.fista(np.asarray(A_blk, dtype=np.float64), b_blk, "lasso", a1b, 0.0, max_iter=15, return_history=True)
    assert harness.rel_err(x_dev, x_ref) <= 1e-10
    np.testing.assert_allclose(h_dev["obj"], h_ref["obj"], rtol=1e-10)
    blk.close()
    des.close()


def test_objective_from_the_residual_recurrence(monkeypatch):
    """Fixed-step solves with history record 0.5 |A x_k - b|^2 from the row-wise residual recurrence
    q_k = (r_y + beta q_{k-1}) / (1 + beta) (GM_QREC) instead of a second dot product per pass.  Against the
    second-dot path (FOS_QREC=0) over 300 iterations -- the recurrence must not drift -- for fista, fista with
    restarts (beta drops to 0 and the chain restarts), fista_delta (beta != 0 from the first update) and ista
    (beta = 0 throughout); the iterates are bit-identical (the objective never feeds back)."""
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import operators as OPS
    from fastoptsolver_b200.design import DeviceDesign
    des = DeviceDesign.synthetic(60_000, 1024, seed=8, noise_std=1.0, rho1=0.8, rho2=0.9)
    des.standardize()
    a1 = 0.02 * des.lambda_max()
    out = {}
    for q in ("1", "0"):
        monkeypatch.setenv("FOS_QREC", q)
        res = []
        np.random.seed(0)
        res.append(S.fista(des, None, "elasticnet", a1, 0.1 * a1, max_iter=300, return_history=True))
        np.random.seed(0)
        res.append(S.fista(des, None, "lasso", a1, 0.0, max_iter=300, adaptive_restart=True, restart_threshold=0.95,
                           return_history=True))
        np.random.seed(0)
        res.append(S.fista_delta(des, None, "lasso", a1, 0.0, 2.5, max_iter=300, return_history=True))
        np.random.seed(0)
        L = S.estimate_lipschitz(des)
        g, grad_g, prox_h = OPS.ista_callables(des, None, a1, 0.0)
        x, log = S.ista(np.zeros(1024), g, grad_g, prox_h, L, max_iter=100, return_history=True)
        res.append((x, {"obj": list(S.last_run["ista_obj"])}))
        out[q] = res
    for (xa, ha), (xb, hb) in zip(out["1"], out["0"]):
        assert xa.tobytes() == xb.tobytes()
        oa, ob = np.asarray(ha["obj"]), np.asarray(hb["obj"])
        assert oa.shape == ob.shape and len(oa) >= 100
        assert np.max(np.abs(oa - ob) / np.abs(ob)) <= 1e-12
    des.close()


def _lda(des):
    import ctypes as C
    from fastoptsolver_b200 import _lib
    n, d, dt, lda = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
    _lib.check(_lib.load().fos_design_shape(des.handle, C.byref(n), C.byref(d), C.byref(dt), C.byref(lda)))
    return lda.value


def des_pointers(des):
    import ctypes as C
    from fastoptsolver_b200 import _lib
    a, b = C.c_void_p(), C.c_void_p()
    _lib.check(_lib.load().fos_design_pointers(des.handle, C.byref(a), C.byref(b)))
    return a.value, b.value


@pytest.mark.parametrize("name,key", harness.all_case_ids(solvers=("lbfgs",)))
def test_lbfgs_traces(be, name, key):
    out, spec = harness.run_case(be, name, key)
    # short, well-conditioned runs agree to ~1e-12; the 50-iteration lasso runs (plain
    # least squares, cond ~ 1e2-1e3) amplify summation-order noise
    harness.check_case(out, spec, name, key, 1e-9, lbfgs_trace_rtol=1e-6)


@pytest.mark.parametrize("kernel", ["generic", "stream"])
def test_both_gradient_kernels(kernel, monkeypatch):
    """Force each kernel on the same design; both must meet the 1e-10 bar."""
    from fastoptsolver_b200 import design as D
    monkeypatch.setenv("FOS_FORCE_KERNEL", kernel)
    D.clear_cache()
    try:
        be_ = harness.cuda_backend()
        for name, key in (("mid", "fista/lasso-fixed-t1.0"), ("mid", "fista/elasticnet-armijo-t2.0"),
                          ("odd", "fista_delta/lasso-armijo-t2.0"), ("c1", "fista/lasso-armijo-t1.0"),
                          ("mid32", "fista/lasso-armijo-t2.0")):
            out, spec = harness.run_case(be_, name, key)
            harness.check_case(out, spec, name, key, RTOL_F64)
    finally:
        D.clear_cache()


def test_api_quirks(be):
    A, b = cases.design("c1")
    np.random.seed(0)
    x1 = be.fista(A, b, "bogus", 1.0, 0.0, max_iter=5)
    np.random.seed(0)
    x2 = be.fista(A, b, "lasso", 1.0, 0.0, max_iter=5)
    np.testing.assert_array_equal(x1, x2)
    assert x1.dtype == np.float64 and x1.shape == (5,)
    with pytest.raises(AssertionError):
        be.fista_delta(A, b, "lasso", 1.0, 0.0, 2.0)
    np.random.seed(0)
    be.fista_delta(A, b, "bogus", 1.0, 0.0, 3.0, max_iter=3)
    with pytest.raises(ValueError):
        be.fista_delta(A, b, "bogus", 1.0, 0.0, 3.0, max_iter=3, return_history=True)
    with pytest.raises(ValueError, match="Unsupported reg_type='l0'"):
        be.lbfgs_cls("l0", 1.0, 1.0)
    s = be.lbfgs_cls("ridge", 0.1, 0.5, max_iter=5)
    s.fit(A, b)
    n1 = len(s.history_)
    s.fit(A, b)
    assert len(s.history_) == 2 * n1
    np.random.seed(0)
    _, h = be.fista(A, b, "lasso", 1.0, 0.0, max_iter=7, return_history=True)
    assert (len(h["x"]), len(h["obj"])) == (8, 7)
    assert all(isinstance(v, np.ndarray) for v in h["x"])
    h["x"][0][0] = 123.0                      # independent copies, like x.copy() in the reference
    assert h["x"][1][0] != 123.0
    np.random.seed(0)
    _, h = be.fista_delta(A, b, "lasso", 1.0, 0.0, 3.0, max_iter=7, return_history=True)
    assert (len(h["x"]), len(h["obj"])) == (7, 7)
    # max_iter = 0
    np.random.seed(0)
    x0, h0 = be.fista(A, b, "lasso", 1.0, 0.0, max_iter=0, return_history=True)
    assert np.all(x0 == 0) and len(h0["x"]) == 1 and h0["obj"] == []


def test_metrics_contract(be):
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import lbfgs as LB
    A, b = cases.design("mid")
    assert LB.grad_call_times is S.grad_call_times
    np.random.seed(0)
    S.fista(A, b, "lasso", 5.0, 0.0, backtracking=True, max_iter=20)
    m = S.get_metrics()
    assert set(m) == {"grad_num_calls", "grad_time_total", "grad_time_mean", "ls_num_calls", "ls_time_total",
                      "ls_time_mean", "ls_iters_total"}
    assert m["grad_num_calls"] == 20 and m["ls_num_calls"] == 20
    assert m["grad_time_total"] > 0 and m["ls_time_total"] > 0
    ident = id(S.grad_call_times)
    S.reset_metrics()
    assert id(S.grad_call_times) == ident and S.get_metrics()["grad_num_calls"] == 0
    # module global C is honoured at call time
    np.random.seed(0)
    xa = S.fista(A, b, "lasso", 5.0, 0.0, backtracking=True, t_init_factor=4.0, max_iter=10)
    ls_a = list(S.ls_call_iters)
    old = S.C
    try:
        S.C = 0.9
        np.random.seed(0)
        S.fista(A, b, "lasso", 5.0, 0.0, backtracking=True, t_init_factor=4.0, max_iter=10)
        ls_b = list(S.ls_call_iters)
    finally:
        S.C = old
    # the same two runs through the oracle with its Armijo constant switched the same way: the
    # shrink counts must agree exactly, and the stricter constant shrinks strictly more here
    import oracle
    from oracle import ref_numpy
    want = []
    try:
        for c_val in (old, 0.9):
            ref_numpy.ARMIJO_C = c_val
            np.random.seed(0)
            oracle.fista(A, b, "lasso", 5.0, 0.0, backtracking=True, t_init_factor=4.0, max_iter=10)
            want.append(list(oracle.METRICS["ls_iters"]))
    finally:
        ref_numpy.ARMIJO_C = old
    assert ls_a == want[0] and ls_b == want[1]
    assert sum(ls_b) > sum(ls_a)
    assert xa.shape == (A.shape[1],)


def test_large_size_properties():
    """At a size the oracle cannot hold a golden for: properties that do not depend on it.
    Linearity of the gradient in b-free mode, determinism, and agreement of the fused
    single-pass loss with the separate objective pass."""
    from fastoptsolver_b200.design import DeviceDesign
    des = DeviceDesign.synthetic(200_000, 2048, seed=5, noise_std=1.0, rho1=0.5, rho2=0.7)
    rng = np.random.default_rng(0)
    d = des.shape[1]
    x = rng.standard_normal(d)
    y = rng.standard_normal(d)
    zero = np.zeros(d)
    l0, g0 = des.grad(zero)               # g0 = -A^T b
    lx, gx = des.grad(x)
    ly, gy = des.grad(y)
    lxy, gxy = des.grad(x + y)
    # A^T A (x+y) = A^T A x + A^T A y
    lhs = gxy - g0
    rhs = (gx - g0) + (gy - g0)
    assert harness.rel_err(lhs, rhs) <= 1e-12
    assert abs(des.objective(x, 0, 0.0, 0.0) - lx) <= 1e-12 * lx
    assert des.grad(x)[1].tobytes() == gx.tobytes()
    # a row block downloaded to the host reproduces its share of the gradient
    import oracle
    A_blk, b_blk = des.download(1000, 4096)
    des_blk = DeviceDesign.from_host(A_blk, b_blk)
    loss_ref, g_ref = oracle.smooth_value_and_grad(x, A_blk, b_blk)
    loss_blk, g_blk = des_blk.grad(x)
    assert abs(loss_blk - loss_ref) <= 1e-12 * loss_ref and harness.rel_err(g_blk, g_ref) <= 1e-12
    des.close()


def test_sweep_driver_small():
    """The notebook replacement: one scenario, all 19 variants, CSV out; ISTA's by-product
    objective equals compute_objective of its iterates."""
    import tempfile
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200 import sweep
    from fastoptsolver_b200.design import DeviceDesign
    from fastoptsolver_b200.operators import ista_callables
    des = DeviceDesign.synthetic(5000, 40, seed=0, noise_std=0.5, rho1=0.5, rho2=0.7)
    traces, timing, meta = sweep.run_scenario(des, max_iter=30)
    assert set(traces) == {"L-BFGS", "ISTA", "FISTA", "FISTA-delta"}
    assert len(traces["FISTA"]) == 6 and len(traces["ISTA"]) == 6 and len(traces["L-BFGS"]) == 2
    sub = sweep.suboptimality(traces)
    assert all(v >= 0 for c in sub.values() for tr in c.values() for v in tr)
    with tempfile.TemporaryDirectory() as tmp:
        sweep.write_csv(os.path.join(tmp, "x.csv"), sub)
        assert os.path.getsize(os.path.join(tmp, "x.csv")) > 1000
    a1 = meta["alpha1"]
    np.random.seed(0)
    L = S.estimate_lipschitz(des)
    g, grad_g, prox_h = ista_callables(des, None, a1, 0.0)
    _, log = S.ista(np.zeros(40), g, grad_g, prox_h, L, max_iter=10, return_history=True)
    direct = [des.objective(x, 1, a1, 0.0) for x in log["x"][1:]]
    np.testing.assert_allclose(S.last_run["ista_obj"], direct, rtol=1e-12)
    des.close()


def _lbfgs_pair(A, b, reg, a1, a2, **kw):
    import oracle
    from fastoptsolver_b200.lbfgs import LBFGSSolver
    ref = oracle.LBFGSSolver(reg, a1, a2, **kw)
    ref.fit(A, b)
    dev = LBFGSSolver(reg, a1, a2, driver="device", **kw)
    dev.fit(A, b)
    return ref, dev


def test_device_lbfgs_well_conditioned():
    """Device L-BFGS (two-loop + More'-Thuente on the GPU) against scipy's L-BFGS-B driven by the
    oracle: on a well-conditioned design the whole objective trace agrees."""
    rng = np.random.default_rng(4)
    n, d = 4000, 64
    A = rng.standard_normal((n, d)) / np.sqrt(n)
    A += 0.1 * rng.standard_normal((n, 1)) / np.sqrt(n)
    x_true = rng.standard_normal(d)
    b = A @ x_true + 0.01 * rng.standard_normal(n)
    for reg, a1, a2 in (("ridge", 0.0, 0.05), ("elasticnet", 0.01, 0.05), ("lasso", 0.01, 0.0)):
        ref, dev = _lbfgs_pair(A, b, reg, a1, a2, max_iter=100, tol=1e-9)
        assert len(dev.history_) == len(ref.history_), (reg, len(dev.history_), len(ref.history_))
        err = np.abs(np.array(dev.history_) - np.array(ref.history_)) / np.abs(ref.history_)
        assert err.max() <= 1e-9, (reg, err.max())
        assert abs(dev.final_obj_ - ref.final_obj_) <= 1e-10 * abs(ref.final_obj_)
        assert harness.rel_err(dev.x_, ref.x_) <= 1e-6


@pytest.mark.parametrize("name", ["mid", "wide", "c1"])
def test_device_lbfgs_golden_designs(name):
    """Device L-BFGS against the reference's golden L-BFGS traces.  The bound is the one the CPU
    model of the same algorithm meets against scipy on these designs
    (tests/test_lbfgs_model_cpu.py::test_model_equals_scipy_on_golden_designs: identical iteration
    counts, objective traces to 1e-10, x to 1e-9), relaxed by one order of magnitude for the
    different summation order of the GPU gradient."""
    from fastoptsolver_b200.lbfgs import LBFGSSolver
    A, b = cases.design(name)
    g = harness.golden(name)
    for key in ("lbfgs/ridge", "lbfgs/elasticnet"):
        a1, a2 = (float(v) for v in g[f"{key}/alpha"])
        spec = cases.solver_specs(name, A, b)[key]
        dev = LBFGSSolver(spec["reg_type"], a1, a2, driver="device", **spec["kw"])
        dev.fit(A, b)
        ref_h = g[f"{key}/hobj"]
        assert len(dev.history_) == len(ref_h), (name, key, len(dev.history_), len(ref_h))
        err = np.abs(np.asarray(dev.history_) - ref_h) / np.abs(ref_h)
        assert err.max() <= 1e-9, (name, key, err.max())
        assert harness.rel_err(dev.x_, g[f"{key}/x"]) <= 1e-8, (name, key, harness.rel_err(dev.x_, g[f"{key}/x"]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("d", [1500, 2048, 3000, 4096])
def test_streaming_builds_mid_widths(d, dtype):
    """The streaming builds for 1024 < lda <= 4096 -- fp32 storage (256,8,4) / (256,16,2), fp64
    (256,8,2) / (256,16,1) -- on row counts that leave partial last stages, with and without column
    padding inside a thread's vector: gradient, objective, a fixed-step and an Armijo FISTA run
    against the oracle.  fp32 storage is compared at the fp64 tolerance (arithmetic is fp64 on both
    sides; north_star allows 1e-5)."""
    import oracle
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    rng = np.random.default_rng(d)
    n = 1237
    A = np.asarray(rng.standard_normal((n, d)) / np.sqrt(n), dtype=dtype)
    A[:, ::3] *= 2.0
    b = rng.standard_normal(n)
    A64 = A.astype(np.float64)
    des = DeviceDesign.from_host(A, b)
    x = rng.standard_normal(d)
    loss, g = des.grad(x, 0.3)
    lr, gr = oracle.smooth_value_and_grad(x, A64, b, 0.3)
    assert abs(loss - lr) <= 1e-12 * lr and harness.rel_err(g, gr) <= 1e-12
    assert abs(des.objective(x, 3, 0.2, 0.3) - oracle.compute_objective(x, A64, b, "elasticnet", 0.2, 0.3)) <= 1e-12 * lr
    a1 = 0.2 * float(np.max(np.abs(A64.T @ b)))
    for kw in (dict(), dict(backtracking=True, t_init_factor=2.0)):
        np.random.seed(0)
        xr, hr = oracle.fista(A64, b, "lasso", a1, 0.0, max_iter=15, return_history=True, **kw)
        ls_ref = list(oracle.METRICS["ls_iters"])
        np.random.seed(0)
        xg, hg = S.fista(des, None, "lasso", a1, 0.0, max_iter=15, return_history=True, **kw)
        assert harness.rel_err(xg, xr) <= 1e-10
        np.testing.assert_allclose(hg["obj"], hr["obj"], rtol=1e-10)
        assert list(S.ls_call_iters) == ls_ref
        assert np.array_equal(xg == 0.0, xr == 0.0), "sparsity pattern differs"
    des.close()


@pytest.mark.parametrize("dtype,d", [(np.float32, 2048), (np.float32, 4096), (np.float64, 4096)])
def test_streaming_ring_wraps(dtype, d):
    """Enough rows per CTA for the shared-memory ring to wrap many times (the small cases above never
    refill a slot): one gradient + both dots against numpy on the downloaded design."""
    import oracle
    from fastoptsolver_b200.design import DeviceDesign
    n = 148 * 150 + 77
    des = DeviceDesign.synthetic(n, d, dtype, seed=4, noise_std=1.0, rho1=0.5, rho2=0.7)
    A, b = des.download()
    A64 = A.astype(np.float64)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(d) * (rng.random(d) < 0.2)
    loss, g = des.grad(x, 0.0)
    lr, gr = oracle.smooth_value_and_grad(x, A64, b, 0.0)
    assert abs(loss - lr) <= 1e-12 * lr and harness.rel_err(g, gr) <= 1e-12
    assert abs(des.objective(x, 1, 0.5, 0.0) - oracle.compute_objective(x, A64, b, "lasso", 0.5, 0.0)) <= 1e-12 * lr
    des.close()


def test_dropin_module_global_C_reaches_the_device():
    """The reference's writable module global ``C`` (iterative_solvers.py:11), set on the DROP-IN
    module the way a notebook would, changes the Armijo decisions taken on the device."""
    import importlib
    import sys
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fastoptsolver_b200", "dropin")
    sys.path.insert(0, d)
    try:
        sys.modules.pop("iterative_solvers", None)
        IS = importlib.import_module("iterative_solvers")
        A, b = cases.design("mid")
        counts = []
        for c_val in (1e-2, 0.9):
            IS.C = c_val
            np.random.seed(0)
            IS.fista(A, b, "lasso", 5.0, 0.0, backtracking=True, t_init_factor=4.0, max_iter=10)
            counts.append(list(IS.ls_call_iters))
        IS.C = 1e-2
        assert counts[0][0] == 1 and counts[1][0] == 4, counts      # reference values (oracle run)
    finally:
        IS.C = 1e-2
        sys.path.remove(d)
        sys.modules.pop("iterative_solvers", None)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_standardize(dtype):
    """z-scoring on the device == datagen.standardize on the host (what the notebook did)."""
    from fastoptsolver_b200 import datagen
    from fastoptsolver_b200.design import DeviceDesign
    A, b, _ = datagen.generate_correlated_design(3001, 35, seed=2)
    A = A.astype(dtype)
    des = DeviceDesign.from_host(A, b)
    mu, sd, bm = des.standardize()
    A_ref, b_ref = datagen.standardize(A.astype(np.float64), b)
    A_dev, b_dev = des.download()
    tol = 1e-12 if dtype == np.float64 else 2e-6
    assert np.max(np.abs(A_dev - A_ref)) <= tol * max(1.0, np.abs(A_ref).max())
    np.testing.assert_allclose(b_dev, b_ref, rtol=0, atol=1e-11 * np.abs(b).max())
    np.testing.assert_allclose(mu, A.astype(np.float64).mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(sd, A.astype(np.float64).std(axis=0), rtol=1e-12)
    des.close()


def test_design_cache_semantics():
    """Default: bare numpy inputs are uploaded on every call (the reference re-reads them), so an
    in-place edit between two calls is always seen.  Opt-in cache: reuse only while the arrays are
    the same, unchanged objects."""
    import oracle
    from fastoptsolver_b200 import design as D
    from fastoptsolver_b200 import iterative_solvers as S
    rng = np.random.default_rng(0)
    A = rng.standard_normal((300, 24))
    b = rng.standard_normal(300)
    D.clear_cache()
    D.set_cache(None)
    assert not D.cache_enabled()
    # drop-in behaviour: edit a feature in place between two calls -> second call sees it
    np.random.seed(0)
    x1 = S.fista(A, b, "lasso", 1.0, 0.0, max_iter=10)
    A[:, 7] *= 0.25
    np.random.seed(0)
    x2 = S.fista(A, b, "lasso", 1.0, 0.0, max_iter=10)
    np.random.seed(0)
    x2_ref = oracle.fista(A, b, "lasso", 1.0, 0.0, max_iter=10)
    assert harness.rel_err(x2, x2_ref) <= 1e-10 and not np.array_equal(x1, x2)
    try:
        D.set_cache(True)
        d1 = D.as_design(A, b)
        assert D.as_design(A, b) is d1                      # same objects, same content: reuse
        b[5] += 1.0                                         # b is fully fingerprinted
        d2 = D.as_design(A, b)
        assert d2 is not d1
        A[17, 3] = 7.0                                      # small matrices are hashed completely
        d3 = D.as_design(A, b)
        assert d3 is not d2
        A2 = A.copy()                                       # equal content, different object: upload again
        assert D.as_design(A2, b) is not d3
        loss, g = d3.grad(np.ones(24))
        lr, gr = oracle.smooth_value_and_grad(np.ones(24), A, b)
        assert abs(loss - lr) <= 1e-12 * lr and harness.rel_err(g, gr) <= 1e-12
        with pytest.raises(ValueError):
            d3.grad(np.ones(5))
        # a large matrix (sampled fingerprint): a column edit and a sparse row edit are both caught
        big = rng.standard_normal((1 << 16, 64))
        bb = rng.standard_normal(1 << 16)
        e1 = D.as_design(big, bb)
        assert D.as_design(big, bb) is e1
        big[:, 33] = 0.0
        e2 = D.as_design(big, bb)
        assert e2 is not e1
        big[1000:9000] += 1.0
        assert any(D.as_design(big, bb) is not e2 for _ in range(20))
    finally:
        D.set_cache(None)
        D.clear_cache()
    with pytest.raises(NotImplementedError):
        D.DeviceDesign.from_host(np.zeros((4, 9000)), np.zeros(4))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("d", [4100, 8192])
def test_wide_rows(d, dtype):
    """4096 < d <= 8192: the 512-thread build (all modes) and the gradient-only 256x32 build."""
    import oracle
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    rng = np.random.default_rng(8)
    n = 700
    A = np.asarray(rng.standard_normal((n, d)) / np.sqrt(n), dtype=dtype)
    b = rng.standard_normal(n)
    A64 = A.astype(np.float64)
    des = DeviceDesign.from_host(A, b)
    x = rng.standard_normal(d)
    loss, g = des.grad(x, 0.3)                                   # gradient-only kernel
    lr, gr = oracle.smooth_value_and_grad(x, A64, b, 0.3)
    assert abs(loss - lr) <= 1e-12 * lr and harness.rel_err(g, gr) <= 1e-12
    assert abs(des.objective(x, 3, 0.2, 0.3) - oracle.compute_objective(x, A64, b, "elasticnet", 0.2, 0.3)) <= 1e-12 * lr
    a1 = 0.2 * float(np.max(np.abs(A64.T @ b)))
    np.random.seed(0)
    xr, hr = oracle.fista(A64, b, "lasso", a1, 0.0, max_iter=15, return_history=True)
    np.random.seed(0)
    xg, hg = S.fista(des, None, "lasso", a1, 0.0, max_iter=15, return_history=True)   # power iteration: lite; loop: full
    assert harness.rel_err(xg, xr) <= 1e-10
    np.testing.assert_allclose(hg["obj"], hr["obj"], rtol=1e-10)
    des.close()


def test_full_size_properties_1Mx4096():
    """BASELINE config 3 at full size (1M x 4096 fp64, 32.8 GB): properties that need no oracle.
    The design equals the concatenation of two independently generated row shards (Philox is
    keyed by the global row), so gradient and loss must be the shard sums; plus linearity,
    determinism, fused loss == objective pass, and a short FISTA run whose recorded objective
    equals the objective of its iterates."""
    from fastoptsolver_b200 import iterative_solvers as S
    from fastoptsolver_b200.design import DeviceDesign
    n, d = 1_000_000, 4096
    sc = dict(seed=3, noise_std=1.0, rho1=0.8, rho2=0.9)
    full = DeviceDesign.synthetic(n, d, **sc)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(d) * (rng.random(d) < 0.1)
    loss, g = full.grad(x, 0.0)
    assert full.grad(x, 0.0)[1].tobytes() == g.tobytes()
    assert abs(full.objective(x, 0, 0.0, 0.0) - loss) <= 1e-12 * loss
    lo = DeviceDesign.synthetic(n // 2, d, row0=0, **sc)
    hi = DeviceDesign.synthetic(n - n // 2, d, row0=n // 2, **sc)
    l1, g1 = lo.grad(x, 0.0)
    l2, g2 = hi.grad(x, 0.0)
    assert abs((l1 + l2) - loss) <= 1e-12 * loss
    assert harness.rel_err(g1 + g2, g) <= 1e-12
    lo.close()
    hi.close()
    y = rng.standard_normal(d)
    g0 = full.grad(np.zeros(d))[1]
    lhs = full.grad(x + y)[1] - g0
    rhs = (g - g0) + (full.grad(y)[1] - g0)
    assert harness.rel_err(lhs, rhs) <= 1e-12
    a1 = 0.1 * full.lambda_max()
    np.random.seed(0)
    xk, h = S.fista(full, None, "lasso", a1, 0.0, max_iter=4, return_history=True)
    for xi, oi in zip(h["x"][1:], h["obj"]):
        assert abs(full.objective(xi, 1, a1, 0.0) - oi) <= 1e-12 * abs(oi)
    assert h["obj"][-1] < h["obj"][0]
    full.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_generator_matches_numpy_model(dtype):
    """csrc/datagen.cu (Philox4x32-10 keyed by (row, group, draw), Box-Muller, the reference's
    5-column recipe) against its independent numpy restatement (oracle/datagen_model.py), which
    `bench.py --impl reference` uses to build the same design without the product.  Agreement is to
    rounding (CUDA's log / sincospi vs numpy's log / sin / cos), for any row offset."""
    from oracle import datagen_model
    from fastoptsolver_b200.design import DeviceDesign
    sc = dict(seed=(7 << 32) | 12345, noise_std=0.5, rho1=0.5, rho2=0.7)
    for n, d, row0 in ((513, 37, 0), (300, 640, 10_000_000_000), (64, 4, 5)):
        des = DeviceDesign.synthetic(n, d, dtype, row0=row0, **sc)
        A, b = des.download()
        A_ref, b_ref = datagen_model.synth_rows(n, d, row0=row0, dtype=dtype, threads=1, **sc)
        tol = 1e-13 if dtype == np.float64 else 2e-7
        assert np.max(np.abs(A.astype(np.float64) - A_ref.astype(np.float64))) <= tol * 8.0
        np.testing.assert_allclose(b, b_ref, rtol=0, atol=(1e-11 if dtype == np.float64 else 1e-4) * max(1.0, np.abs(b_ref).max()))
        des.close()
    # the virtual design is row-addressable: a shard generated at an offset equals those rows of the whole
    whole, _ = datagen_model.synth_rows(100, 20, row0=0, threads=1, **sc)
    part, _ = datagen_model.synth_rows(30, 20, row0=50, threads=1, **sc)
    assert np.array_equal(whole[50:80], part)
