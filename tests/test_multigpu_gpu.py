"""GPU, >= 2 devices: the row-sharded path with the fused peer-memory all-reduce against the
golden traces and against the single-GPU result (differences are summation order only)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = [("wide", "fista/lasso-fixed-t1.0"), ("wide", "fista/lasso-armijo-t2.0"),
         ("mid", "fista_delta/elasticnet-armijo-t2.0"), ("mid", "ista/lasso-armijo-t2.0"),
         ("mid", "lbfgs/ridge"), ("odd", "fista/lasso-armijo-t2.0"), ("mid", "fista/lasso-tol"),
         # 1200 rows per rank at world 2: the persistent solve kernel with its push-model slice exchange
         ("widex", "fista/restart-armijo"), ("widex", "fista_delta/step-stop"), ("widex", "ista/nonzero-start-armijo")]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import cases
        import harness
        from fastoptsolver_b200 import multigpu
        base = harness.cuda_backend()
        for name, key in CASES:
            A, b = cases.design(name)
            lo, hi = multigpu.shard_bounds(A.shape[0], rank, world)
            shard = multigpu.sharded_from_host(np.ascontiguousarray(A[lo:hi]), b[lo:hi], dist, device=rank)

            class _Sharded:        # present the shard wherever the harness passes (A, b)
                pass
            be = harness.Backend(
                name="cuda-sharded",
                fista=lambda A_, b_, *a, **k: base.fista(shard, None, *a, **k),
                fista_delta=lambda A_, b_, *a, **k: base.fista_delta(shard, None, *a, **k),
                ista=base.ista,
                estimate_lipschitz=lambda A_, *a, **k: base.estimate_lipschitz(shard, *a, **k),
                lbfgs_cls=type("L", (base.lbfgs_cls,), {"fit": lambda self, A_, b_=None: base.lbfgs_cls.fit(self, shard)}),
                ista_callables=lambda A_, b_, a1, a2: base.ista_callables(shard, None, a1, a2),
                ls_iters=base.ls_iters, grad_calls=base.grad_calls)
            out, spec = harness.run_case(be, name, key)
            harness.check_case(out, spec, name, key, 1e-10, lbfgs_trace_rtol=1e-7)
            if key.startswith("lbfgs/"):
                # device driver on the sharded design: same converged objective, identical on all ranks
                g = harness.golden(name)
                a1, a2 = (float(v) for v in g[f"{key}/alpha"])
                dev = base.lbfgs_cls(spec["reg_type"], a1, a2, driver="device", **spec["kw"])
                dev.fit(shard)
                ref_h = g[f"{key}/hobj"]
                np.testing.assert_allclose(dev.history_[:3], ref_h[:3], rtol=1e-9)
                assert abs(dev.history_[-1] - ref_h[-1]) <= 1e-7 * abs(ref_h[-1])
                out = {"x": dev.x_}
            # all ranks hold bit-identical iterates
            t = torch.from_numpy(np.asarray(out["x"]).copy()).cuda()
            lst = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            assert all(torch.equal(lst[0], v) for v in lst), "ranks diverged"
            shard.close()
        # A shard large enough for the streaming kernel (and, from 4 ranks on, the rate-weighted
        # SM-indexed row blocks): the sharded solve equals the single-GPU solve of the same design
        from fastoptsolver_b200 import iterative_solvers as S
        from fastoptsolver_b200.design import DeviceDesign
        n_tot, d_big = world * 40000, 1024
        sc = dict(seed=9, noise_std=0.5, rho1=0.5, rho2=0.7)
        shard = multigpu.sharded_synthetic(n_tot, d_big, dist, device=rank, **sc)
        full = DeviceDesign.synthetic(n_tot, d_big, device=rank, **sc)
        lam = full.lambda_max()
        assert abs(shard.lambda_max() - lam) <= 1e-12 * lam
        for kw in (dict(), dict(backtracking=True, t_init_factor=2.0)):
            np.random.seed(0)
            xs, hs = S.fista(shard, None, "lasso", 0.1 * lam, 0.0, max_iter=30, return_history=True, **kw)
            ls_s = list(S.ls_call_iters)
            np.random.seed(0)
            xf, hf = S.fista(full, None, "lasso", 0.1 * lam, 0.0, max_iter=30, return_history=True, **kw)
            assert harness.rel_err(xs, xf) <= 1e-10
            np.testing.assert_allclose(hs["obj"], hf["obj"], rtol=1e-10)
            assert ls_s == list(S.ls_call_iters)
        multigpu.close(shard, dist)
        full.close()
        # Gram mode on row shards: local SYRK + one all-reduce of G == the single-matrix Gram
        from fastoptsolver_b200.gram import GramDesign
        rng = np.random.default_rng(1)
        A = rng.standard_normal((2001, 128))
        b = rng.standard_normal(2001)
        lo, hi = multigpu.shard_bounds(2001, rank, world)
        shard = multigpu.sharded_from_host(np.ascontiguousarray(A[lo:hi]), b[lo:hi], dist, device=rank)
        gram = GramDesign(shard)
        gram.allreduce(dist)
        G, c = gram.download()
        assert harness.rel_err(G, A.T @ A) <= 1e-13 and harness.rel_err(c, A.T @ b) <= 1e-13
        assert abs(gram.btb - b @ b) <= 1e-12 * (b @ b)
        gram.close()
        shard.close()
        # Gram accumulated under each rank's upload, summed over the ranks (forced on for this
        # small shape): estimate_lipschitz on it == the streaming estimate == numpy, all ranks equal
        rng = np.random.default_rng(2)
        A = rng.standard_normal((world * 6000 + 13, 256))
        A[:, ::4] *= 2.0
        b = rng.standard_normal(A.shape[0])
        lo, hi = multigpu.shard_bounds(A.shape[0], rank, world)
        os.environ["FOS_UPLOAD_GRAM"] = "1"
        shard = multigpu.sharded_from_host(np.ascontiguousarray(A[lo:hi]), b[lo:hi], dist, device=rank)
        os.environ["FOS_UPLOAD_GRAM"] = "0"
        plain = multigpu.sharded_from_host(np.ascontiguousarray(A[lo:hi]), b[lo:hi], dist, device=rank)
        os.environ.pop("FOS_UPLOAD_GRAM")
        assert shard.upload_gram()["state"] == 2 and plain.upload_gram()["state"] == 0
        import oracle
        np.random.seed(3)
        L_ref = oracle.estimate_lipschitz(A)
        np.random.seed(3)
        L_g = S.estimate_lipschitz(shard)
        assert S.last_run["lipschitz"]["via"] == "gram"
        np.random.seed(3)
        L_s = S.estimate_lipschitz(plain)
        assert S.last_run["lipschitz"]["via"] == "stream"
        assert abs(L_g - L_ref) <= 1e-12 * L_ref and abs(L_s - L_ref) <= 1e-12 * L_ref
        t = torch.tensor([L_g], dtype=torch.float64).cuda()
        lst = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        assert all(torch.equal(lst[0], v) for v in lst), "ranks hold different Lipschitz estimates"
        np.random.seed(0)
        xg, hg = S.fista(shard, None, "lasso", 50.0, 0.0, max_iter=25, return_history=True)
        np.random.seed(0)
        xr, hr = oracle.fista(A, b, "lasso", 50.0, 0.0, max_iter=25, return_history=True)
        assert harness.rel_err(xg, xr) <= 1e-10
        np.testing.assert_allclose(hg["obj"], hr["obj"], rtol=1e-10)
        multigpu.close(shard, dist)
        multigpu.close(plain, dist)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_traces(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
