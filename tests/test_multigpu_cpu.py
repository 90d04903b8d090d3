"""CPU, world_size 2 over gloo: the host-side logic of the row-sharded path -- shard
bounds, the handle exchange, and the algebra the fused all-reduce relies on
(sum over shards of A_g^T(A_g y - b_g) == A^T(Ay - b); the oracle stands in for the kernel)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from fastoptsolver_b200 import multigpu
        # 1. handle exchange: every rank ends up with everybody's 64 bytes, in rank order
        mine = bytes([rank]) * 64
        got = multigpu.exchange_bytes(mine, dist)
        assert got == [bytes([r]) * 64 for r in range(world)]
        # 1b. descriptor exchange (the transport of the VMM exchange windows): every rank hands out
        # the read end of a pipe it keeps writing to; the descriptors received from the peers must
        # be live duplicates in this process
        rd, wr = os.pipe()
        fds = multigpu.exchange_fds(rd, dist)
        assert fds[rank] == -1 and all(f >= 0 for r, f in enumerate(fds) if r != rank)
        os.write(wr, bytes([65 + rank]) * (world - 1))       # one byte for every peer to read
        dist.barrier()
        for r, f in enumerate(fds):
            if r != rank:
                assert os.read(f, 1) == bytes([65 + r])
                os.close(f)
        os.close(rd)
        os.close(wr)
        assert multigpu._all_ok(True, dist, None) and not multigpu._all_ok(rank == 0, dist, None)
        # 1c. the collective logic of wiring the exchange windows, driven with stand-ins for the
        # library calls: descriptor route, fall-back to the handle route when one rank cannot do
        # descriptors (the others must drop their windows), forced handle route; the piggy-backed
        # objects of all ranks come back in rank order
        log = []

        def make(vmm_ok):
            pipe = os.pipe()
            os.write(pipe[1], bytes([97 + rank]) * world)

            def alloc_fd():
                log.append("alloc_fd")
                return pipe[0] if vmm_ok else None

            def attach_fd(fds):
                log.append("attach_fd")
                return all(os.read(f, 1) == bytes([97 + r]) for r, f in enumerate(fds) if r != rank)

            def alloc_ipc():
                log.append("alloc_ipc")
                return bytes([rank]) * 64

            def attach_ipc(handles):
                log.append("attach_ipc")
                assert handles == [bytes([r]) * 64 for r in range(world)]

            return alloc_fd, attach_fd, lambda: log.append("free"), alloc_ipc, attach_ipc

        kind, extras = multigpu.share_windows(*make(True), dist, None, piggyback=("gram", rank))
        assert kind == "vmm" and extras == [("gram", r) for r in range(world)]
        assert log == ["alloc_fd", "attach_fd"]
        del log[:]
        kind, extras = multigpu.share_windows(*make(rank != 1), dist, None, piggyback=rank * 10)
        assert kind == "ipc" and extras == [r * 10 for r in range(world)]
        assert log == (["alloc_fd", "alloc_ipc", "attach_ipc"] if rank == 1 else
                       ["alloc_fd", "free", "alloc_ipc", "attach_ipc"])
        del log[:]
        kind, extras = multigpu.share_windows(*make(True), dist, None, piggyback=None, prefer_vmm=False)
        assert kind == "ipc" and extras == [None] * world and log == ["alloc_ipc", "attach_ipc"]
        # 2. sharded gradient algebra on an uneven split
        rng = np.random.default_rng(0)
        n, d = 1001, 17
        A = rng.standard_normal((n, d))
        b = rng.standard_normal(n)
        y = rng.standard_normal(d)
        lo, hi = multigpu.shard_bounds(n, rank, world)
        loss_l, g_l = oracle.smooth_value_and_grad(y, A[lo:hi], b[lo:hi])
        t = torch.from_numpy(np.concatenate([g_l, [loss_l]]))
        dist.all_reduce(t)
        loss, g = oracle.smooth_value_and_grad(y, A, b)
        np.testing.assert_allclose(t.numpy()[:-1], g, rtol=1e-12)
        np.testing.assert_allclose(t.numpy()[-1], loss, rtol=1e-12)
        q.put((rank, "ok", (lo, hi)))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"fail: {e!r}", None))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_rows():
    from fastoptsolver_b200.multigpu import shard_bounds
    for n in (1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == ["ok", "ok"], res
    assert res[0][2] == (0, 500) and res[1][2] == (500, 1001)


def test_world3_gloo():
    """Three ranks: every descriptor / handle reaches every peer (uneven shard sizes too)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == ["ok"] * 3, res
    assert [r[2] for r in res] == [(0, 333), (333, 667), (667, 1001)]


def test_descriptor_server_only_serves_the_announced_ranks():
    """The abstract-namespace socket is visible to every local process.  A connection that does not
    come from one of the announced process ids (SO_PEERCRED) gets no descriptor and does not use up a
    hand-over; an announced one does."""
    import subprocess
    import warnings
    sys.path.insert(0, ROOT)
    from fastoptsolver_b200 import multigpu
    rd, wr = os.pipe()
    srv = multigpu._FdServer(world=2, timeout=20.0)
    try:
        srv.serve(rd, allowed_pids=[os.getpid()])
        # an intruder: another process of the same user connecting first
        code = ("import socket,sys\n"
                "c=socket.socket(socket.AF_UNIX,socket.SOCK_STREAM);c.settimeout(10)\n"
                f"c.connect({srv.name!r})\n"
                "m,f,_,_=socket.recv_fds(c,16,1)\n"
                "print(len(f))\n")
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
        assert out.stdout.strip() == "0", out
        # the announced rank (this process) still gets its descriptor
        fds = multigpu._receive_fds([srv.name, "unused"], rank=1, timeout=10.0)
        assert fds[0] >= 0 and fds[1] == -1
        os.write(wr, b"k")
        assert os.read(fds[0], 1) == b"k"
        os.close(fds[0])
    finally:
        srv.close()
        os.close(rd)
        os.close(wr)
    assert len(srv.rejected) == 1 and not srv.errors
    # a sharded design reclaimed by the garbage collector (no barrier) warns; an explicit close does not
    from fastoptsolver_b200 import design as D
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        D._destroy(0, {"sharded": True, "explicit": True})
        assert not w
        D._destroy(0, {"sharded": True, "explicit": False})
        assert len(w) == 1 and issubclass(w[0].category, ResourceWarning)
