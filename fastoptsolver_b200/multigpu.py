"""Row-sharded designs across the GPUs of one box: one process per GPU.

A^T(Ay - b) = sum_g A_g^T (A_g y - b_g) over contiguous row blocks, so the only exchange
step of the hot path is one all-reduce of the (d + 2)-vector [A_g^T r_g, ||r_g||^2,
||A_g x - b_g||^2] per pass.  That all-reduce is fused into the epilogue kernel over
peer memory (csrc/epilogue_kernels.cu: peer_exchange): every rank publishes its partial in
an IPC-mapped window, signals its peers over NVLink and sums all windows in rank order, so
all ranks hold bit-identical iterates and the solver state is simply replicated.

torch.distributed is plumbing here: it carries the 64-byte IPC handles once and provides
the barriers around timed regions; the data path never calls NCCL.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .design import DeviceDesign


def shard_bounds(n, rank, world):
    """Contiguous row block [lo, hi) of rank `rank`; blocks differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return (n * rank) // world, (n * (rank + 1)) // world


def exchange_bytes(payload: bytes, dist, group=None):
    """All-gather one fixed-size byte string per rank (works on gloo and nccl)."""
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, payload, group=group)
    if any(len(o) != len(payload) for o in out):
        raise RuntimeError("ranks published handles of different sizes")
    return out


def exchange_fds(my_fd: int, dist, group=None, timeout=60.0):
    """Hand one file descriptor per rank to every other rank of the box (unix socket in the abstract
    namespace, SCM_RIGHTS).  Returns a list with, for every peer rank, a descriptor valid in THIS
    process (own entry -1); the caller closes them."""
    import os
    import socket
    import threading
    import uuid
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    name = f"\0fos-b200-{os.getpid()}-{uuid.uuid4().hex}"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.settimeout(timeout)
    srv.bind(name)
    srv.listen(world)
    names = [None] * world
    dist.all_gather_object(names, name, group=group)
    errors = []

    def serve():
        try:
            for _ in range(world - 1):
                conn, _addr = srv.accept()
                with conn:
                    socket.send_fds(conn, [b"w"], [my_fd])
        except Exception as e:  # reported by the receiving side as a missing descriptor
            errors.append(e)

    t = threading.Thread(target=serve, daemon=True)
    t.start()
    fds = [-1] * world
    try:
        for r in range(world):
            if r == rank:
                continue
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.settimeout(timeout)
                c.connect(names[r])
                _msg, got, _flags, _addr = socket.recv_fds(c, 16, 1)
                if not got:
                    raise RuntimeError(f"rank {r} sent no descriptor")
                fds[r] = got[0]
    finally:
        t.join(timeout)
        srv.close()
    if errors:
        raise RuntimeError(f"could not hand the window descriptor to every peer: {errors[0]}")
    return fds


def _all_ok(ok: bool, dist, group):
    flags = [None] * dist.get_world_size(group)
    dist.all_gather_object(flags, bool(ok), group=group)
    return all(flags)


def _attach_vmm(des, dist, group):
    """Windows as cuMemCreate allocations shared by file descriptor (include/fos.h,
    fos_comm_window_alloc_fd).  Returns False -- with nothing left behind on any rank -- if some
    rank cannot do it; the caller then falls back to the cudaIpc windows."""
    import os
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    fd = C.c_int(-1)
    ok = lib.fos_comm_window_alloc_fd(des.handle, rank, world, C.byref(fd)) == _lib.FOS_OK
    if not _all_ok(ok, dist, group):
        if ok:
            _lib.check(lib.fos_comm_window_free(des.handle))
        return False
    fds, err = [-1] * world, None
    try:
        fds = exchange_fds(fd.value, dist, group)
        arr = (C.c_int * world)(*fds)
        ok = lib.fos_comm_attach_fd(des.handle, arr, world) == _lib.FOS_OK
    except Exception as e:      # socket trouble: treated like an import failure
        ok, err = False, e
    finally:
        for f in fds:
            if f >= 0:
                os.close(f)
    if not _all_ok(ok, dist, group):
        # some rank could not import: nobody has signalled a peer yet, so the windows can go
        raise RuntimeError("sharing the exchange windows by file descriptor failed half-way"
                           + (f": {err}" if err else "") + "; set FOS_COMM=ipc to use cudaIpc windows")
    return True


def attach(des: DeviceDesign, dist, group=None):
    """Allocate this rank's exchange window, hand it to the peers, map theirs."""
    import os
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise ValueError("at most 8 ranks (one NVSwitch box) are supported")
    if os.environ.get("FOS_COMM", "vmm") != "ipc" and _attach_vmm(des, dist, group):
        dist.barrier(group)      # nobody signals a peer before every window is mapped
        return group
    buf = C.create_string_buffer(64)
    _lib.check(lib.fos_comm_window_alloc(des.handle, rank, world, buf))
    handles = exchange_bytes(buf.raw, dist, group)
    blob = C.create_string_buffer(b"".join(handles), 64 * world)
    _lib.check(lib.fos_comm_attach(des.handle, blob, world))
    dist.barrier(group)          # nobody signals a peer before every window is mapped
    return group


def sharded_from_host(A_local, b_local, dist, group=None, device=None):
    """Upload this rank's row block and wire it to its peers.  The result is accepted by the
    drop-in solvers wherever they take ``A`` (pass ``b=None``)."""
    if device is None:
        import torch
        device = torch.cuda.current_device()
    des = DeviceDesign.from_host(A_local, b_local, device=device)
    attach(des, dist, group)
    allreduce_upload_gram(des, dist, group)
    return des


def allreduce_upload_gram(des: DeviceDesign, dist, group=None):
    """If every rank accumulated the Gram matrix of its rows under the upload (include/fos.h,
    fos_design_upload_gram), sum them over the ranks in place -- d^2 doubles, once, a plain library
    all-reduce -- so that ``estimate_lipschitz`` iterates on G instead of streaming A; if any
    rank has none, all ranks discard theirs and keep the streaming power iteration."""
    import torch
    info = des.upload_gram()
    if "nccl" not in str(dist.get_backend(group)):
        # no device collective on this group (same answer on every rank): keep streaming
        if info["state"] != 0:
            des.set_upload_gram(0)
        return False
    dev = f"cuda:{des.device}"
    flag = torch.tensor([1 if info["state"] == 1 else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if info["state"] != 0:
            des.set_upload_gram(0)
        return False
    d = des.shape[1]
    arr = type("_G", (), {"__cuda_array_interface__": {"shape": (d, d), "typestr": "<f8", "data": (int(info["ptr"]), False),
                                                       "version": 3, "strides": None}})()
    G = torch.as_tensor(arr, device=dev)
    dist.all_reduce(G, group=group)
    torch.cuda.synchronize(G.device)
    des.set_upload_gram(2)
    return True


def sharded_synthetic(n_total, d, dist, group=None, device=None, dtype=np.float64, **scenario):
    """Each rank generates its row block of the same virtual n_total x d design in HBM."""
    if device is None:
        import torch
        device = torch.cuda.current_device()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(n_total, rank, world)
    des = DeviceDesign.synthetic(hi - lo, d, dtype, row0=lo, device=device, **scenario)
    attach(des, dist, group)
    return des


def close(des: DeviceDesign, dist, group=None):
    """Free a sharded design.  The exchange window of a rank is read by its peers' epilogue
    kernels, so every rank must be done with its last solve before any window goes away:
    barrier first, then close."""
    dist.barrier(group)
    des.close()
