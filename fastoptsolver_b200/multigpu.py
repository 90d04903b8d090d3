"""Row-sharded designs across the GPUs of one box: one process per GPU.

A^T(Ay - b) = sum_g A_g^T (A_g y - b_g) over contiguous row blocks, so the only exchange
step of the hot path is one all-reduce of the (d + 2)-vector [A_g^T r_g, ||r_g||^2,
||A_g x - b_g||^2] per pass.  That all-reduce is fused into the epilogue kernel over
peer memory (csrc/epilogue_kernels.cu: peer_exchange): every rank publishes its partial in
a window mapped into every peer, signals its peers over NVLink and sums all windows in rank order, so
all ranks hold bit-identical iterates and the solver state is simply replicated.

torch.distributed is plumbing here: it wires the windows once (socket names / handles and an
agreement flag ride on two object all-gathers) and provides the barriers around timed regions;
the data path never calls NCCL.
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


def _gather(obj, dist, group):
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def _all_ok(ok: bool, dist, group):
    return all(_gather(bool(ok), dist, group))


def _peer_credentials(conn):
    """(pid, uid, gid) of the process at the other end of a unix socket (SO_PEERCRED)."""
    import socket
    import struct
    raw = conn.getsockopt(socket.SOL_SOCKET, socket.SO_PEERCRED, struct.calcsize("3i"))
    return struct.unpack("3i", raw)


class _FdServer:
    """Hands this rank's descriptor to every peer rank that connects (unix socket in the abstract
    namespace, SCM_RIGHTS).  Abstract names are visible to every local process, so a connection
    only gets the descriptor if the kernel says (SO_PEERCRED) that it comes from this user AND from
    one of the process ids the ranks published over the process group; anything else is dropped
    and does not use up one of the world-1 hand-overs."""

    def __init__(self, world, timeout=60.0):
        import os
        import socket
        import uuid
        self.world, self.timeout = world, timeout
        self.name = f"\0fos-b200-{os.getpid()}-{uuid.uuid4().hex}"
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.sock.settimeout(timeout)
        self.sock.bind(self.name)
        self.sock.listen(world + 8)
        self.errors = []
        self.rejected = []
        self.thread = None

    def serve(self, fd, allowed_pids=None):
        import os
        import socket
        import threading
        allowed = None if allowed_pids is None else set(int(p) for p in allowed_pids)

        def run():
            try:
                served = 0
                while served < self.world - 1:
                    conn, _addr = self.sock.accept()
                    with conn:
                        pid, uid, _gid = _peer_credentials(conn)
                        if uid != os.getuid() or (allowed is not None and pid not in allowed):
                            self.rejected.append((pid, uid))
                            continue
                        socket.send_fds(conn, [b"w"], [fd])
                        served += 1
            except Exception as e:      # the receiving side reports the missing descriptor
                self.errors.append(e)

        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()

    def close(self):
        if self.thread is not None:
            self.thread.join(self.timeout)
        self.sock.close()


def _receive_fds(names, rank, timeout=60.0):
    import socket
    fds = [-1] * len(names)
    for r, name in enumerate(names):
        if r == rank:
            continue
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
            c.settimeout(timeout)
            c.connect(name)
            _msg, got, _flags, _addr = socket.recv_fds(c, 16, 1)
            if not got:
                raise RuntimeError(f"rank {r} sent no descriptor")
            fds[r] = got[0]
    return fds


def exchange_fds(my_fd: int, dist, group=None, timeout=60.0):
    """Hand one file descriptor per rank to every other rank of the box.  Returns a list with, for
    every peer rank, a descriptor valid in THIS process (own entry -1); the caller closes them."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    import os
    srv = _FdServer(world, timeout)
    try:
        pairs = _gather((srv.name, os.getpid()), dist, group)
        names = [n for n, _ in pairs]
        srv.serve(my_fd, allowed_pids=[p for _, p in pairs])
        fds = _receive_fds(names, rank, timeout)
    finally:
        srv.close()
    if srv.errors:
        raise RuntimeError(f"could not hand the descriptor to every peer: {srv.errors[0]}")
    return fds


def share_windows(alloc_fd, attach_fd, free_window, alloc_ipc, attach_ipc, dist, group=None, piggyback=None,
                  prefer_vmm=True):
    """The collective part of wiring the exchange windows, independent of the library calls (which
    are passed in, so that the CPU tests can drive it with stand-ins):

      alloc_fd()        -> descriptor of this rank's VMM window, or None if unsupported here
      attach_fd(fds)    -> True once the peers' windows (descriptors valid in this process) are mapped
      free_window()     -> drop a window that was allocated but is not attached
      alloc_ipc()       -> 64-byte cudaIpc handle of a freshly allocated window
      attach_ipc(blobs) -> map the peers' windows from their handles

    Two object all-gathers in total: (VMM ok?, socket name) before and (attached ok?, piggyback)
    after; the second doubles as the barrier "nobody signals a peer before every window is mapped".
    Returns (kind, [piggyback of every rank])."""
    import os
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if prefer_vmm:
        srv = _FdServer(world)
        fds = [-1] * world
        try:
            fd = alloc_fd()
            first = _gather((fd is not None, srv.name, os.getpid()), dist, group)
            if all(ok for ok, _, _ in first):
                ok, err = True, None
                try:
                    srv.serve(fd, allowed_pids=[p for _, _, p in first])
                    fds = _receive_fds([name for _, name, _ in first], rank)
                    ok = bool(attach_fd(fds))
                except Exception as e:      # socket trouble counts like an import failure
                    ok, err = False, e
                second = _gather((ok and not srv.errors, piggyback), dist, group)
                if not all(ok for ok, _ in second):
                    raise RuntimeError("sharing the exchange windows by file descriptor failed half-way"
                                       + (f": {err}" if err else "") + "; set FOS_COMM=ipc to use cudaIpc windows")
                return "vmm", [p for _, p in second]
            if fd is not None:
                free_window()
        finally:
            srv.close()
            for f in fds:
                if f >= 0:
                    os.close(f)
    handles = _gather(alloc_ipc(), dist, group)
    if any(len(hd) != 64 for hd in handles):
        raise RuntimeError("ranks published handles of different sizes")
    attach_ipc(handles)
    return "ipc", _gather(piggyback, dist, group)


def attach(des: DeviceDesign, dist, group=None, piggyback=None):
    """Allocate this rank's exchange window, hand it to the peers, map theirs.  Returns the
    ``piggyback`` objects of all ranks (they ride on the closing all-gather)."""
    import os
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise ValueError("at most 8 ranks (one NVSwitch box) are supported")

    def alloc_fd():
        fd = C.c_int(-1)
        ok = lib.fos_comm_window_alloc_fd(des.handle, rank, world, C.byref(fd)) == _lib.FOS_OK
        return fd.value if ok else None

    def attach_fd(fds):
        return lib.fos_comm_attach_fd(des.handle, (C.c_int * world)(*fds), world) == _lib.FOS_OK

    def alloc_ipc():
        buf = C.create_string_buffer(64)
        _lib.check(lib.fos_comm_window_alloc(des.handle, rank, world, buf))
        return buf.raw

    def attach_ipc(handles):
        blob = C.create_string_buffer(b"".join(handles), 64 * world)
        _lib.check(lib.fos_comm_attach(des.handle, blob, world))

    # Persistent solve kernel (push-model exchange) or launch pairs (pull-model): the same on every rank.
    # The candidate flag rides on the closing all-gather of the wiring.
    ok = C.c_int(0)
    _lib.check(lib.fos_design_solve_kernel_ok(des.handle, world, C.byref(ok)))
    kind, extras = share_windows(alloc_fd, attach_fd, lambda: _lib.check(lib.fos_comm_window_free(des.handle)),
                                 alloc_ipc, attach_ipc, dist, group, (bool(ok.value), piggyback),
                                 prefer_vmm=os.environ.get("FOS_COMM", "vmm") != "ipc")
    if not all(f for f, _ in extras):
        _lib.check(lib.fos_design_solve_kernel_disable(des.handle))
    extras = [p for _, p in extras]
    des.comm_kind = kind
    if hasattr(des, "_life"):
        des._life["sharded"] = world > 1      # reclaiming it without multigpu.close() warns
    return extras


def sharded_from_host(A_local, b_local, dist, group=None, device=None):
    """Upload this rank's row block and wire it to its peers.  The result is accepted by the
    drop-in solvers wherever they take ``A`` (pass ``b=None``)."""
    if device is None:
        import torch
        device = torch.cuda.current_device()
    import threading
    # The PCIe copy runs on a side thread (ctypes releases the GIL) while this thread wires the exchange
    # windows of the same handle: socket hand-off of the descriptors + two object all-gathers, ~0.1 s
    # that used to FOLLOW the copy (the windows do not depend on the data).
    des, upload = DeviceDesign.begin_from_host(A_local, b_local, device=device)
    failure = []

    def run():
        try:
            upload()
        except BaseException as e:      # re-raised on the caller's thread below
            failure.append(e)

    th = threading.Thread(target=run, name="fos-upload")
    th.start()
    wiring_error = None
    try:
        attach(des, dist, group)
    except BaseException as e:
        wiring_error = e
    th.join()
    if wiring_error is not None:        # the collective itself broke: nothing to agree on any more
        des.close()
        raise wiring_error
    # one more (tiny) all-gather: did every rank's copy succeed, and does every rank hold a Gram matrix
    have = (not failure) and des.upload_gram()["state"] == 1
    everyone = _gather((not failure, have), dist, group)
    if not all(ok for ok, _ in everyone):
        des.close()
        raise failure[0] if failure else RuntimeError("the upload failed on another rank")
    # Every rank accumulated the Gram matrix of its own rows under the upload (include/fos.h,
    # fos_design_upload_gram)?  Then estimate_lipschitz iterates on them -- local product, then the
    # same fused peer-memory all-reduce of the d-vector as a streaming pass; the matrices are never
    # summed.  All ranks must take the same branch: if any rank has none (shard below the size
    # threshold, workspace allocation failed), all discard theirs and keep streaming.
    if all(h for _, h in everyone):
        des.set_upload_gram(2)
    elif have:
        des.set_upload_gram(0)
    return des


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
    des.close()     # explicit: no ResourceWarning from the finalizer
