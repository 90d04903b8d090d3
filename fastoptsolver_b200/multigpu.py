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


def attach(des: DeviceDesign, dist, group=None):
    """Allocate this rank's exchange window, swap IPC handles, map the peers."""
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise ValueError("at most 8 ranks (one NVSwitch box) are supported")
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
