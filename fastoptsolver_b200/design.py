"""DeviceDesign: the pair (A, b) resident in HBM, plus the (opt-in) cache that lets the drop-in
solvers reuse the device copy of plain numpy arrays.

The reference re-reads its numpy arrays on every call (iterative_solvers.py:133-134), so by
default ``as_design(A, b)`` uploads bare host arrays on every call.  The notebook calls ~19
solver variants per scenario on the same ``A`` ("A residency"): pass a
``DeviceDesign`` in place of ``A`` to keep it resident, or opt into the cache (``FOS_CACHE=1`` /
``set_cache(True)``), which reuses a copy only for the same live array object with a matching
content fingerprint and a matching fresh row sample.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
import zlib

import numpy as np

from . import _lib

_DTYPES = {np.dtype(np.float64): _lib.FOS_F64, np.dtype(np.float32): _lib.FOS_F32}


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class DeviceDesign:
    """Handle to a design matrix and response vector resident on one GPU.

    Quacks enough like the ``A`` argument of the reference solvers (``.shape``,
    ``.dtype``) to be passed wherever they take ``A``; ``b`` may then be ``None``.
    """

    def __init__(self, handle, n, d, dtype, device, keepalive=None):
        self._h = C.c_void_p(handle)
        self.shape = (int(n), int(d))
        self.dtype = np.dtype(np.float64 if dtype == _lib.FOS_F64 else np.float32)
        self.device = device
        self._keepalive = keepalive
        # shared with the finalizer: was the design closed explicitly, does it hold an exchange
        # window that peer ranks read (multigpu.attach sets "sharded")
        self._life = {"explicit": False, "sharded": False}
        self._finalizer = weakref.finalize(self, _destroy, handle, self._life)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_host(cls, A, b, device=0):
        des, upload = cls.begin_from_host(A, b, device=device)
        try:
            upload()
        except Exception:
            des.close()
            raise
        return des

    @classmethod
    def begin_from_host(cls, A, b, device=0):
        """Two-step creation (include/fos.h: fos_design_create_begin / fos_design_upload): returns the
        design with its matrix still undefined and a callable that performs the copy (it may run on
        another thread; ctypes releases the GIL).  Row-sharded callers wire the exchange windows of the
        handle meanwhile (multigpu.sharded_from_host).  Nothing else may use the design before the
        callable has returned."""
        lib = _lib.load()
        A = np.asarray(A)
        if A.ndim != 2:
            raise ValueError(f"A must be 2-D, got shape {A.shape}")
        if A.dtype not in _DTYPES:
            A = A.astype(np.float64)
        n, d = A.shape
        es = A.itemsize
        rs, cs = A.strides[0] // es, A.strides[1] // es
        c_like = (cs == 1 or d == 1) and rs >= d
        f_like = (rs == 1 or n == 1) and cs >= n
        if A.strides[0] % es or A.strides[1] % es or not (c_like or f_like):
            A = np.ascontiguousarray(A)
            rs, cs = d, 1
        elif c_like:
            cs = 1
        else:
            rs = 1
        b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1))
        if b.shape[0] != n:
            raise ValueError(f"b has {b.shape[0]} entries, A has {n} rows")
        out = C.c_void_p()
        _lib.check(lib.fos_design_create_begin(n, d, _DTYPES[A.dtype], device, C.byref(out)))
        des = cls(out.value, n, d, _DTYPES[A.dtype], device)

        def upload(A=A, b=b, rs=rs, cs=cs):      # keeps the (possibly converted) host arrays alive until done
            _lib.check(lib.fos_design_upload(des.handle, _ptr(A), _ptr(b), rs, cs))

        return des, upload

    @classmethod
    def synthetic(cls, n, d, dtype=np.float64, seed=0, noise_std=1.0, rho1=0.8, rho2=0.9, row0=0, device=0):
        """Correlated-column design generated in HBM (csrc/datagen.cu)."""
        lib = _lib.load()
        out = C.c_void_p()
        code = _DTYPES[np.dtype(dtype)]
        _lib.check(lib.fos_design_create_synthetic(n, d, code, seed, noise_std, rho1, rho2, row0, device,
                                                   C.byref(out)))
        return cls(out.value, n, d, code, device)

    @classmethod
    def from_device_pointers(cls, a_ptr, b_ptr, n, d, dtype, lda, device=0, keepalive=None):
        """Borrow device memory (e.g. torch CUDA tensors: pass ``t.data_ptr()``)."""
        lib = _lib.load()
        out = C.c_void_p()
        code = _DTYPES[np.dtype(dtype)]
        _lib.check(lib.fos_design_create_device(C.c_void_p(a_ptr), C.c_void_p(b_ptr), n, d, code, lda, device,
                                                C.byref(out)))
        return cls(out.value, n, d, code, device, keepalive=keepalive)

    # ------------------------------------------------------------------ queries
    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("DeviceDesign has been closed")
        return self._h

    def close(self):
        if self._h is not None:
            self._life["explicit"] = True
            self._finalizer()
            self._h = None

    def download(self, row0=0, rows=None):
        """Rows [row0, row0+rows) back on the host as (A, b) numpy arrays."""
        n, d = self.shape
        rows = n - row0 if rows is None else rows
        A = np.empty((rows, d), dtype=self.dtype)
        b = np.empty(rows, dtype=np.float64)
        _lib.check(_lib.load().fos_design_download(self.handle, row0, rows, _ptr(A), _ptr(b)))
        return A, b

    def column_sums(self, center=None, squared=False):
        """Per-column sum (or centered sum of squares) over this rank's rows, plus the same for b."""
        d = self.shape[1]
        out = np.empty(d)
        b_out = C.c_double()
        cen = None if center is None else np.ascontiguousarray(center, dtype=np.float64)
        _lib.check(_lib.load().fos_design_column_sums(self.handle, None if cen is None else _ptr(cen), int(squared),
                                                      _ptr(out), C.byref(b_out)))
        return out, b_out.value

    def standardize(self, dist=None, group=None, n_total=None):
        """z-score the columns of A and centre b in place on the device (two statistics passes +
        one rewrite pass), like ``datagen.standardize`` on the host.  With ``dist`` the statistics
        are summed over the row-sharded ranks.  Returns (mean, std, b_mean)."""
        n = self.shape[0]

        def allsum(vec):
            if dist is None:
                return vec
            import torch
            t = torch.from_numpy(np.ascontiguousarray(vec)).to(f"cuda:{self.device}")
            dist.all_reduce(t, group=group)
            return t.cpu().numpy()

        n_all = float(allsum(np.array([float(n)]))[0]) if n_total is None else float(n_total)
        s, bs = self.column_sums()
        tot = allsum(np.concatenate([s, [bs]]))
        mean = tot / n_all
        sq, _ = self.column_sums(center=mean, squared=True)
        sq = allsum(sq)
        sd = np.sqrt(sq / n_all)
        sd = np.where(sd > 0, sd, 1.0)
        _lib.check(_lib.load().fos_design_affine(self.handle, _ptr(np.ascontiguousarray(mean[:-1])),
                                                 _ptr(np.ascontiguousarray(sd)), float(mean[-1])))
        return mean[:-1], sd, float(mean[-1])

    def upload_gram(self):
        """Gram matrix accumulated under the host->device upload, if any (include/fos.h,
        fos_design_upload_gram): dict(ptr, state, copy_ms, tail_ms); state 0 none, 1 local rows,
        2 summed over all ranks."""
        g, st, cm, tm = C.c_void_p(), C.c_int(), C.c_float(), C.c_float()
        _lib.check(_lib.load().fos_design_upload_gram(self.handle, C.byref(g), C.byref(st), C.byref(cm), C.byref(tm)))
        return {"ptr": g.value, "state": st.value, "copy_ms": cm.value, "tail_ms": tm.value}

    def comm_info(self):
        """(rank, world) of a row-sharded design; (0, 1) otherwise."""
        r, w = C.c_int(), C.c_int()
        _lib.check(_lib.load().fos_comm_info(self.handle, C.byref(r), C.byref(w)))
        return r.value, w.value

    def set_upload_gram(self, state):
        _lib.check(_lib.load().fos_design_upload_gram_set(self.handle, int(state)))

    def lambda_max(self):
        out = C.c_double()
        _lib.check(_lib.load().fos_design_lambda_max(self.handle, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ operators
    def grad(self, x, alpha2=0.0):
        """(loss, g) with loss = 0.5||Ax-b||^2 + 0.5 a2 ||x||^2, g = A^T(Ax-b) + a2 x."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.shape[1],):
            raise ValueError(f"x must have shape ({self.shape[1]},), got {x.shape}")
        g = np.empty(self.shape[1])
        loss = C.c_double()
        _lib.check(_lib.load().fos_grad(self.handle, _ptr(x), float(alpha2), _ptr(g), C.byref(loss)))
        return loss.value, g

    def objective(self, x, reg_bits, alpha1, alpha2):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.shape[1],):
            raise ValueError(f"x must have shape ({self.shape[1]},), got {x.shape}")
        out = C.c_double()
        _lib.check(_lib.load().fos_objective(self.handle, _ptr(x), reg_bits, float(alpha1), float(alpha2),
                                             C.byref(out)))
        return out.value

    def power_iter(self, v0, n_iter=100, tol=1e-6):
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        L = C.c_double()
        its = C.c_int()
        ms = C.c_float()
        _lib.check(_lib.load().fos_power_iter(self.handle, _ptr(v0), int(n_iter), float(tol), C.byref(L),
                                              C.byref(its), C.byref(ms)))
        return L.value, its.value, ms.value


def _destroy(handle, life=None):
    if life is not None and life.get("sharded") and not life.get("explicit"):
        # The exchange window of a row-sharded design is read by the peers' epilogue kernels: it must
        # only go away after a barrier (multigpu.close).  Garbage collection cannot provide one.
        import warnings
        warnings.warn("a row-sharded DeviceDesign was reclaimed by the garbage collector; its exchange window "
                      "is freed without a barrier and a peer still inside a solve may read freed memory -- "
                      "call fastoptsolver_b200.multigpu.close(design, dist) instead", ResourceWarning, stacklevel=2)
    try:
        _lib.load().fos_design_destroy(C.c_void_p(handle))
    except Exception:
        pass


# ---------------------------------------------------------------------------------- cache
# The reference re-reads its numpy arrays on every call (iterative_solvers.py:133-134, :173), so a
# caller may edit ``A`` in place between two solver calls.  A device copy can only be reused on the
# caller's word that this does not happen: reuse is therefore OPT-IN (``FOS_CACHE=1`` or
# ``set_cache(True)``); by default every call on bare numpy arrays uploads them again, exactly what
# the reference's re-read means.  Callers that want residency without the promise pass a
# ``DeviceDesign`` in place of ``A`` (the sweep driver, the benchmarks).
_CACHE = {}          # key -> (DeviceDesign, fingerprint, weakref to the host matrix)
_CACHE_MAX = 4
_CACHE_ON = None     # None: follow the environment (FOS_CACHE=1); True / False: set_cache()
_FRESH_ROWS = 16     # rows compared with the device copy, freshly drawn, on every cache hit


def set_cache(enabled):
    """Switch reuse of uploaded host arrays on or off for this process (None: follow FOS_CACHE)."""
    global _CACHE_ON
    _CACHE_ON = enabled
    if not cache_enabled():
        clear_cache()


def cache_enabled():
    if os.environ.get("FOS_NO_CACHE") == "1":
        return False
    if _CACHE_ON is not None:
        return bool(_CACHE_ON)
    return os.environ.get("FOS_CACHE") == "1"


def _fingerprint_matrix(A):
    """Content check of A for cache reuse (a full hash of a 32 GB matrix costs more than uploading it
    again): CRC of 65536 elements at pseudo-random positions (fixed seed, so two calls sample the same
    positions whatever the shape -- a fixed stride aliases with the row length), of 64 complete
    pseudo-random rows (every column is covered: ``A[:, j] = 0`` or a rescaled feature is seen), and of
    the first and last rows.  Small matrices (<= 1 MB) are hashed completely."""
    n, d = A.shape
    if A.size * A.itemsize <= (1 << 20):
        return zlib.crc32(np.ascontiguousarray(A).tobytes())
    rng = np.random.default_rng(0x5EED)
    m = 65536
    crc = zlib.crc32(np.ascontiguousarray(A[rng.integers(0, n, m), rng.integers(0, d, m)]).tobytes())
    rows = np.unique(rng.integers(0, n, 64))
    crc = zlib.crc32(np.ascontiguousarray(A[rows]).tobytes(), crc)
    crc = zlib.crc32(np.ascontiguousarray(A[0]).tobytes(), crc)
    crc = zlib.crc32(np.ascontiguousarray(A[-1]).tobytes(), crc)
    return crc


def _fingerprint(A, b):
    """(matrix fingerprint, CRC of ALL of b)."""
    return _fingerprint_matrix(A), zlib.crc32(np.ascontiguousarray(b).tobytes())


def _device_rows_match(des, A):
    """Freshly drawn rows of the host matrix against the same rows of the device copy, bit for bit.
    The positions change from call to call (OS entropy; numpy's legacy global stream, which
    estimate_lipschitz consumes like the reference, is not touched), so an edit the fixed sample
    missed cannot survive repeated calls."""
    n = A.shape[0]
    if not hasattr(des, "download"):
        return True
    for r in np.random.default_rng().integers(0, n, min(n, _FRESH_ROWS)):
        dev, _ = des.download(int(r), 1)
        if dev.tobytes() != np.ascontiguousarray(A[int(r)], dtype=dev.dtype).tobytes():
            return False
    return True


def _matrix_key(A):
    return (A.__array_interface__["data"][0], A.shape, A.strides, A.dtype.str)


def as_design(A, b=None, device=0):
    """Return a DeviceDesign for (A, b).

    Default: bare host arrays are uploaded on every call (the reference re-reads them on every
    call).  With the cache switched on (``FOS_CACHE=1`` / ``set_cache(True)``) a device copy is reused
    for the *same array objects* (identity, still alive) with unchanged shape/strides/dtype, a
    matching content fingerprint and a matching fresh row sample; ``FOS_NO_CACHE=1`` overrides."""
    if isinstance(A, DeviceDesign):
        return A
    A = np.asarray(A)
    if b is None:
        raise ValueError("b is required when A is a host array")
    b = np.asarray(b)
    if not cache_enabled():
        return DeviceDesign.from_host(A, b, device=device)
    key = _matrix_key(A) + (b.__array_interface__["data"][0], b.shape, device)
    fp = _fingerprint(A, b.reshape(-1))
    hit = _CACHE.get(key)
    if (hit is not None and hit[1] == fp and hit[0]._h is not None and hit[2]() is A
            and _device_rows_match(hit[0], A)):
        return hit[0]
    des = DeviceDesign.from_host(A, b, device=device)
    if len(_CACHE) >= _CACHE_MAX:
        old_key = next(iter(_CACHE))
        _CACHE.pop(old_key)      # freed by its finalizer once nobody else holds it
    try:
        ref = weakref.ref(A, lambda _r, k=key: _evict(k))   # host array gone: free the HBM copy too
    except TypeError:            # not weak-referenceable: never reuse
        ref = lambda: None       # noqa: E731
    _CACHE[key] = (des, fp, ref)
    return des


def _evict(key):
    # dropping the entry drops the cache's reference: the device copy is freed by the
    # DeviceDesign finalizer unless the caller still holds the handle
    try:
        _CACHE.pop(key, None)
    except Exception:       # interpreter shutdown: module globals may already be gone
        pass


def find_by_matrix(A, device=0):
    """A cached design whose matrix is this very host array (same object, still alive, same
    content fingerprint; whatever its b), or None.  Lets ``estimate_lipschitz(A)``, which never
    reads b, reuse the copy a solver uploaded."""
    if isinstance(A, DeviceDesign):
        return A
    if not cache_enabled():
        return None
    A = np.asarray(A)
    akey = _matrix_key(A)
    fp_a = None
    for key, (des, fp, ref) in _CACHE.items():
        if key[:4] == akey and key[6] == device and des._h is not None and ref() is A:
            if fp_a is None:
                fp_a = _fingerprint_matrix(A)
            if fp[0] == fp_a and _device_rows_match(des, A):
                return des
    return None


def trim():
    """Release device memory the library keeps between calls (fos_trim: the split workspace of the
    upload-time Gram accumulation)."""
    _lib.check(_lib.load().fos_trim())


def clear_cache():
    for des, _, _r in _CACHE.values():
        des.close()
    _CACHE.clear()
