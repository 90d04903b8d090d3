"""CPU: the host-side design cache (fastoptsolver_b200/design.py) with a stand-in for the device
upload -- off by default (the reference re-reads its arrays on every call); when opted into, reuse
only for the same live array object with unchanged content (fingerprint + fresh row sample against
the device copy), no stale hits when an address is recycled, eviction when the host array dies."""
import gc

import numpy as np
import pytest

from fastoptsolver_b200 import design as D


class _FakeDesign:
    def __init__(self, A, b, device):
        self.shape = A.shape
        self.device = device
        self._h = object()
        self.closed = False

    def close(self):
        self._h = None
        self.closed = True


@pytest.fixture
def fake_upload(monkeypatch):
    made = []

    def from_host(A, b, device=0):
        d = _FakeDesign(A, b, device)
        made.append(d)
        return d

    monkeypatch.setattr(D.DeviceDesign, "from_host", staticmethod(from_host))
    monkeypatch.delenv("FOS_NO_CACHE", raising=False)
    D._CACHE.clear()
    D.set_cache(True)
    yield made
    D.set_cache(None)
    D._CACHE.clear()


def test_reuse_requires_same_object_and_content(fake_upload):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((200, 7))
    b = rng.standard_normal(200)
    d1 = D.as_design(A, b)
    assert D.as_design(A, b) is d1 and len(fake_upload) == 1          # unchanged: reused
    b2 = b.copy()
    assert D.as_design(A, b2) is not d1                                 # other b: new upload
    A[3, 2] += 1.0                                                      # small matrices are hashed completely
    d3 = D.as_design(A, b)
    assert d3 is not d1 and len(fake_upload) == 3
    assert D.as_design(A, b) is d3
    b[5] = 9.0                                                          # all of b is fingerprinted
    assert D.as_design(A, b) is not d3


def test_find_by_matrix_never_returns_a_recycled_address(fake_upload):
    """estimate_lipschitz(A) looks a design up by matrix only.  A new array that happens to sit at
    the address (and have the shape) of a dead one must not get the dead one's device copy."""
    A = np.arange(60, dtype=np.float64).reshape(12, 5)
    b = np.zeros(12)
    d1 = D.as_design(A, b)
    assert D.find_by_matrix(A) is d1
    view_same_memory = np.ndarray(A.shape, dtype=A.dtype, buffer=A.data)   # same address, another object
    assert D.find_by_matrix(view_same_memory) is None
    key = next(iter(D._CACHE))
    # simulate the recycled address: entry whose weak reference points at nothing any more
    des, fp, _ref = D._CACHE[key]
    D._CACHE[key] = (des, fp, lambda: None)
    assert D.find_by_matrix(A) is None
    # and a changed matrix under the same object is not found either
    D._CACHE[key] = (des, fp, _ref)
    A[0, 0] = -1.0
    assert D.find_by_matrix(A) is None


def test_entry_goes_when_the_host_array_dies(fake_upload):
    A = np.ones((50, 3))
    b = np.ones(50)
    D.as_design(A, b)
    assert len(D._CACHE) == 1
    del A
    gc.collect()
    assert len(D._CACHE) == 0


def test_cache_is_bounded_and_device_designs_pass_through(fake_upload):
    keep = []
    for i in range(D._CACHE_MAX + 3):
        A = np.full((20, 4), float(i))
        b = np.zeros(20)
        keep.append((A, b))
        D.as_design(A, b)
    assert len(D._CACHE) <= D._CACHE_MAX
    fake = fake_upload[0]
    real = object.__new__(D.DeviceDesign)        # isinstance check only; never dereferenced
    assert D.as_design(real) is real and D.find_by_matrix(real) is real
    with pytest.raises(ValueError):
        D.as_design(np.ones((3, 3)))             # host matrix without b
    assert fake is fake_upload[0]


def test_no_cache_env(fake_upload, monkeypatch):
    monkeypatch.setenv("FOS_NO_CACHE", "1")
    A = np.ones((10, 2))
    b = np.ones(10)
    assert D.as_design(A, b) is not D.as_design(A, b)
    assert len(D._CACHE) == 0


def test_cache_is_off_by_default(fake_upload, monkeypatch):
    """The reference re-reads A on every call: without the opt-in every call uploads."""
    D.set_cache(None)
    monkeypatch.delenv("FOS_CACHE", raising=False)
    A = np.ones((10, 2))
    b = np.ones(10)
    assert D.as_design(A, b) is not D.as_design(A, b)
    assert len(D._CACHE) == 0 and D.find_by_matrix(A) is None
    monkeypatch.setenv("FOS_CACHE", "1")
    d1 = D.as_design(A, b)
    assert D.as_design(A, b) is d1 and D.find_by_matrix(A) is d1
    monkeypatch.setenv("FOS_NO_CACHE", "1")                  # the override wins
    assert D.as_design(A, b) is not d1


@pytest.mark.parametrize("shape", [(1 << 16, 64), (4000, 4096), (62500, 48)])
def test_fingerprint_sees_column_edits_whatever_the_shape(shape):
    """Round-1 sampled with a fixed stride n*d // 65536: for n a multiple of 65536 that hit column 0
    only.  A zeroed / rescaled feature must change the fingerprint for any shape."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal(shape, dtype=np.float32) if shape[1] == 4096 else rng.standard_normal(shape)
    fp0 = D._fingerprint_matrix(A)
    assert D._fingerprint_matrix(A) == fp0
    for j in (1, shape[1] // 2, shape[1] - 1):
        keep = A[:, j].copy()
        A[:, j] = 0.0
        assert D._fingerprint_matrix(A) != fp0, f"zeroed column {j} not seen"
        A[:, j] = keep * 1.5
        assert D._fingerprint_matrix(A) != fp0, f"rescaled column {j} not seen"
        A[:, j] = keep
    assert D._fingerprint_matrix(A) == fp0


def test_fresh_row_sample_catches_what_the_fingerprint_missed(fake_upload, monkeypatch):
    """An edit confined to rows outside the fixed sample slips past the fingerprint; the freshly
    drawn rows compared with the device copy catch it within a few calls."""
    rng = np.random.default_rng(2)
    A = rng.standard_normal((40000, 8))
    b = rng.standard_normal(40000)
    snapshot = A.copy()

    def download(self, row0=0, rows=None):
        return snapshot[row0: row0 + rows].copy(), None

    monkeypatch.setattr(_FakeDesign, "download", download, raising=False)
    d1 = D.as_design(A, b)
    assert D.as_design(A, b) is d1
    fp0 = D._fingerprint_matrix(A)
    # rows the fixed sample looks at (same generator, same draws as design._fingerprint_matrix)
    g = np.random.default_rng(0x5EED)
    seen = set(g.integers(0, 40000, 65536).tolist())
    g.integers(0, 8, 65536)
    seen |= set(g.integers(0, 40000, 64).tolist()) | {0, 39999}
    edit = np.array([r for r in range(40000) if r not in seen][:3000])
    assert edit.size == 3000
    A[edit] += 1.0
    assert D._fingerprint_matrix(A) == fp0          # the fixed sample cannot see this edit ...
    hits = sum(D.as_design(A, b) is d1 for _ in range(40))
    assert hits < 40                                # ... the freshly drawn rows do (3000 of 40000 rows, 16 per call)
