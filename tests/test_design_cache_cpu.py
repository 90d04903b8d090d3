"""CPU: the host-side design cache (fastoptsolver_b200/design.py) with a stand-in for the device
upload -- reuse only for the same live array object with unchanged content, no stale hits when
an address is recycled, eviction when the host array dies."""
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
    D._CACHE.clear()
    yield made
    D._CACHE.clear()


def test_reuse_requires_same_object_and_content(fake_upload):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((200, 7))
    b = rng.standard_normal(200)
    d1 = D.as_design(A, b)
    assert D.as_design(A, b) is d1 and len(fake_upload) == 1          # unchanged: reused
    b2 = b.copy()
    assert D.as_design(A, b2) is not d1                                 # other b: new upload
    A[3, 2] += 1.0                                                      # row 3 is not sampled... but b is unchanged
    A[0, 0] += 1.0                                                      # first row IS part of the fingerprint
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
