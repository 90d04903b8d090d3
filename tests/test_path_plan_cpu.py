"""CPU: the host-side schedule choice of the batched path iteration (csrc/gram_kernels.cu: path_plan, reached
through the C ABI's fos_debug_path_plan -- no GPU involved): padded penalty count, tile shape, stream-K or one
tile per CTA.  The kernels rely on the invariants checked here (grid sizes are plain divisions)."""
import ctypes as C

import pytest

from fastoptsolver_b200 import _lib


def plan(d, n_lambda, sms=148):
    lp, pm, tn, sk, nt = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    _lib.check(_lib.load().fos_debug_path_plan(d, n_lambda, sms, C.byref(lp), C.byref(pm), C.byref(tn), C.byref(sk), C.byref(nt)))
    return {"Lpad": lp.value, "pm": pm.value, "tn": tn.value, "sk": bool(sk.value), "tiles": nt.value}


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    monkeypatch.delenv("FOS_PATH_SK", raising=False)
    monkeypatch.delenv("FOS_PATH_TN", raising=False)


def test_invariants_over_a_grid_of_shapes():
    for d in (128, 256, 384, 1024, 2048, 4096):
        for n_lambda in list(range(1, 70)) + [95, 96, 97, 128, 129, 160, 161, 200, 256, 257, 500, 1000]:
            for sms in (148, 132, 16):
                p = plan(d, n_lambda, sms)
                assert p["Lpad"] >= n_lambda and p["Lpad"] - n_lambda < 64
                assert p["tn"] in (32, 64, 128) and p["Lpad"] % p["tn"] == 0
                assert p["pm"] in (32, 64, 128) and d % p["pm"] == 0
                if p["sk"]:
                    assert p["pm"] == 128
                    assert p["tiles"] == (d // 128) * (p["Lpad"] // p["tn"])
                    # enough k-steps for every CTA to get a non-empty range, and a tile never has more contributors
                    # than the fix-up was designed for
                    assert p["tiles"] * (d // 16) >= sms
                else:
                    assert p["tn"] == 64 and p["Lpad"] % 64 == 0
                if p["tn"] == 32:
                    assert p["sk"] and p["Lpad"] <= 160 and p["Lpad"] == (n_lambda + 31) // 32 * 32
                    assert p["Lpad"] < (n_lambda + 63) // 64 * 64, "32-wide tiles only when they remove padding"


def test_the_shapes_the_benchmarks_use():
    # config 5 on one GPU: 256 penalties, d = 4096 -> stream-K, 128 x 128 tiles, 64 of them
    assert plan(4096, 256) == {"Lpad": 256, "pm": 128, "tn": 128, "sk": True, "tiles": 64}
    # the 32 penalties a rank holds on 8 GPUs: 128 x 32 tiles, no padding
    assert plan(4096, 32) == {"Lpad": 32, "pm": 128, "tn": 32, "sk": True, "tiles": 32}
    # 33 penalties would need 64 either way
    assert plan(4096, 33)["tn"] == 64 and plan(4096, 33)["Lpad"] == 64
    # a small problem keeps one tile per CTA and shrinks the row tile so that every SM has one
    small = plan(384, 64)
    assert not small["sk"] and small["tn"] == 64 and small["pm"] == 32


def test_environment_switches(monkeypatch):
    monkeypatch.setenv("FOS_PATH_SK", "0")
    p = plan(4096, 256)
    assert not p["sk"] and p["tn"] == 64 and p["Lpad"] == 256
    monkeypatch.delenv("FOS_PATH_SK")
    monkeypatch.setenv("FOS_PATH_TN", "64")
    assert plan(4096, 256)["tn"] == 64 and plan(4096, 32) == {"Lpad": 64, "pm": 128, "tn": 64, "sk": True, "tiles": 32}
    monkeypatch.setenv("FOS_PATH_TN", "128")
    assert plan(4096, 256)["tn"] == 128 and plan(4096, 70)["tn"] == 128      # 70 -> 128 padded either way
    assert plan(4096, 32)["tn"] == 64                                         # 128 does not divide 64


def test_bad_arguments_are_refused():
    lib = _lib.load()
    assert lib.fos_debug_path_plan(100, 8, 148, None, None, None, None, None) != 0
    assert lib.fos_debug_path_plan(4096, 0, 148, None, None, None, None, None) != 0
