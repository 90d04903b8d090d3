"""TEST INFRASTRUCTURE ONLY -- numpy model of the device-side design generator
(fastoptsolver_b200/csrc/datagen.cu: synth_kernel).

The product generates its large synthetic designs in HBM: Philox4x32-10 keyed by the seed, counter
= (global row, column group, draw), Box-Muller in double precision, the reference's 5-column recipe
(easy_boston_data.py:23-43) per group, population-standardised.  This file restates that generator
with numpy integer arithmetic so that

* tests can pin the kernel's output against an independent implementation
  (tests/test_cuda_parity.py::test_device_generator_matches_numpy_model), and
* ``bench.py --impl reference`` can build its row sample of the benchmark design on the host
  without loading the product library.

Agreement with the kernel is to rounding, not to the bit: CUDA's log / sincospi and numpy's
log / sin / cos differ in the last ulp, and b's dot product is summed in another order.
Only tests and the reference arm of bench.py may import this module.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_SH = np.uint64(32)
_COEF = np.array([5.0, 0.0, -0.02, -0.05, 1.5])


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on arrays of 32-bit counters held in uint64 (csrc/datagen.cu: philox4x32_10)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> _SH) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> _SH) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & _MASK, n2, p0 & _MASK
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normal2(a, b):
    """Two standard normals from two 32-bit words (csrc/datagen.cu: normal2)."""
    u1 = (a.astype(np.float64) + 1.0) * (1.0 / 4294967296.0)
    u2 = b.astype(np.float64) * (1.0 / 4294967296.0)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = np.pi * (2.0 * u2)
    return rad * np.cos(ang), rad * np.sin(ang)


def _rows_block(r_lo, r_hi, d, seed, noise, rho1, rho2, out_A, out_b, off):
    rows = np.arange(r_lo, r_hi, dtype=np.uint64)
    groups, rest = divmod(d, 5)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c1 = np.sqrt(1.0 - rho1 * rho1)
    c2 = np.sqrt(1.0 - rho2 * rho2)
    lo32 = (rows & _MASK)[:, None]
    hi32 = (rows >> _SH)[:, None]
    A = out_A[r_lo - off: r_hi - off]
    if groups:
        g = np.arange(groups, dtype=np.uint64)[None, :]
        shape = (rows.size, groups)
        X = np.broadcast_to(lo32, shape)
        Y = np.broadcast_to(hi32, shape)
        Z = np.broadcast_to(g, shape)
        ra = philox4x32_10(X, Y, Z, np.zeros(shape, np.uint64), k0, k1)
        rb = philox4x32_10(X, Y, Z, np.ones(shape, np.uint64), k0, k1)
        z0, z1 = normal2(ra[0], ra[1])
        z2, z3 = normal2(ra[2], ra[3])
        z4, _ = normal2(rb[0], rb[1])
        A[:, 0:5 * groups:5] = z0
        A[:, 1:5 * groups:5] = rho1 * z0 + c1 * z1
        A[:, 2:5 * groups:5] = z2
        A[:, 3:5 * groups:5] = rho2 * z2 + c2 * z3
        A[:, 4:5 * groups:5] = z4
    if rest:
        j = (np.arange(rest, dtype=np.uint64) + np.uint64(groups))[None, :]
        shape = (rows.size, rest)
        ra = philox4x32_10(np.broadcast_to(lo32, shape), np.broadcast_to(hi32, shape), np.broadcast_to(j, shape),
                           np.full(shape, 2, np.uint64), k0, k1)
        A[:, 5 * groups:] = normal2(ra[0], ra[1])[0]
    # b = A x_true + noise * N(0,1) on the STORED values (float32 storage rounds first, like the kernel)
    # (row-wise reductions instead of a BLAS product: the result must not depend on how the rows are
    # blocked or threaded)
    if groups:
        dot = (A[:, :5 * groups].astype(np.float64).reshape(rows.size, groups, 5) * _COEF).sum(axis=2).sum(axis=1)
    else:
        dot = np.zeros(rows.size)
    ra = philox4x32_10(lo32[:, 0], hi32[:, 0], np.full(rows.size, 0xFFFFFFFF, np.uint64),
                       np.full(rows.size, 3, np.uint64), k0, k1)
    out_b[r_lo - off: r_hi - off] = dot + noise * normal2(ra[0], ra[1])[0]


def synth_rows(n, d, seed=0, noise_std=1.0, rho1=0.8, rho2=0.9, row0=0, dtype=np.float64, threads=None, block=2048):
    """Rows [row0, row0 + n) of the virtual design ``DeviceDesign.synthetic(..., row0=row0)`` generates:
    returns (A [n, d] of ``dtype``, b [n] float64)."""
    A = np.empty((n, d), dtype=dtype)
    b = np.empty(n, dtype=np.float64)
    threads = threads or len(os.sched_getaffinity(0))
    spans = [(lo, min(lo + block, row0 + n)) for lo in range(row0, row0 + n, block)]

    def work(span):
        _rows_block(span[0], span[1], d, int(seed), float(noise_std), float(rho1), float(rho2), A, b, row0)

    if threads > 1 and len(spans) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, spans))
    else:
        for s in spans:
            work(s)
    return A, b
