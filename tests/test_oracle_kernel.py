"""Pins the ArrayKernel oracle -- the reference's own native kernel compiled from /root/reference against the
shim (oracle/_ref/libsst_ref.so) -- to the reference's Java-side known answers. CPU only."""
import numpy as np
import pytest

import oracle
from conftest import golden
from kernel_kats import check_all

ref = oracle.ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref/libsst_ref.so not built (needs /root/reference)")


def test_call_level_kats():
    assert check_all(ref) >= 50  # ArrayKernelTest.java:61-378


def far(dims):
    s = [1] * len(dims)
    for i in range(len(dims) - 2, -1, -1):
        s[i] = s[i + 1] * dims[i + 1]
    return s


def test_rrop_sum_kat():
    # RealArrayTest.java:726-758: a.rSum(1) on 3x4x5 FAR
    recs = golden("RealArrayTest")["testRrOps"]
    a, exp = recs[0], recs[1]
    src = np.array(a["values"])
    dst = np.zeros(len(exp["values"]))
    ref.rrOp(0, src, a["dims"], far(a["dims"]), dst, exp["dims"], far(exp["dims"]), 1)
    assert np.array_equal(dst, np.array(exp["values"]))


def test_mul_kat():
    # MatrixTest.java:69-90: a (4x3) . a^T; :92-123: complex 3x5 . 5x3
    recs = golden("MatrixTest")["testMMul"]
    a, exp = recs[0], recs[1]
    at = np.ascontiguousarray(np.array(a["values"]).reshape(4, 3).T).ravel()
    dst = np.zeros(16)
    ref.mul(np.array(a["values"]), at, 4, 4, dst, False)
    assert np.max(np.abs(dst - np.array(exp["values"]))) < 1e-8
    c1, c2, cexp = recs[2], recs[3], recs[4]
    dst = np.zeros(18)
    ref.mul(np.array(c1["values"]), np.array(c2["values"]), 3, 3, dst, True)
    assert np.max(np.abs(dst - np.array(cexp["values"]))) < 1e-8


def test_error_text():
    with pytest.raises(oracle.refkernel.RefKernelError, match="Invalid dimensions and/or strides"):
        ref.rrOp(0, np.zeros(6), [2, 3], [1, 1], np.zeros(2), [2, 1], [1, 1], 1)
    with pytest.raises(oracle.refkernel.RefKernelError, match="Operation type not recognized"):
        ref.ruOp(99, 0.0, np.zeros(3))


def test_invert_and_find_reference_properties():
    # MatrixTest.java:227-242: (rnd + 2I) . inverse == I, sum |.| < 1e-8; "Matrix is singular" on a zero pivot
    rng = np.random.default_rng(5)
    n = 48
    a = rng.uniform(0, 1, (n, n)) + 2 * np.eye(n)
    inv = np.zeros(n * n)
    ref.invert(a.ravel().copy(), inv, n)
    assert np.abs(a @ inv.reshape(n, n) - np.eye(n)).sum() < 1e-8
    with pytest.raises(oracle.refkernel.RefKernelError, match="Matrix is singular"):
        ref.invert(np.zeros(9), np.zeros(9), 3)
    # MatrixTest.java:186-203: A == U S V^T (sum |.| < 1e-8), singular values = numpy's
    a = rng.uniform(0, 1, (40, 30))
    u, sv, v = np.zeros(1200), np.zeros(30), np.zeros(900)
    ref.svd(a.ravel().copy(), 30, 1, u, sv, v, 40, 30)
    assert np.abs(a - u.reshape(40, 30) @ np.diag(sv) @ v.reshape(30, 30).T).sum() < 1e-8
    assert np.max(np.abs(sv - np.linalg.svd(a, compute_uv=False))) < 1e-12
    # IndexOps.java:44-91 on a riOp-shaped array: positions, then -1 padding
    v = np.array([3, 1, -1, -1, 0, 2, 5, -1], dtype=np.int32)
    assert list(ref.find(v, [2, 4], [4, 1], [1, -1])) == [0, 2, 5]
    assert list(ref.find(v, [2, 4], [4, 1], [0, -1])) == [3, 1]
    assert list(ref.find(v, [2, 4], [4, 1], [-1, 3])) == []
    with pytest.raises(oracle.refkernel.RefKernelError, match="Invalid index"):
        ref.find(v, [2, 4], [4, 1], [2, -1])
