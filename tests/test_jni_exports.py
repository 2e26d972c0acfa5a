"""Every JNI export of shared_b200/csrc/jni/jni_glue.cpp CALLED on the CPU (tests/jni_harness): the Java-side arguments
are {kind, data, len} array records of the functional JNI model, the C ABI behind the glue is served by the reference's
own native code, and the result must equal a direct call of the reference with the Java argument order. What this pins
down is the glue's marshalling -- which argument lands in which C-ABI parameter (lr / rc, strides, lengths, flags), which
arrays are written back (pin modes), what is returned -- the one layer no GPU test reaches and no JDK is here to run."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from kernel_kats import OPCODES as OP

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ref = oracle.ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref/libsst_ref.so missing")
AK = "Java_org_sharedx_cuda_CudaArrayKernel_"


class JObj(ctypes.Structure):  # struct _jobject of oracle/ref/jni.h
    _fields_ = [("kind", ctypes.c_int), ("data", ctypes.c_void_p), ("len", ctypes.c_int), ("owned", ctypes.c_int)]


def J(a):
    """A Java double[] / int[] over a numpy array."""
    kind = {np.dtype(np.float64): 1, np.dtype(np.int32): 2}[a.dtype]
    obj = JObj(kind, a.ctypes.data, a.size, 0)
    obj.keep = a  # the record only holds the address: keep temporaries alive for as long as the record lives
    return ctypes.byref(obj)


def I(values):
    return np.ascontiguousarray(values, dtype=np.int32)


class JavaException(RuntimeError):
    pass


class Glue:
    def __init__(self, lib):
        self.lib = lib
        lib.jh_env.restype = ctypes.c_void_p
        self.env = ctypes.c_void_p(lib.jh_env())
        lib.jh_onload(ctypes.create_string_buffer(1024), 1024, ctypes.byref(ctypes.c_int()))

    def call(self, name, *args, restype=None):
        fn = getattr(self.lib, name)
        fn.restype = restype
        out = fn(self.env, None, *args)
        err = ctypes.create_string_buffer(512)
        if self.lib.jh_take_exception(err, 512):
            raise JavaException(err.value.decode())
        return out

    def array_result(self, ptr, dtype):
        obj = ctypes.cast(ptr, ctypes.POINTER(JObj)).contents
        n = obj.len
        out = np.frombuffer((ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(obj.data), dtype=dtype).copy() \
            if n else np.zeros(0, dtype=dtype)
        self.lib.jh_free_array(ctypes.c_void_p(ptr))
        return out


@pytest.fixture(scope="module")
def g():
    if shutil.which("g++") is None or shutil.which("make") is None:
        pytest.skip("no g++ / make")
    res = subprocess.run(["make", "-s", "-k", "-C", os.path.join(HERE, "jni_harness"), "all"], capture_output=True, text=True)
    path = os.path.join(HERE, "jni_harness", "_build", "libjni_harness.so")
    if not os.path.exists(path):
        pytest.fail("harness build failed:\n" + res.stdout[-3000:] + res.stderr[-3000:])
    lib = ctypes.CDLL(path)
    oracle.lib()  # makes sure oracle/libsst_oracle.so is built
    got = lib.jh_set_backends(os.path.join(ROOT, "oracle", "_ref", "libsst_ref.so").encode(),
                              os.path.join(ROOT, "oracle", "libsst_oracle.so").encode())
    assert got == 3
    return Glue(lib)


ci, cd, cb = ctypes.c_int, ctypes.c_double, ctypes.c_ubyte
rng = np.random.default_rng(20261027)


def test_unary_and_binary_maps(g):
    v = rng.uniform(-1, 1, 12)
    a, b = v.copy(), v.copy()
    g.call(AK + "ruOp", ci(OP["RU_MUL"]), cd(2.5), J(a))
    ref.ruOp(OP["RU_MUL"], 2.5, b)
    assert np.array_equal(a, b) and not np.array_equal(a, v)
    a, b = v.copy(), v.copy()
    g.call(AK + "cuOp", ci(OP["CU_MUL"]), cd(0.5), cd(-1.5), J(a))  # aRe / aIm must not be swapped
    ref.cuOp(OP["CU_MUL"], 0.5, -1.5, b)
    assert np.array_equal(a, b)
    ia = rng.integers(-9, 9, 7).astype(np.int32)
    ib = ia.copy()
    g.call(AK + "iuOp", ci(OP["IU_ADD"]), ci(7), J(ia))
    ref.iuOp(OP["IU_ADD"], 7, ib)
    assert np.array_equal(ia, ib)
    x, y = rng.uniform(-1, 1, 8), rng.uniform(1, 2, 8)
    for op, cplx in ((OP["RE_SUB"], False), (OP["RE_DIV"], False), (OP["CE_DIV"], True), (OP["CE_MUL"], True)):
        d1, d2 = np.zeros(8), np.zeros(8)
        g.call(AK + "eOp", ci(op), J(x), J(y), J(d1), cb(cplx))  # lhs / rhs order matters for SUB and DIV
        ref.eOp(op, x, y, d2, cplx)
        assert np.array_equal(d1, d2) and np.any(d1 != 0)
    ix, iy = rng.integers(-9, 9, 6).astype(np.int32), rng.integers(-9, 9, 6).astype(np.int32)
    d1, d2 = np.zeros(6, dtype=np.int32), np.zeros(6, dtype=np.int32)
    g.call(AK + "eOp", ci(OP["IE_SUB"]), J(ix), J(iy), J(d1), cb(0))
    ref.eOp(OP["IE_SUB"], ix, iy, d2, False)
    assert np.array_equal(d1, d2)
    with pytest.raises(JavaException, match="Invalid array types"):
        g.call(AK + "eOp", ci(0), J(x), J(ix), J(d1), cb(0))
    with pytest.raises(JavaException, match="Invalid array lengths"):
        g.call(AK + "eOp", ci(0), J(x), J(y[:4].copy()), J(np.zeros(8)), cb(0))


def test_convert(g):
    r = rng.uniform(-1, 1, 5)
    c = rng.uniform(-1, 1, 10)
    i = rng.integers(-9, 9, 5).astype(np.int32)
    for op, src, sc, n_dst, dc in ((OP["R_TO_C_IM"], r, False, 10, True), (OP["R_TO_C_RE"], r, False, 10, True),
                                   (OP["C_TO_R_ABS"], c, True, 5, False), (OP["C_TO_R_IM"], c, True, 5, False),
                                   (OP["I_TO_R"], i, False, 5, False)):
        d1, d2 = np.full(n_dst, 9.0), np.full(n_dst, 9.0)
        g.call(AK + "convert", ci(op), J(src), cb(sc), J(d1), cb(dc))
        ref.convert(op, src, sc, d2, dc)
        assert np.array_equal(d1, d2), op
    with pytest.raises(JavaException, match="Invalid array types"):
        g.call(AK + "convert", ci(0), J(r), cb(1), J(np.zeros(5)), cb(1))


def test_reductions_scans_indices(g):
    v = rng.uniform(0.5, 1.5, 60)
    for op in ("RA_SUM", "RA_VAR", "RA_MAX"):
        assert g.call(AK + "raOp", ci(OP[op]), J(v), restype=ctypes.c_double) == ref.raOp(OP[op], v)
    got = g.array_result(g.call(AK + "caOp", ci(OP["CA_PROD"]), J(v), restype=ctypes.c_void_p), np.float64)
    assert np.array_equal(got, ref.caOp(OP["CA_PROD"], v))
    sD, sS = I([3, 4, 5]), I([20, 5, 1])
    dD, dS = I([1, 4, 1]), I([4, 1, 1])
    d1, d2 = np.zeros(4), np.zeros(4)
    od1, od2 = I([2, 0]), [2, 0]  # unsorted: the reference sorts opDims in place
    g.call(AK + "rrOp", ci(OP["RR_SUM"]), J(v), J(sD), J(sS), J(d1), J(dD), J(dS), J(od1))
    ref.rrOp(OP["RR_SUM"], v, sD, sS, d2, dD, dS, *od2)
    assert np.array_equal(d1, d2) and np.array_equal(od1, [0, 2])
    idx1, idx2 = np.zeros(60, dtype=np.int32), np.zeros(60, dtype=np.int32)
    w = np.round(v * 3)
    g.call(AK + "riOp", ci(OP["RI_MAX"]), J(w), J(sD), J(sS), J(idx1), ci(1))
    ref.riOp(OP["RI_MAX"], w.copy(), sD, sS, idx2, 1)
    assert np.array_equal(idx1, idx2)
    s1, s2 = v.copy(), v.copy()
    g.call(AK + "riOp", ci(OP["RI_SORT"]), J(s1), J(sD), J(sS), J(idx1), ci(2))  # sorts the SOURCE in place
    ref.riOp(OP["RI_SORT"], s2, sD, sS, idx2, 2)
    assert np.array_equal(s1, s2) and np.array_equal(idx1, idx2) and not np.array_equal(s1, v)
    d1, d2 = np.zeros(60), np.zeros(60)
    g.call(AK + "rdOp", ci(OP["RD_SUM"]), J(v), J(sD), J(sS), J(d1), J(I([0, 2])))
    ref.rdOp(OP["RD_SUM"], v, sD, sS, d2, 0, 2)
    assert np.array_equal(d1, d2)
    line = I([2, 0, -1, -1, 3, 1, -1, -1])
    got = g.array_result(g.call(AK + "find", J(line), J(I([2, 4])), J(I([4, 1])), J(I([1, -1])), restype=ctypes.c_void_p),
                         np.int32)
    assert np.array_equal(got, ref.find(line, [2, 4], [4, 1], [1, -1])) and got.size == 2


def test_map_and_slice_primitive_arrays(g):
    src = rng.uniform(-1, 1, 12)
    for dtype in (np.float64, np.int32):
        s = (src * 100).astype(dtype)
        d1, d2 = np.zeros(20, dtype=dtype), np.zeros(20, dtype=dtype)
        bounds = I([1, -1, 5, 2, 1, 3])  # wraps on both sides
        g.call(AK + "map", J(bounds), J(s), J(I([3, 4])), J(I([4, 1])), J(d1), J(I([4, 5])), J(I([5, 1])))
        ref.map(bounds, s, [3, 4], [4, 1], d2, [4, 5], [5, 1])
        assert np.array_equal(d1, d2) and np.any(d1 != 0)
        d1[:], d2[:] = 0, 0
        slices = I([0, 3, 0, 2, 1, 0, 1, 4, 1, 3, 0, 1])
        g.call(AK + "slice", J(slices), J(s), J(I([3, 4])), J(I([4, 1])), J(d1), J(I([4, 5])), J(I([5, 1])))
        ref.slice(slices, s, [3, 4], [4, 1], d2, [4, 5], [5, 1])
        assert np.array_equal(d1, d2) and np.any(d1 != 0)
    with pytest.raises(JavaException, match="Invalid array types"):
        g.call(AK + "map", J(bounds), J(src), J(I([3, 4])), J(I([4, 1])), J(np.zeros(20, dtype=np.int32)), J(I([4, 5])),
               J(I([5, 1])))


def test_matrices(g):
    a, b = rng.uniform(-1, 1, 12), rng.uniform(-1, 1, 20)
    d1, d2 = np.zeros(15), np.zeros(15)
    g.call(AK + "mul", J(a), J(b), ci(3), ci(5), J(d1), cb(0))  # lr = 3, rc = 5: a swap would be caught
    ref.mul(a, b, 3, 5, d2, False)
    assert np.array_equal(d1, d2)
    ca, cb_ = rng.uniform(-1, 1, 12), rng.uniform(-1, 1, 12)
    d1, d2 = np.zeros(8), np.zeros(8)
    g.call(AK + "mul", J(ca), J(cb_), ci(2), ci(2), J(d1), cb(1))
    ref.mul(ca, cb_, 2, 2, d2, True)
    assert np.array_equal(d1, d2)
    m = rng.uniform(-1, 1, 18)
    d1, d2 = np.zeros(6), np.zeros(6)
    g.call(AK + "diag", J(m), J(d1), ci(3), cb(1))
    ref.diag(m, d2, 3, True)
    assert np.array_equal(d1, d2)
    sq = rng.uniform(0, 1, 25) + 2 * np.eye(5).ravel()
    d1, d2 = np.zeros(25), np.zeros(25)
    g.call(AK + "invert", J(sq), J(d1), ci(5))
    ref.invert(sq, d2, 5)
    assert np.array_equal(d1, d2)
    tall = rng.uniform(0, 1, 24)  # 6 x 4 row-major, then the same buffer read as the transpose of a 4 x 6
    for rs, cs in ((4, 1), (1, 6)):
        u1, s1, v1, u2, s2, v2 = np.zeros(24), np.zeros(4), np.zeros(16), np.zeros(24), np.zeros(4), np.zeros(16)
        g.call(AK + "svd", J(tall), ci(rs), ci(cs), J(u1), J(s1), J(v1), ci(6), ci(4))
        ref.svd(tall, rs, cs, u2, s2, v2, 6, 4)
        assert np.array_equal(u1, u2) and np.array_equal(s1, s2) and np.array_equal(v1, v2) and s1[0] > 0
    e = rng.uniform(0, 1, 36)
    vec1, val1, vec2, val2 = np.zeros(36), np.zeros(12), np.zeros(36), np.zeros(12)
    g.call(AK + "eigs", J(e), J(vec1), J(val1), ci(6))
    ref.eigs(e, vec2, val2, 6)
    assert np.array_equal(vec1, vec2) and np.array_equal(val1, val2)
    with pytest.raises(JavaException, match="Invalid arguments"):
        g.call(AK + "eigs", J(e), J(vec1), J(np.zeros(11)), ci(6))


def test_fft_service_transform(g):
    name = "Java_org_sharedx_cuda_CudaFftService_transform"
    dims = [4, 6]
    n = 24
    x = rng.uniform(-0.5, 0.5, 2 * n)
    out = np.zeros(2 * n)
    g.call(name, ci(2), J(I(dims)), J(x), J(out))  # Plan.java:73-88: 2 = forward, 3 = backward, 0 = r2c, 1 = c2r
    assert np.array_equal(out, oracle.fft(dims, x))
    back = np.zeros(2 * n)
    g.call(name, ci(3), J(I(dims)), J(out), J(back))
    assert np.array_equal(back, oracle.ifft(dims, out)) and np.allclose(back, x)
    r = rng.uniform(-0.5, 0.5, n)
    half = np.zeros(2 * 4 * 4)
    g.call(name, ci(0), J(I(dims)), J(r), J(half))
    assert np.array_equal(half, oracle.rfft(dims, r))
    rr = np.zeros(n)
    g.call(name, ci(1), J(I(dims)), J(half), J(rr))
    assert np.array_equal(rr, oracle.rifft(dims, half)) and np.allclose(rr, r)
    with pytest.raises(JavaException, match="expected sizes"):
        g.call(name, ci(2), J(I(dims)), J(x), J(np.zeros(n)))


def test_image_kernel(g):
    src = rng.uniform(0, 1, 12)
    sD, sS, dD, dS = I([3, 4]), I([4, 1]), I([4, 5]), I([5, 1])
    d1, d2 = np.zeros(20), np.zeros(20)
    g.call("Java_org_sharedx_cuda_CudaImageKernel_createIntegralImage", J(src), J(sD), J(sS), J(d1), J(dD), J(dS))
    ref.createIntegralImage(src, sD, sS, d2, dD, dS)
    assert np.array_equal(d1, d2) and d1[-1] > 0
    mem = rng.integers(0, 3, 12).astype(np.int32)
    hD, hS = I([4, 5, 3]), I([15, 3, 1])
    d1, d2 = np.zeros(60), np.zeros(60)
    g.call("Java_org_sharedx_cuda_CudaImageKernel_createIntegralHistogram", J(src), J(sD), J(sS), J(mem), J(d1), J(hD),
           J(hS))
    ref.createIntegralHistogram(src, sD, sS, mem, d2, hD, hS)
    assert np.array_equal(d1, d2) and d1[-3:].sum() > 0
