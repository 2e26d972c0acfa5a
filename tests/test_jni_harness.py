"""The JNI glue of libsst_cuda.so (shared_b200/csrc/jni/jni_glue.cpp) EXECUTED on the CPU against the functional JNI model
of oracle/ref/jni.h (tests/jni_harness): provider registration in JNI_OnLoad, SparseArrayState construction, and the
Object[] paths (map / slice moved on the JVM side; insertSparse / sliceSparse through positions). No JDK and no GPU are
involved: the two sparse entry points of the C ABI are served by the host executor of tests/emu, every other one by a
failing stub -- what is under test is the glue's own logic."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from sparse_cases import dim_offsets, empty_state, insert_cases, row_major, slice_cases

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def jh():
    if shutil.which("g++") is None or shutil.which("make") is None:
        pytest.skip("no g++ / make")
    res = subprocess.run(["make", "-s", "-k", "-C", os.path.join(HERE, "jni_harness"), "all"], capture_output=True, text=True)
    path = os.path.join(HERE, "jni_harness", "_build", "libjni_harness.so")
    if not os.path.exists(path):
        pytest.fail("harness build failed:\n" + res.stdout[-3000:] + res.stderr[-3000:])
    return ctypes.CDLL(path)


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class GlueError(RuntimeError):
    pass


class GlueKernel:
    """insertSparse / sliceSparse as the Java side would see them through the glue. as_objects=True hands the values over
    as Object[] whose elements carry the (int32) values as identities."""

    def __init__(self, lib, as_objects=False):
        self.lib, self.as_objects = lib, as_objects

    def _kind(self, v):
        return 3 if self.as_objects else (1 if v.dtype == np.float64 else 2)

    def _finish(self, rc, out, err, dtype):
        if rc:
            raise GlueError(err.value.decode())
        v, i, io, ii, counts = out
        return v[:counts[0]].astype(dtype), i[:counts[0]].copy(), io[:counts[1]].copy(), ii[:counts[2]].copy()

    def _out(self, dtype, cap, nd, dims):
        cap = max(1, cap)
        return (np.zeros(cap, dtype=np.int32 if self.as_objects else dtype), np.zeros(cap, dtype=np.int32),
                np.zeros(max(1, int(np.sum(dims)) + nd), dtype=np.int32), np.zeros(max(1, nd * cap), dtype=np.int32),
                np.zeros(3, dtype=np.int32))

    def insertSparse(self, oldV, oldD, oldS, oldDo, oldI, newV, newI):
        oD, oS, oDo, oI = i32(oldD), i32(oldS), i32(oldDo), i32(oldI)
        out = self._out(oldV.dtype, oldV.size + newV.size, oD.size, oD)
        err = ctypes.create_string_buffer(256)
        rc = self.lib.jh_insert_sparse(self._kind(oldV), p(oldV), oldV.size, p(oD), p(oS), p(oDo), oD.size, p(oI), p(newV),
                                       newV.size, p(newI), p(out[0]), p(out[1]), p(out[2]), p(out[3]), p(out[4]), err, 256)
        return self._finish(rc, out, err, oldV.dtype)

    def sliceSparse(self, slices, srcV, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, dstD, dstS, dstDo, dstI, dstIo, dstIi):
        a = [i32(x) for x in (slices, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstD, dstS, dstDo, dstI, dstIo, dstIi)]
        nd = a[1].size
        cap = dstV.size + int(np.prod([max(1, int(np.sum(a[0][2::3] == d))) for d in range(nd)]))
        out = self._out(srcV.dtype, cap, nd, a[7])
        err = ctypes.create_string_buffer(256)
        rc = self.lib.jh_slice_sparse(self._kind(srcV), p(a[0]), a[0].size, p(srcV), srcV.size, p(a[1]), p(a[2]), p(a[3]),
                                      p(a[4]), p(a[5]), a[5].size, p(a[6]), p(dstV), dstV.size, p(a[7]), p(a[8]), p(a[9]),
                                      p(a[10]), p(a[11]), a[11].size, p(a[12]), nd, p(out[0]), p(out[1]), p(out[2]),
                                      p(out[3]), p(out[4]), err, 256)
        return self._finish(rc, out, err, srcV.dtype)


def test_onload_registers_the_three_providers(jh):
    """Library.cpp:55-77: registerService(spec, impl) for every provider, then Library.initialized = true."""
    buf = ctypes.create_string_buffer(1024)
    init = ctypes.c_int(-1)
    n = jh.jh_onload(buf, 1024, ctypes.byref(init))
    assert n == 3 and init.value == 1
    assert buf.value.decode().rstrip(";").split(";") == [
        "org/shared/array/kernel/ArrayKernel->org/sharedx/cuda/CudaArrayKernel",
        "org/shared/fft/FftService->org/sharedx/cuda/CudaFftService",
        "org/shared/image/kernel/ImageKernel->org/sharedx/cuda/CudaImageKernel"]


def test_sparse_through_the_glue_matches_the_fixture(jh):
    """double[] and int[] values: the SparseArrayState the glue builds equals the reference's (golden fixture)."""
    import sparse_checks as sc
    jh.jh_onload(ctypes.create_string_buffer(1024), 1024, ctypes.byref(ctypes.c_int()))
    k = GlueKernel(jh)
    assert sc.check_insert_cases(k, oracle.ref()) >= 20
    assert sc.check_slice_cases(k, oracle.ref()) >= 20
    sc.check_errors(k, GlueError)


def test_object_values_follow_the_int_path(jh):
    """Object[] values: the glue runs the same call on positions and gathers the references; the objects that come back
    must be the ones the int[] path would have moved (identities = the int values)."""
    from sparse_cases import run_insert, run_slice
    jh.jh_onload(ctypes.create_string_buffer(1024), 1024, ctypes.byref(ctypes.c_int()))
    ints, objs = GlueKernel(jh), GlueKernel(jh, as_objects=True)
    checked = 0
    for c in insert_cases():
        if c["old_vals"].dtype != np.int32:
            continue
        for a, b in zip(run_insert(ints, c), run_insert(objs, c)):
            assert np.array_equal(a, b), c["dims"]
        checked += 1
    for c in slice_cases():
        if c["svals"].dtype != np.int32:
            continue
        for a, b in zip(run_slice(ints, c), run_slice(objs, c)):
            assert np.array_equal(a, b), (c["sdims"], c["ddims"])
        checked += 1
    assert checked >= 20


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libsst_ref.so missing")
def test_object_map_and_slice_match_the_reference(jh):
    """Object[] map / slice are moved by the glue itself (no array contents in HBM). The same bounds / slices on an int[]
    through the reference's native MappingOps must move the same elements (ObjectArrayTest exercises this path)."""
    jh.jh_onload(ctypes.create_string_buffer(1024), 1024, ctypes.byref(ctypes.c_int()))
    ref = oracle.ref()
    rng = np.random.default_rng(20261026)
    err = ctypes.create_string_buffer(256)
    for trial in range(300):
        nd = int(rng.integers(1, 4))
        sD = [int(rng.integers(1, 6)) for _ in range(nd)]
        dD = [int(rng.integers(1, 6)) for _ in range(nd)]
        sS, dS = row_major(sD), row_major(dD)
        if rng.random() < 0.3:  # column-major destination: a transpose through map
            dS = [int(np.prod(dD[:i])) for i in range(nd)]
        src = rng.integers(0, 1000, int(np.prod(sD))).astype(np.int32)
        dst0 = rng.integers(1000, 2000, int(np.prod(dD))).astype(np.int32)
        if trial % 2 == 0:  # map: (srcStart, dstStart, size) per dimension, wrapping, negative starts
            spec = np.array([v for d in range(nd) for v in (int(rng.integers(-3, 6)), int(rng.integers(-3, 6)),
                                                            int(rng.integers(0, 2 * max(sD[d], dD[d]) + 1)))], dtype=np.int32)
            want = dst0.copy()
            ref.map(spec, src, sD, sS, want, dD, dS)
            is_map = 1
        else:  # slice: (src index, dst index, dim) triples, repeated destination indices included
            tr = [[int(rng.integers(0, sD[d])), int(rng.integers(0, dD[d])), d] for d in range(nd)
                  for _ in range(int(rng.integers(0, 2 * dD[d] + 1)))]
            spec = np.array(tr, dtype=np.int32).reshape(-1, 3).ravel()
            want = dst0.copy()
            ref.slice(spec, src, sD, sS, want, dD, dS)
            is_map = 0
        got = dst0.copy()
        rc = jh.jh_move_objects(is_map, p(spec), spec.size, p(src), src.size, p(i32(sD)), p(i32(sS)), p(got), got.size,
                                p(i32(dD)), p(i32(dS)), nd, err, 256)
        assert rc == 0, err.value
        assert np.array_equal(got, want), (is_map, sD, dD, spec)
    # the reference's error texts on the Object[] path
    src, dst = np.arange(4, dtype=np.int32), np.arange(4, dtype=np.int32)
    one = i32([4]), i32([1])
    rc = jh.jh_move_objects(0, p(i32([0, 5, 0])), 3, p(src), 4, p(one[0]), p(one[1]), p(dst), 4, p(one[0]), p(one[1]), 1, err, 256)
    assert rc == 1 and err.value == b"Invalid index"
    rc = jh.jh_move_objects(1, p(i32([0, 0, -1])), 3, p(src), 4, p(one[0]), p(one[1]), p(dst), 4, p(one[0]), p(one[1]), 1, err, 256)
    assert rc == 1 and err.value == b"Invalid mapping parameters"
    rc = jh.jh_move_objects(0, p(i32([0, 0, 0])), 3, p(src), 4, p(i32([5])), p(one[1]), p(dst), 4, p(one[0]), p(one[1]), 1, err, 256)
    assert rc == 1 and err.value == b"Invalid dimensions and/or strides"
