"""Host-side mirror of the reference's ArrayKernel SPI for the CUDA provider.

`CudaArrayKernel` has the method names, argument order and error behaviour of
org.shared.array.kernel.ArrayKernel (src/org/shared/array/kernel/ArrayKernel.java:285-674) as implemented by
NativeArrayKernel (src/org/shared/array/jni/NativeArrayKernel.java:41-154): arrays are flat numpy float64 /
int32 arrays standing in for Java double[] / int[], mutated in place where the Java contract mutates them.
It is the Python twin of shared_b200/java/org/sharedx/cuda/CudaArrayKernel.java; both sit on the host tier of
include/sst_cuda.h. `DeviceArrayKernel` is the device tier (device pointers + stream) used by bench.py.
The members outside the north-star families (invert, svd, eigs, find, insertSparse, sliceSparse) are provided as well;
everything runs on the GPU -- there is no CPU fallback.
"""
import collections
import ctypes
import time

import numpy as np

from . import _lib

REAL, COMPLEX, INT = 0, 1, 2            # SSTC_ELEM_*
R_TO_C, C_TO_R, I_TO_R = 0, 1, 2        # SSTC_CONV_*


def _i32(values):
    return np.ascontiguousarray(values, dtype=np.int32)


def _pi(a):
    return a.ctypes.data_as(_lib.c_i32_p)


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


class SparseStateStruct(ctypes.Structure):
    """sstc_sparse_state (include/sst_cuda.h)."""
    _fields_ = [("count", ctypes.c_int32), ("n_offsets", ctypes.c_int32), ("n_indirections", ctypes.c_int32),
                ("elem_bytes", ctypes.c_int32), ("values", ctypes.c_void_p), ("indices", ctypes.c_void_p),
                ("indirection_offsets", ctypes.c_void_p), ("indirections", ctypes.c_void_p)]


# the four fields of org.shared.array.sparse.SparseArrayState (SparseArrayState.java:44-64)
SparseArrayState = collections.namedtuple("SparseArrayState", "values indices indirectionOffsets indirections")


def _check_array(a, name):
    if not isinstance(a, np.ndarray) or a.dtype not in (np.float64, np.int32) or not a.flags.c_contiguous:
        raise _lib.SstCudaError("Invalid array types")  # NativeArrayKernel.cpp:134 / ObjectArrayTest.java:301
    return a


class CudaArrayKernel:
    """Drop-in for JavaArrayKernel / NativeArrayKernel behind ModalArrayKernel (ModalArrayKernel.java:41-235)."""

    def __init__(self, device=0):
        _lib.init(device)
        self._l = _lib.lib()

    # -- randomness (Bindings.cpp:26-32) --------------------------------------------------------------------
    def randomize(self):
        _lib.check(self._l.sstc_srand(ctypes.c_uint64(time.time_ns() & 0xFFFFFFFFFFFFFFFF)))

    def derandomize(self):
        _lib.check(self._l.sstc_srand(ctypes.c_uint64(0xdeadbeefcacabead)))  # Arithmetic.java:43

    # -- movement ---------------------------------------------------------------------------------------------
    def _move(self, fn, spec, srcV, srcD, srcS, dstV, dstD, dstS):
        _check_array(srcV, "srcV")
        _check_array(dstV, "dstV")
        if srcV.dtype != dstV.dtype:
            raise _lib.SstCudaError("Invalid array types")
        sD, sS, dD, dS, sp = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS), _i32(spec)
        if not (sD.size == sS.size == dD.size == dS.size):
            raise _lib.SstCudaError("Invalid arguments")
        _lib.check(fn(srcV.dtype.itemsize, _pi(sp), sp.size, _vp(srcV), srcV.size, _pi(sD), _pi(sS), _vp(dstV),
                      dstV.size, _pi(dD), _pi(dS), sD.size))

    def map(self, bounds, srcV, srcD, srcS, dstV, dstD, dstS):
        self._move(self._l.sstc_map, bounds, srcV, srcD, srcS, dstV, dstD, dstS)

    def slice(self, slices, srcV, srcD, srcS, dstV, dstD, dstS):
        self._move(self._l.sstc_slice, slices, srcV, srcD, srcS, dstV, dstD, dstS)

    # -- dimension operations -----------------------------------------------------------------------------------
    def rrOp(self, type_, srcV, srcD, srcS, dstV, dstD, dstS, *opDims):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        od = _i32(list(opDims))
        if not (sD.size == sS.size == dD.size == dS.size):
            raise _lib.SstCudaError("Invalid arguments")
        _lib.check(self._l.sstc_rrop(int(type_), _vp(srcV), srcV.size, _pi(sD), _pi(sS), _vp(dstV), dstV.size, _pi(dD),
                                     _pi(dS), sD.size, _pi(od), od.size))

    def riOp(self, type_, srcV, srcD, srcS, dstV, dim):
        sD, sS = _i32(srcD), _i32(srcS)
        if sD.size != sS.size or dstV.dtype != np.int32:
            raise _lib.SstCudaError("Invalid arguments")
        _lib.check(self._l.sstc_riop(int(type_), _vp(srcV), srcV.size, _pi(sD), _pi(sS), sD.size, _vp(dstV), dstV.size,
                                     int(dim)))

    def rdOp(self, type_, srcV, srcD, srcS, dstV, *opDims):
        sD, sS, od = _i32(srcD), _i32(srcS), _i32(list(opDims))
        if sD.size != sS.size:
            raise _lib.SstCudaError("Invalid arguments")
        _lib.check(self._l.sstc_rdop(int(type_), _vp(srcV), srcV.size, _pi(sD), _pi(sS), sD.size, _vp(dstV), dstV.size,
                                     _pi(od), od.size))

    # -- accumulators ---------------------------------------------------------------------------------------------
    def raOp(self, type_, srcV):
        out = ctypes.c_double()
        _lib.check(self._l.sstc_raop(int(type_), _vp(srcV), srcV.size, ctypes.byref(out)))
        return out.value

    def caOp(self, type_, srcV):
        out = np.zeros(2)
        _lib.check(self._l.sstc_caop(int(type_), _vp(srcV), srcV.size, _vp(out)))
        return out

    # -- element-wise -----------------------------------------------------------------------------------------------
    def ruOp(self, type_, a, srcV):
        _lib.check(self._l.sstc_ruop(int(type_), float(a), _vp(srcV), srcV.size))

    def cuOp(self, type_, aRe, aIm, srcV):
        _lib.check(self._l.sstc_cuop(int(type_), float(aRe), float(aIm), _vp(srcV), srcV.size))

    def iuOp(self, type_, a, srcV):
        if srcV.dtype != np.int32:
            raise _lib.SstCudaError("Invalid array types")
        _lib.check(self._l.sstc_iuop(int(type_), int(a), _vp(srcV), srcV.size))

    def eOp(self, type_, lhsV, rhsV, dstV, complex_):
        for a in (lhsV, rhsV, dstV):
            _check_array(a, "eOp operand")
        # ElementOps.cpp:285-321: all three double[] (complex or real by the flag), or all three int[]
        if lhsV.dtype == rhsV.dtype == dstV.dtype == np.float64:
            elem = COMPLEX if complex_ else REAL
        elif lhsV.dtype == rhsV.dtype == dstV.dtype == np.int32:
            elem = INT
        else:
            raise _lib.SstCudaError("Invalid array types")
        if not (lhsV.size == rhsV.size == dstV.size):
            raise _lib.SstCudaError("Invalid array lengths")  # ElementOps.cpp:570-572
        _lib.check(self._l.sstc_eop(int(type_), elem, _vp(lhsV), _vp(rhsV), _vp(dstV), dstV.size))

    def convert(self, type_, srcV, isSrcComplex, dstV, isDstComplex):
        _check_array(srcV, "srcV")
        _check_array(dstV, "dstV")
        # ElementOps.cpp:329-401
        if srcV.dtype == np.float64 and not isSrcComplex and dstV.dtype == np.float64 and isDstComplex:
            conv = R_TO_C
        elif srcV.dtype == np.float64 and isSrcComplex and dstV.dtype == np.float64 and not isDstComplex:
            conv = C_TO_R
        elif srcV.dtype == np.int32 and not isSrcComplex and dstV.dtype == np.float64 and not isDstComplex:
            conv = I_TO_R
        else:
            raise _lib.SstCudaError("Invalid array types")
        _lib.check(self._l.sstc_convert(int(type_), conv, _vp(srcV), srcV.size, _vp(dstV), dstV.size))

    # -- matrices ---------------------------------------------------------------------------------------------------
    def mul(self, lhsV, rhsV, lr, rc, dstV, complex_):
        _lib.check(self._l.sstc_mul(_vp(lhsV), lhsV.size, _vp(rhsV), rhsV.size, int(lr), int(rc), _vp(dstV), dstV.size,
                                    int(bool(complex_))))

    def diag(self, srcV, dstV, size, complex_):
        _lib.check(self._l.sstc_diag(_vp(srcV), srcV.size, _vp(dstV), dstV.size, int(size), int(bool(complex_))))

    def invert(self, srcV, dstV, size):
        """ArrayKernel.invert (ArrayKernel.java:576-586; LinearAlgebraOpsLu.cpp:26-76)."""
        _lib.check(self._l.sstc_invert(_vp(srcV), srcV.size, _vp(dstV), dstV.size, int(size)))

    def svd(self, srcV, srcStrideRow, srcStrideCol, uV, sV, vV, nRows, nCols):
        """ArrayKernel.svd (ArrayKernel.java:558-560; LinearAlgebraOpsSvd.cpp:26-72)."""
        _lib.check(self._l.sstc_svd(_vp(srcV), srcV.size, int(srcStrideRow), int(srcStrideCol), _vp(uV), uV.size, _vp(sV),
                                    sV.size, _vp(vV), vV.size, int(nRows), int(nCols)))

    def eigs(self, srcV, vecV, valV, size):
        """ArrayKernel.eigs (ArrayKernel.java:574; LinearAlgebraOpsEigs.cpp:26-615): bit-identical to the reference."""
        _lib.check(self._l.sstc_eigs(_vp(srcV), srcV.size, _vp(vecV), vecV.size, _vp(valV), valV.size, int(size)))

    def find(self, srcV, srcD, srcS, logical):
        """ArrayKernel.find (ArrayKernel.java:603; IndexOps.java:44-91): returns a new int32 array."""
        if srcV.dtype != np.int32:
            raise _lib.SstCudaError("Invalid array types")
        sD, sS, lg = _i32(srcD), _i32(srcS), _i32(logical)
        if not (sD.size == sS.size == lg.size):
            raise _lib.SstCudaError("Invalid arguments")
        active = [d for d in range(lg.size) if lg[d] == -1]
        line = np.empty(int(sD[active[0]]) if len(active) == 1 else 0, dtype=np.int32)
        count = ctypes.c_int32()
        _lib.check(self._l.sstc_find(_vp(srcV), srcV.size, _pi(sD), _pi(sS), sD.size, _pi(lg), _vp(line), line.size,
                                     ctypes.byref(count)))
        return line[:count.value].copy()

    # -- sparse arrays (ArrayKernel.java:628-674) ------------------------------------------------------------------------
    def _sparse_state(self, st, dtype):
        def take(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            raw = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(raw, dtype=dt).copy()  # the library reuses its buffers on this thread's next call
        return SparseArrayState(take(st.values, st.count, dtype), take(st.indices, st.count, np.int32),
                                take(st.indirection_offsets, st.n_offsets, np.int32),
                                take(st.indirections, st.n_indirections, np.int32))

    def insertSparse(self, oldV, oldD, oldS, oldDo, oldI, newV, newI):
        """ArrayKernel.insertSparse (ArrayKernel.java:628-650; SparseOpsInsert.cpp:26-168): returns the new
        SparseArrayState; newI (an int32 array) comes back sorted, as with both reference providers."""
        _check_array(oldV, "oldV")
        _check_array(newV, "newV")
        if oldV.dtype != newV.dtype:
            raise _lib.SstCudaError("Invalid array types")
        oD, oS, oDo, oI = _i32(oldD), _i32(oldS), _i32(oldDo), _i32(oldI)
        if not (isinstance(newI, np.ndarray) and newI.dtype == np.int32 and newI.flags.c_contiguous):
            raise _lib.SstCudaError("Invalid array types")
        if not (oD.size == oS.size and oDo.size == oD.size + 1 and oI.size == oldV.size and newI.size == newV.size):
            raise _lib.SstCudaError("Invalid arguments")  # SparseOpsInsert.cpp:77-83
        st = SparseStateStruct()
        _lib.check(self._l.sstc_insert_sparse(oldV.dtype.itemsize, _vp(oldV), oldV.size, _pi(oD), _pi(oS), _pi(oDo), oD.size,
                                              _vp(oI), _vp(newV), newV.size, _vp(newI), ctypes.byref(st)))
        return self._sparse_state(st, oldV.dtype)

    def sliceSparse(self, slices, srcV, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, dstD, dstS, dstDo, dstI, dstIo,
                    dstIi):
        """ArrayKernel.sliceSparse (ArrayKernel.java:652-674; SparseOpsSlice.cpp:26-410)."""
        _check_array(srcV, "srcV")
        _check_array(dstV, "dstV")
        if srcV.dtype != dstV.dtype:
            raise _lib.SstCudaError("Invalid array types")
        sl = _i32(slices)
        sD, sS, sDo, sI, sIo, sIi = (_i32(a) for a in (srcD, srcS, srcDo, srcI, srcIo, srcIi))
        dD, dS, dDo, dI, dIo, dIi = (_i32(a) for a in (dstD, dstS, dstDo, dstI, dstIo, dstIi))
        nd = sD.size
        if not (sS.size == nd and dD.size == nd and dS.size == nd and sl.size % 3 == 0 and sI.size == srcV.size and
                dI.size == dstV.size and sDo.size == nd + 1 and dDo.size == nd + 1 and sIi.size == nd * srcV.size and
                dIi.size == nd * dstV.size and sIo.size == int(sD.sum()) + nd and dIo.size == int(dD.sum()) + nd):
            raise _lib.SstCudaError("Invalid arguments")  # SparseOpsSlice.cpp:87-98, 128-131
        st = SparseStateStruct()
        _lib.check(self._l.sstc_slice_sparse(srcV.dtype.itemsize, _pi(sl), sl.size, _vp(srcV), srcV.size, _pi(sD), _pi(sS),
                                             _pi(sDo), _vp(sI), _pi(sIo), _vp(sIi), _vp(dstV), dstV.size, _pi(dD), _pi(dS),
                                             _pi(dDo), _vp(dI), _pi(dIo), _vp(dIi), nd, ctypes.byref(st)))
        return self._sparse_state(st, srcV.dtype)


class DeviceArrayKernel:
    """Device tier: value arrays are device pointers (ints), metadata stays on the host; asynchronous on `stream`."""

    def __init__(self, device=0, stream=0):
        _lib.init(device)
        self._l = _lib.lib()
        self.stream = ctypes.c_void_p(int(stream))

    def ruOp(self, type_, a, ptr, n):
        _lib.check(self._l.sstc_ruop_dev(int(type_), float(a), ctypes.c_void_p(ptr), int(n), self.stream))

    def eOp(self, type_, elem, lhs, rhs, dst, n):
        _lib.check(self._l.sstc_eop_dev(int(type_), int(elem), ctypes.c_void_p(lhs), ctypes.c_void_p(rhs),
                                        ctypes.c_void_p(dst), int(n), self.stream))

    def raOp(self, type_, ptr, n, out_ptr):
        _lib.check(self._l.sstc_raop_dev(int(type_), ctypes.c_void_p(ptr), int(n), ctypes.c_void_p(out_ptr), self.stream))

    def rrOp(self, type_, src, nsrc, srcD, srcS, dst, ndst, dstD, dstS, *opDims):
        sD, sS, dD, dS, od = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS), _i32(list(opDims))
        _lib.check(self._l.sstc_rrop_dev(int(type_), ctypes.c_void_p(src), int(nsrc), _pi(sD), _pi(sS),
                                         ctypes.c_void_p(dst), int(ndst), _pi(dD), _pi(dS), sD.size, _pi(od), od.size,
                                         self.stream))

    def riOp(self, type_, src, nsrc, srcD, srcS, dst, ndst, dim):
        sD, sS = _i32(srcD), _i32(srcS)
        _lib.check(self._l.sstc_riop_dev(int(type_), ctypes.c_void_p(src), int(nsrc), _pi(sD), _pi(sS), sD.size,
                                         ctypes.c_void_p(dst), int(ndst), int(dim), self.stream))

    def rdOp(self, type_, src, nsrc, srcD, srcS, dst, ndst, *opDims):
        sD, sS, od = _i32(srcD), _i32(srcS), _i32(list(opDims))
        _lib.check(self._l.sstc_rdop_dev(int(type_), ctypes.c_void_p(src), int(nsrc), _pi(sD), _pi(sS), sD.size,
                                         ctypes.c_void_p(dst), int(ndst), _pi(od), od.size, self.stream))

    def map(self, elem_bytes, bounds, src, nsrc, srcD, srcS, dst, ndst, dstD, dstS):
        b, sD, sS, dD, dS = _i32(bounds), _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        _lib.check(self._l.sstc_map_dev(int(elem_bytes), _pi(b), b.size, ctypes.c_void_p(src), int(nsrc), _pi(sD),
                                        _pi(sS), ctypes.c_void_p(dst), int(ndst), _pi(dD), _pi(dS), sD.size, self.stream))

    def mul(self, lhs, nl, rhs, nr, lr, rc, dst, ndst, complex_):
        _lib.check(self._l.sstc_mul_dev(ctypes.c_void_p(lhs), int(nl), ctypes.c_void_p(rhs), int(nr), int(lr), int(rc),
                                        ctypes.c_void_p(dst), int(ndst), int(bool(complex_)), self.stream))


def smoke_check():
    """A few host-tier calls on cuda:0 (used by __graft_entry__.smoke): KAT values from ArrayKernelTest.java."""
    k = CudaArrayKernel(0)
    assert abs(k.raOp(0, np.array([1.0, -0.5, 0.25, -0.125])) - 0.625) < 1e-12  # ArrayKernelTest.java:69
    v = np.array([0.0, -1, 1, 0, -1, 1, 0, -1, 1])
    k.ruOp(1, 2.0, v)  # RU_MUL, ArrayKernelTest.java:102
    assert np.array_equal(v, np.array([0.0, -2, 2, 0, -2, 2, 0, -2, 2]))
    a = np.arange(60, dtype=np.float64)
    out = np.zeros(15)
    k.rrOp(0, a, [3, 4, 5], [20, 5, 1], out, [3, 1, 5], [5, 5, 1], 1)  # RealArrayTest.java:726-758
    assert np.array_equal(out, a.reshape(3, 4, 5).sum(1).ravel())
    m = np.array([[2.0, 0, 0], [0, 0, -1], [0, 1, 0]])  # eigenvalues 2, +-i
    vec, val = np.zeros(9), np.zeros(6)
    k.eigs(m.ravel().copy(), vec, val, 3)
    assert sorted(val[0::2].tolist()) == [0.0, 0.0, 2.0] and sorted(val[1::2].tolist()) == [-1.0, 0.0, 1.0]
    st = k.insertSparse(np.zeros(0), [2, 3], [3, 1], [0, 3, 7], np.zeros(0, dtype=np.int32), np.array([5.0, 4.0]),
                        np.array([4, 1], dtype=np.int32))
    assert np.array_equal(st.indices, [1, 4]) and np.array_equal(st.values, [4.0, 5.0])
    assert np.array_equal(st.indirectionOffsets, [0, 1, 2, 0, 0, 2, 2]) and np.array_equal(st.indirections, [0, 1, 0, 1])
