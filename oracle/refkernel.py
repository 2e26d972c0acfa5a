"""TEST INFRASTRUCTURE ONLY -- the reference's own native ArrayKernel as a Python object.

`RefKernel` exposes the 16 array-kernel methods of org.shared.array.kernel.ArrayKernel
(src/org/shared/array/kernel/ArrayKernel.java:285-534) with the Java argument order, backed by
oracle/_ref/libsst_ref.so = /root/reference/native/src/shared/*.cpp compiled UNMODIFIED against the shim
oracle/ref/jni.h (recipe: oracle/Makefile, wrappers: oracle/ref/ref_driver.cpp). Arrays are numpy float64 /
int32 (Java double[] / int[]), mutated in place exactly as the JNI code mutates pinned arrays. A Java
RuntimeException surfaces as RefKernelError(message). This is `cpu_baseline.kind == "reference"` material:
the reference's code, not a port.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libsst_ref.so")
_KERNEL = None

DOUBLE, INT = 1, 2  # shim_kind in oracle/ref/jni.h


class RefKernelError(RuntimeError):
    pass


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f64(a, writable=False):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous, "need a float64 array"
    return a


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def _kind(a):
    if a.dtype == np.float64:
        return DOUBLE
    if a.dtype == np.int32:
        return INT
    raise RefKernelError("Invalid array types")


class RefKernel:
    def __init__(self, lib):
        self._l = lib
        lib.sstref_last_error.restype = ctypes.c_char_p

    def _chk(self, rc):
        if rc != 0:
            raise RefKernelError(self._l.sstref_last_error().decode())

    def derandomize(self):
        self._l.sstref_srand(0)

    def ruOp(self, type_, a, srcV):
        self._chk(self._l.sstref_ruop(int(type_), ctypes.c_double(a), _p(_f64(srcV)), srcV.size))

    def cuOp(self, type_, aRe, aIm, srcV):
        self._chk(self._l.sstref_cuop(int(type_), ctypes.c_double(aRe), ctypes.c_double(aIm), _p(_f64(srcV)), srcV.size))

    def iuOp(self, type_, a, srcV):
        assert srcV.dtype == np.int32
        self._chk(self._l.sstref_iuop(int(type_), int(a), _p(srcV), srcV.size))

    def eOp(self, type_, lhsV, rhsV, dstV, complex_):
        k = _kind(lhsV)
        if _kind(rhsV) != k or _kind(dstV) != k:
            raise RefKernelError("Invalid array types")
        if not (lhsV.size == rhsV.size == dstV.size):
            raise RefKernelError("Invalid array lengths")
        self._chk(self._l.sstref_eop(int(type_), k, _p(lhsV), _p(rhsV), _p(dstV), dstV.size, int(bool(complex_))))

    def convert(self, type_, srcV, isSrcComplex, dstV, isDstComplex):
        self._chk(self._l.sstref_convert(int(type_), _p(srcV), _kind(srcV), srcV.size, int(bool(isSrcComplex)),
                                         _p(_f64(dstV)), dstV.size, int(bool(isDstComplex))))

    def raOp(self, type_, srcV):
        out = ctypes.c_double()
        self._chk(self._l.sstref_raop(int(type_), _p(_f64(srcV)), srcV.size, ctypes.byref(out)))
        return out.value

    def caOp(self, type_, srcV):
        out = (ctypes.c_double * 2)()
        self._chk(self._l.sstref_caop(int(type_), _p(_f64(srcV)), srcV.size, out))
        return np.array([out[0], out[1]])

    def rrOp(self, type_, srcV, srcD, srcS, dstV, dstD, dstS, *opDims):
        sD, sS, dD, dS, od = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS), _i32(list(opDims))
        self._chk(self._l.sstref_rrop(int(type_), _p(_f64(srcV)), srcV.size, _p(sD), _p(sS), sD.size, _p(_f64(dstV)),
                                      dstV.size, _p(dD), _p(dS), _p(od), od.size))

    def riOp(self, type_, srcV, srcD, srcS, dstV, dim):
        sD, sS = _i32(srcD), _i32(srcS)
        assert dstV.dtype == np.int32
        self._chk(self._l.sstref_riop(int(type_), _p(_f64(srcV)), srcV.size, _p(sD), _p(sS), sD.size, _p(dstV),
                                      dstV.size, int(dim)))

    def rdOp(self, type_, srcV, srcD, srcS, dstV, *opDims):
        sD, sS, od = _i32(srcD), _i32(srcS), _i32(list(opDims))
        self._chk(self._l.sstref_rdop(int(type_), _p(_f64(srcV)), srcV.size, _p(sD), _p(sS), sD.size, _p(_f64(dstV)),
                                      dstV.size, _p(od), od.size))

    def map(self, bounds, srcV, srcD, srcS, dstV, dstD, dstS):
        b, sD, sS, dD, dS = _i32(bounds), _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        if _kind(srcV) != _kind(dstV):
            raise RefKernelError("Invalid array types")
        self._chk(self._l.sstref_map(_p(b), b.size, _p(srcV), _kind(srcV), srcV.size, _p(sD), _p(sS), sD.size,
                                     _p(dstV), dstV.size, _p(dD), _p(dS), dD.size))

    def slice(self, slices, srcV, srcD, srcS, dstV, dstD, dstS):
        b, sD, sS, dD, dS = _i32(slices), _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        if _kind(srcV) != _kind(dstV):
            raise RefKernelError("Invalid array types")
        self._chk(self._l.sstref_slice(_p(b), b.size, _p(srcV), _kind(srcV), srcV.size, _p(sD), _p(sS), sD.size,
                                       _p(dstV), dstV.size, _p(dD), _p(dS), dD.size))

    def mul(self, lhsV, rhsV, lr, rc, dstV, complex_):
        self._chk(self._l.sstref_mul(_p(_f64(lhsV)), lhsV.size, _p(_f64(rhsV)), rhsV.size, int(lr), int(rc),
                                     _p(_f64(dstV)), dstV.size, int(bool(complex_))))

    def diag(self, srcV, dstV, size, complex_):
        self._chk(self._l.sstref_diag(_p(_f64(srcV)), srcV.size, _p(_f64(dstV)), dstV.size, int(size),
                                      int(bool(complex_))))

    def invert(self, srcV, dstV, size):
        self._chk(self._l.sstref_invert(_p(_f64(srcV)), srcV.size, _p(_f64(dstV)), dstV.size, int(size)))

    def svd(self, srcV, srcStrideRow, srcStrideCol, uV, sV, vV, nRows, nCols):
        self._chk(self._l.sstref_svd(_p(_f64(srcV)), srcV.size, int(srcStrideRow), int(srcStrideCol), _p(_f64(uV)), uV.size,
                                     _p(_f64(sV)), sV.size, _p(_f64(vV)), vV.size, int(nRows), int(nCols)))

    def eigs(self, srcV, vecV, valV, size):
        self._chk(self._l.sstref_eigs(_p(_f64(srcV)), srcV.size, _p(_f64(vecV)), vecV.size, _p(_f64(valV)), valV.size,
                                      int(size)))

    def find(self, srcV, srcD, srcS, logical):
        sD, sS, lg = _i32(srcD), _i32(srcS), _i32(logical)
        assert srcV.dtype == np.int32
        if not (sD.size == sS.size == lg.size):
            raise RefKernelError("Invalid arguments")
        out = np.empty(max(1, srcV.size), dtype=np.int32)
        n = self._l.sstref_find(_p(srcV), srcV.size, _p(sD), _p(sS), sD.size, _p(lg), _p(out), out.size)
        if n < 0:
            raise RefKernelError(self._l.sstref_last_error().decode())
        return out[:n].copy()

    # -- sparse arrays (ArrayKernel.java:628-674; SparseOpsInsert.cpp, SparseOpsSlice.cpp, SparseOps.cpp) ------
    # Both return (values, indices, indirectionOffsets, indirections) -- the fields of SparseArrayState.
    def _sparse_out(self, dtype, cap, nd, dims):
        cap = max(1, cap)
        return (np.zeros(cap, dtype=dtype), np.zeros(cap, dtype=np.int32),
                np.zeros(max(1, int(np.sum(dims)) + nd), dtype=np.int32), np.zeros(max(1, nd * cap), dtype=np.int32),
                np.zeros(3, dtype=np.int32))

    def _sparse_done(self, rc, out):
        if rc == 1:
            raise RefKernelError(self._l.sstref_last_error().decode())
        assert rc == 0, "oracle buffers too small"
        v, i, io, ii, counts = out
        return v[:counts[0]].copy(), i[:counts[0]].copy(), io[:counts[1]].copy(), ii[:counts[2]].copy()

    def insertSparse(self, oldV, oldD, oldS, oldDo, oldI, newV, newI):
        """newI (an int32 array) is sorted in place, as the reference does (SparseOpsInsert.cpp:118-126)."""
        oD, oS, oDo, oI = _i32(oldD), _i32(oldS), _i32(oldDo), _i32(oldI)
        assert newI.dtype == np.int32 and newI.flags.c_contiguous
        if _kind(oldV) != _kind(newV):
            raise RefKernelError("Invalid array types")
        if not (oD.size == oS.size and oDo.size == oD.size + 1 and oI.size == oldV.size and newI.size == newV.size):
            raise RefKernelError("Invalid arguments")
        out = self._sparse_out(oldV.dtype, oldV.size + newV.size, oD.size, oD)
        rc = self._l.sstref_insert_sparse(_kind(oldV), _p(oldV), oldV.size, _p(oD), _p(oS), _p(oDo), oD.size, _p(oI),
                                          _p(newV), newV.size, _p(newI), _p(out[0]), out[0].size, _p(out[1]), _p(out[2]),
                                          out[2].size, _p(out[3]), out[3].size, _p(out[4]))
        return self._sparse_done(rc, out)

    def sliceSparse(self, slices, srcV, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, dstD, dstS, dstDo, dstI, dstIo,
                    dstIi):
        sl = _i32(slices)
        sD, sS, sDo, sI, sIo, sIi = (_i32(a) for a in (srcD, srcS, srcDo, srcI, srcIo, srcIi))
        dD, dS, dDo, dI, dIo, dIi = (_i32(a) for a in (dstD, dstS, dstDo, dstI, dstIo, dstIi))
        if _kind(srcV) != _kind(dstV):
            raise RefKernelError("Invalid array types")
        nd = sD.size
        if not (sS.size == nd and dD.size == nd and dS.size == nd and sDo.size == nd + 1 and dDo.size == nd + 1 and
                sI.size == srcV.size and dI.size == dstV.size and sIi.size == nd * srcV.size and
                dIi.size == nd * dstV.size):
            raise RefKernelError("Invalid arguments")
        # the destination slice can hold at most prod(number of destination slices per dimension) new cells
        cap = dstV.size + int(np.prod([max(1, int(np.sum(sl[2::3] == d))) for d in range(nd)])) if nd else dstV.size + 1
        out = self._sparse_out(dstV.dtype, cap, nd, dD)
        rc = self._l.sstref_slice_sparse(_kind(srcV), _p(sl), sl.size, _p(srcV), srcV.size, _p(sD), _p(sS), _p(sDo), _p(sI),
                                         _p(sIo), sIo.size, _p(sIi), _p(dstV), dstV.size, _p(dD), _p(dS), _p(dDo), _p(dI),
                                         _p(dIo), dIo.size, _p(dIi), nd, _p(out[0]), out[0].size, _p(out[1]), _p(out[2]),
                                         out[2].size, _p(out[3]), out[3].size, _p(out[4]))
        return self._sparse_done(rc, out)

    # -- ImageKernel (src/org/shared/image/kernel/ImageKernel.java:56-80) -------------------------------------
    def createIntegralImage(self, srcV, srcD, srcS, dstV, dstD, dstS):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        if not (sD.size == sS.size == dD.size == dS.size):
            raise RefKernelError("Invalid arguments")
        self._chk(self._l.sstref_integral_image(_p(_f64(srcV)), srcV.size, _p(sD), _p(sS), _p(_f64(dstV)), dstV.size,
                                                _p(dD), _p(dS), sD.size))

    def createIntegralHistogram(self, srcV, srcD, srcS, memV, dstV, dstD, dstS):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        assert memV.dtype == np.int32
        if not (sD.size == sS.size and dD.size == dS.size == sD.size + 1):
            raise RefKernelError("Invalid arguments")
        self._chk(self._l.sstref_integral_histogram(_p(_f64(srcV)), srcV.size, _p(sD), _p(sS), sD.size, _p(memV),
                                                    memV.size, _p(_f64(dstV)), dstV.size, _p(dD), _p(dS)))


def load():
    """RefKernel, or None if oracle/_ref/libsst_ref.so is absent (it is built only where /root/reference exists
    and then travels with the repo snapshot)."""
    global _KERNEL
    if _KERNEL is None and os.path.exists(_PATH):
        _KERNEL = RefKernel(ctypes.CDLL(_PATH))
    return _KERNEL
