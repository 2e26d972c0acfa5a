"""TEST INFRASTRUCTURE ONLY -- ctypes front end of tests/emu/_build/libsparse_emu.so: the sparse pipeline of
shared_b200/csrc/sparse_core.cuh run by a serial host executor, with the method names of ArrayKernel."""
import ctypes
import os
import shutil
import subprocess

import numpy as np

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


class _State(ctypes.Structure):
    _fields_ = [("count", ctypes.c_int32), ("n_offsets", ctypes.c_int32), ("n_indirections", ctypes.c_int32),
                ("elem_bytes", ctypes.c_int32), ("values", ctypes.c_void_p), ("indices", ctypes.c_void_p),
                ("offsets", ctypes.c_void_p), ("indirections", ctypes.c_void_p)]


def build():
    """-> (build dir, make's output); raises RuntimeError when the tool chain is missing."""
    if shutil.which("g++") is None or shutil.which("make") is None:
        raise RuntimeError("no g++ / make")
    res = subprocess.run(["make", "-s", "-k", "-C", EMU_DIR, "all"], capture_output=True, text=True)
    return os.path.join(EMU_DIR, "_build"), res.stdout[-2000:] + res.stderr[-2000:]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class EmuError(RuntimeError):
    pass


class EmuSparseKernel:
    def __init__(self, path):
        self._l = ctypes.CDLL(path)
        self._l.emu_last_error.restype = ctypes.c_char_p

    def _unpack(self, rc, st, dtype):
        if rc:
            raise EmuError(self._l.emu_last_error().decode())

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            return np.frombuffer((ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt).copy()
        return (arr(st.values, st.count, dtype), arr(st.indices, st.count, np.int32),
                arr(st.offsets, st.n_offsets, np.int32), arr(st.indirections, st.n_indirections, np.int32))

    def insertSparse(self, oldV, oldD, oldS, oldDo, oldI, newV, newI):
        oD, oS, oDo, oI = _i32(oldD), _i32(oldS), _i32(oldDo), _i32(oldI)
        st = _State()
        rc = self._l.emu_insert_sparse(oldV.dtype.itemsize, _p(oldV), ctypes.c_int64(oldV.size), _p(oD), _p(oS), _p(oDo),
                                       oD.size, _p(oI), _p(newV), ctypes.c_int64(newV.size), _p(newI), ctypes.byref(st))
        return self._unpack(rc, st, oldV.dtype)

    def sliceSparse(self, slices, srcV, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, dstD, dstS, dstDo, dstI, dstIo,
                    dstIi):
        a = [_i32(x) for x in (slices, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstD, dstS, dstDo, dstI, dstIo, dstIi)]
        st = _State()
        rc = self._l.emu_slice_sparse(srcV.dtype.itemsize, _p(a[0]), a[0].size, _p(srcV), ctypes.c_int64(srcV.size), _p(a[1]),
                                      _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]), _p(dstV), ctypes.c_int64(dstV.size),
                                      _p(a[7]), _p(a[8]), _p(a[9]), _p(a[10]), _p(a[11]), _p(a[12]), a[1].size,
                                      ctypes.byref(st))
        return self._unpack(rc, st, srcV.dtype)
