"""Host-side mirror of the reference's ImageKernel SPI for the CUDA provider.

`CudaImageKernel` has the method names, argument order and error behaviour of
org.shared.image.kernel.ImageKernel (src/org/shared/image/kernel/ImageKernel.java:56-80) as implemented by
NativeImageKernel (src/org/shared/image/jni/NativeImageKernel.java:40-60): flat numpy float64 / int32 arrays stand
in for Java double[] / int[], dst is mutated in place. It is the Python twin of
shared_b200/java/org/sharedx/cuda/CudaImageKernel.java; both sit on the host tier of include/sst_cuda.h.
`DeviceImageKernel` is the device tier (device pointers + stream). `ilut` / `query` restate the host-side lookup
of IntegralImage.query / IntegralHistogram.query (IntegralImage.java:66-87, ImageOps.java:51-66) so that the
parity tests can replay the reference's own IntegralImageTest / IntegralHistogramTest.
"""
import ctypes

import numpy as np

from . import _lib
from .array_kernel import _i32, _pi, _vp


class CudaImageKernel:
    """Drop-in for JavaImageKernel / NativeImageKernel behind ModalImageKernel (ModalImageKernel.java:86-97)."""

    def __init__(self, device=0):
        _lib.init(device)
        self._l = _lib.lib()

    def createIntegralImage(self, srcV, srcD, srcS, dstV, dstD, dstS):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        if not (sD.size == sS.size == dD.size == dS.size) or srcV.dtype != np.float64 or dstV.dtype != np.float64:
            raise _lib.SstCudaError("Invalid arguments")  # NativeImageKernel.cpp:38-50
        _lib.check(self._l.sstc_integral_image(_vp(srcV), srcV.size, _pi(sD), _pi(sS), _vp(dstV), dstV.size, _pi(dD),
                                               _pi(dS), sD.size))

    def createIntegralHistogram(self, srcV, srcD, srcS, memV, dstV, dstD, dstS):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        if not (sD.size == sS.size and dD.size == dS.size == sD.size + 1) or memV.dtype != np.int32 or \
                srcV.dtype != np.float64 or dstV.dtype != np.float64:
            raise _lib.SstCudaError("Invalid arguments")  # NativeImageKernel.cpp:136-150
        _lib.check(self._l.sstc_integral_histogram(_vp(srcV), srcV.size, _pi(sD), _pi(sS), sD.size, _vp(memV), memV.size,
                                                   _vp(dstV), dstV.size, _pi(dD), _pi(dS)))


class DeviceImageKernel:
    """Device tier: value arrays are device pointers (ints), metadata stays on the host; asynchronous on `stream`."""

    def __init__(self, device=0, stream=0):
        _lib.init(device)
        self._l = _lib.lib()
        self.stream = ctypes.c_void_p(int(stream))

    def createIntegralImage(self, src, nsrc, srcD, srcS, dst, ndst, dstD, dstS):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        _lib.check(self._l.sstc_integral_image_dev(ctypes.c_void_p(src), int(nsrc), _pi(sD), _pi(sS), ctypes.c_void_p(dst),
                                                   int(ndst), _pi(dD), _pi(dS), sD.size, self.stream))

    def createIntegralHistogram(self, src, nsrc, srcD, srcS, mem, dst, ndst, dstD, dstS):
        sD, sS, dD, dS = _i32(srcD), _i32(srcS), _i32(dstD), _i32(dstS)
        _lib.check(self._l.sstc_integral_histogram_dev(ctypes.c_void_p(src), int(nsrc), _pi(sD), _pi(sS), sD.size,
                                                       ctypes.c_void_p(mem), int(nsrc), ctypes.c_void_p(dst), int(ndst),
                                                       _pi(dD), _pi(dS), self.stream))


def ilut(nd):
    """ImageOps.createIlut (ImageOps.java:51-66): corner selectors and inclusion-exclusion signs."""
    stride = nd + 1
    t = np.zeros((1 << nd) * stride, dtype=np.int64)
    parity = nd % 2
    for i in range(1 << nd):
        for d in range(nd):
            t[i * stride + d] = (d << 1) + ((i >> d) & 1)
        t[i * stride + nd] = 1 - (((bin(i).count("1") + parity) % 2) << 1)
    return t


def query(values, strides, bounds, bin_offset=0):
    """IntegralImage.query (IntegralImage.java:66-87): sum over [lo, hi) per dimension, bounds = (lo0, hi0, lo1, ...)."""
    nd = len(bounds) // 2
    t, stride, total = ilut(nd), nd + 1, 0.0
    for i in range(1 << nd):
        index = bin_offset
        for d in range(nd):
            index += bounds[t[i * stride + d]] * strides[d]
        total += values[index] * t[i * stride + nd]
    return total
