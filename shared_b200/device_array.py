"""Device-resident twins of the reference's RealArray / ComplexArray for the FFT + convolution path.

In the reference every array is a JVM `double[]` (`ProtoArray.values`, src/org/shared/array/ProtoArray.java:56), so a
provider behind the FftService / ArrayKernel SPIs must cross PCIe twice per call. These classes keep `values` in HBM
and chain device-tier calls of include/sst_cuda.h on one CUDA stream, with the reference's method names and
semantics for the methods the convolution path uses:

    RealArray    rfft, eMul/eAdd/eSub/eDiv, uAdd/uMul, aSum, rSum, subarray, map   (AbstractRealArray.java:486-652,
                                                                                   AbstractArray.java:176-318)
    ComplexArray fft, ifft, rifft, eMul, uConj, fftShift, ifftShift, subarray      (AbstractComplexArray.java:67-97)
    ConvolutionCache.convolve(cIm, ker), createCacheable, pad                      (ConvolutionCache.java:70-239,
                                                                                   FftCache.java:83-105)

`DeviceConvolutionCache.convolve` keeps conj(rfft(padded kernel)) on the device and runs the fused kernel path
(sstc_fft_convolve_dev): product on load of the first inverse pass, crop on store of the last. Arrays are row-major
(IndexingOrder.FAR, the reference's default, ArrayBase.java:66); complex arrays carry a trailing dimension of 2
(ComplexArray.java:59,82). The Java twin would be a `CudaRealArray extends ProtoArray` holding a device handle;
INTEGRATION.md sketches it. No CPU fallback: every method is a CUDA call.
"""
import ctypes
import weakref

import numpy as np

from . import _lib
from .array_kernel import COMPLEX, REAL, _i32, _pi
from .fft_service import BACKWARD, C2R, FORWARD, R2C, FftPlan

RE_ADD, RE_SUB, RE_MUL, RE_DIV = 0, 1, 2, 3          # ArrayKernel.java:208-230 (CE_* share the values)
RU_ADD, RU_MUL = 0, 1
CU_CONJ = 4
RA_SUM, RR_SUM = 0, 0


def far_strides(dims):
    s = [1] * len(dims)
    for i in range(len(dims) - 2, -1, -1):
        s[i] = s[i + 1] * dims[i + 1]
    return s


class DeviceContext:
    """One device + stream, with the plan cache the reference keeps in FftwService (FftwService.java:58-67)."""

    def __init__(self, device=0, stream=0):
        _lib.init(device)
        self.device, self.stream = int(device), ctypes.c_void_p(int(stream))
        self._plans = {}
        self.lib = _lib.lib()

    def plan(self, kind, dims):
        key = (kind, tuple(int(d) for d in dims))
        if key not in self._plans:
            self._plans[key] = FftPlan(kind, key[1], device=self.device)
        return self._plans[key]

    def sync(self):
        _lib.check(self.lib.sstc_stream_sync(self.stream))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = DeviceContext()
    return _default_ctx


class _DeviceArray:
    """Flat float64 values in HBM + logical dims (row-major)."""

    def __init__(self, dims, ctx=None, zero=True):
        self.ctx = ctx or default_context()
        self._dims = [int(d) for d in dims]
        self.n = int(np.prod(self._dims)) if self._dims else 0
        p = ctypes.c_void_p()
        _lib.check(self.ctx.lib.sstc_malloc(ctypes.byref(p), max(self.n, 1) * 8))
        self.ptr = p.value
        if zero and self.n:  # Java arrays are born zero-filled
            _lib.check(self.ctx.lib.sstc_memset(ctypes.c_void_p(self.ptr), 0, self.n * 8, self.ctx.stream))

    @classmethod
    def from_host(cls, values, dims, ctx=None, **kw):
        values = np.ascontiguousarray(values, dtype=np.float64).ravel()
        a = cls(dims, ctx=ctx, zero=False, **kw)
        if values.size != a.n:
            raise _lib.SstCudaError("Invalid dimensions and/or strides")
        _lib.check(a.ctx.lib.sstc_h2d(ctypes.c_void_p(a.ptr), values.ctypes.data, a.n * 8, a.ctx.stream))
        a.ctx.sync()  # `values` may be a temporary
        return a

    def values(self):
        """ProtoArray.values(): here a device -> host copy."""
        out = np.empty(self.n, dtype=np.float64)
        _lib.check(self.ctx.lib.sstc_d2h(out.ctypes.data, ctypes.c_void_p(self.ptr), self.n * 8, self.ctx.stream))
        self.ctx.sync()
        return out

    def dims(self):
        return list(self._dims)

    def size(self, d):
        return self._dims[d]

    def _like(self, dims=None, **kw):
        return type(self)(self._dims if dims is None else dims, ctx=self.ctx, **kw)

    def _eop(self, op, other, elem):
        if self._dims != other._dims:
            raise _lib.SstCudaError("Dimensionality mismatch")
        out = self._like(zero=False)
        _lib.check(self.ctx.lib.sstc_eop_dev(op, elem, ctypes.c_void_p(self.ptr), ctypes.c_void_p(other.ptr),
                                             ctypes.c_void_p(out.ptr), self.n, self.ctx.stream))
        return out

    def clone(self):
        out = self._like(zero=False)
        _lib.check(self.ctx.lib.sstc_d2d(ctypes.c_void_p(out.ptr), ctypes.c_void_p(self.ptr), self.n * 8, self.ctx.stream))
        return out

    def map(self, dst, *bounds):
        """ProtoArray.map(dst, srcStart, dstStart, size per dimension) (ProtoArray.java:246-259): modular block copy."""
        b, sD, dD = _i32(bounds), _i32(self._dims), _i32(dst._dims)
        sS, dS = _i32(far_strides(self._dims)), _i32(far_strides(dst._dims))
        _lib.check(self.ctx.lib.sstc_map_dev(8, _pi(b), b.size, ctypes.c_void_p(self.ptr), self.n, _pi(sD), _pi(sS),
                                             ctypes.c_void_p(dst.ptr), dst.n, _pi(dD), _pi(dS), sD.size, self.ctx.stream))
        return dst

    def subarray(self, *bounds):
        """ProtoArray.subarray(lo0, hi0, lo1, hi1, ...) (ProtoArray.java:371-391)."""
        nd = len(self._dims)
        if len(bounds) != 2 * nd:
            raise _lib.SstCudaError("Invalid subarray bounds")
        out = self._like([bounds[2 * d + 1] - bounds[2 * d] for d in range(nd)], zero=False)
        spec = []
        for d in range(nd):
            spec += [bounds[2 * d], 0, bounds[2 * d + 1] - bounds[2 * d]]
        return self.map(out, *spec)

    def slice_into(self, dst, slices):
        """ProtoArray.splice(dst, (srcIndex, dstIndex, dim) triples) (ProtoArray.java:393-437)."""
        sl, sD, dD = _i32(slices), _i32(self._dims), _i32(dst._dims)
        sS, dS = _i32(far_strides(self._dims)), _i32(far_strides(dst._dims))
        _lib.check(self.ctx.lib.sstc_slice_dev(8, _pi(sl), sl.size, ctypes.c_void_p(self.ptr), self.n, _pi(sD), _pi(sS),
                                               ctypes.c_void_p(dst.ptr), dst.n, _pi(dD), _pi(dS), sD.size,
                                               self.ctx.stream))
        return dst

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.ctx.lib.sstc_free(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class DeviceRealArray(_DeviceArray):
    def eAdd(self, o):
        return self._eop(RE_ADD, o, REAL)

    def eSub(self, o):
        return self._eop(RE_SUB, o, REAL)

    def eMul(self, o):
        return self._eop(RE_MUL, o, REAL)

    def eDiv(self, o):
        return self._eop(RE_DIV, o, REAL)

    def _uop(self, op, a):
        out = self.clone()
        _lib.check(self.ctx.lib.sstc_ruop_dev(op, float(a), ctypes.c_void_p(out.ptr), out.n, self.ctx.stream))
        return out

    def uAdd(self, a):
        return self._uop(RU_ADD, a)

    def uMul(self, a):
        return self._uop(RU_MUL, a)

    def aSum(self):
        """AbstractRealArray.aSum (raOp RA_SUM): one double back to the host."""
        d = DeviceRealArray([2], ctx=self.ctx, zero=False)
        _lib.check(self.ctx.lib.sstc_raop_dev(RA_SUM, ctypes.c_void_p(self.ptr), self.n, ctypes.c_void_p(d.ptr),
                                              self.ctx.stream))
        return float(d.values()[0])

    def rSum(self, *op_dims):
        """AbstractRealArray.rSum(dims...) (rrOp RR_SUM): reduced dims keep size 1."""
        od = [1 if d in op_dims else s for d, s in enumerate(self._dims)]
        out = DeviceRealArray(od, ctx=self.ctx, zero=False)
        sD, sS, dD, dS, ops = (_i32(self._dims), _i32(far_strides(self._dims)), _i32(od), _i32(far_strides(od)),
                               _i32(sorted(op_dims)))
        _lib.check(self.ctx.lib.sstc_rrop_dev(RR_SUM, ctypes.c_void_p(self.ptr), self.n, _pi(sD), _pi(sS),
                                              ctypes.c_void_p(out.ptr), out.n, _pi(dD), _pi(dS), sD.size, _pi(ops),
                                              ops.size, self.ctx.stream))
        return out

    def rfft(self):
        """AbstractArray.rfft (AbstractArray.java:176-187): half-spectrum dims (..., last/2+1, 2), remembers the parity
        of the last real dimension for rifft (:181, :240)."""
        dims = self._dims
        out = DeviceComplexArray(dims[:-1] + [dims[-1] // 2 + 1, 2], ctx=self.ctx, zero=False, parity=dims[-1] % 2)
        self.ctx.plan(R2C, dims).execute(self.ptr, out.ptr, self.ctx.stream.value or 0)
        return out


class DeviceComplexArray(_DeviceArray):
    def __init__(self, dims, ctx=None, zero=True, parity=None):
        if not dims or dims[-1] != 2:
            raise _lib.SstCudaError("Invalid dimensions")  # ComplexArray.java:82
        super().__init__(dims, ctx=ctx, zero=zero)
        self.parity = parity  # AbstractArray.java:63-68: odd / even length of the last real dimension, if known

    def _like(self, dims=None, **kw):
        return DeviceComplexArray(self._dims if dims is None else dims, ctx=self.ctx, parity=self.parity, **kw)

    def eAdd(self, o):
        return self._eop(RE_ADD, o, COMPLEX)

    def eSub(self, o):
        return self._eop(RE_SUB, o, COMPLEX)

    def eMul(self, o):
        return self._eop(RE_MUL, o, COMPLEX)

    def eDiv(self, o):
        return self._eop(RE_DIV, o, COMPLEX)

    def uConj(self):
        out = self.clone()
        _lib.check(self.ctx.lib.sstc_cuop_dev(CU_CONJ, 0.0, 0.0, ctypes.c_void_p(out.ptr), out.n, self.ctx.stream))
        return out

    def _c2c(self, kind):
        out = self._like(zero=False)
        self.ctx.plan(kind, self._dims[:-1]).execute(self.ptr, out.ptr, self.ctx.stream.value or 0)
        return out

    def fft(self):
        return self._c2c(FORWARD)

    def ifft(self):
        return self._c2c(BACKWARD)

    def rifftDimensions(self):
        """AbstractArray.rifftDimensions (AbstractArray.java:229-246)."""
        if self.parity is None:
            raise _lib.SstCudaError("Parity not set")
        d = self._dims[:-1]
        return d[:-1] + [(d[-1] - 1) * 2 + self.parity]

    def rifft(self):
        rd = self.rifftDimensions()
        out = DeviceRealArray(rd, ctx=self.ctx, zero=False)
        self.ctx.plan(C2R, rd).execute(self.ptr, out.ptr, self.ctx.stream.value or 0)
        return out

    def _shift(self, sign):
        # AbstractComplexArray.fftShift / ifftShift (AbstractComplexArray.java:67-97): circular shift through map
        out = self._like(zero=False)
        spec = []
        for s in self._dims[:-1]:
            spec += [0, sign * (s // 2), s]  # negative starts wrap modulo the dimension (MappingOps.java:237-249)
        spec += [0, 0, 2]
        return self.map(out, *spec)

    def fftShift(self):
        return self._shift(+1)

    def ifftShift(self):
        return self._shift(-1)


def pad(im, *margins):
    """ConvolutionCache.pad (ConvolutionCache.java:149-239): mirror-extrapolating pad. The reference issues 2^nDims
    splice calls (one per corner / edge partition); their union is one Cartesian product of per-dimension index
    lists, so a single slice call does it."""
    dims = im.dims()
    if len(dims) != len(margins):
        raise _lib.SstCudaError("Dimensionality mismatch")
    out = DeviceRealArray([s + 2 * m for s, m in zip(dims, margins)], ctx=im.ctx, zero=False)
    slices = []
    for d, (s, m) in enumerate(zip(dims, margins)):
        for j in range(m):
            slices += [m - 1 - j, j, d]                    # left margin, mirrored about the first sample
        for k in range(s):
            slices += [k, m + k, d]
        for i in range(m):
            slices += [s - 1 - i, s + m + i, d]            # right margin, mirrored about the last sample
    return im.slice_into(out, slices)


class DeviceConvolutionCache:
    """ConvolutionCache (ConvolutionCache.java:45-139) with device-resident entries. Entries are keyed by the
    kernel object and the padded dims, and die with the kernel (the reference holds them through SoftReferences,
    FftCache.java:83-105)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._cache = weakref.WeakKeyDictionary()
        self.hits = self.misses = 0

    def createCacheable(self, ker, dims):
        """conj(rfft(kernel zero-padded to dims)) (ConvolutionCache.java:117-138)."""
        padded = DeviceRealArray(dims, ctx=self.ctx)
        spec = []
        for d in range(len(dims)):
            spec += [0, 0, ker.size(d)]
        return ker.map(padded, *spec).rfft().uConj()

    def get(self, ker, dims):
        per_kernel = self._cache.setdefault(ker, {})
        key = tuple(dims)
        if key not in per_kernel:
            self.misses += 1
            per_kernel[key] = self.createCacheable(ker, dims)
        else:
            self.hits += 1
        return per_kernel[key]

    def convolve(self, c_im, ker, fused=True):
        """ConvolutionCache.convolve(ComplexArray cIm, RealArray ker) (ConvolutionCache.java:99-114): valid-region
        correlation of the image whose rfft is `c_im` with `ker`."""
        dims_t = c_im.rifftDimensions()
        ker_f = self.get(ker, dims_t)
        out_dims = [dims_t[d] - ker.size(d) + 1 for d in range(len(dims_t))]
        if min(out_dims) <= 0:
            raise _lib.SstCudaError("Invalid kernel size")
        if not fused:  # the reference's three whole-array steps, device-resident
            res = c_im.eMul(ker_f).rifft()
            bounds = []
            for d in out_dims:
                bounds += [0, d]
            return res.subarray(*bounds)
        out = DeviceRealArray(out_dims, ctx=self.ctx, zero=False)
        plan = self.ctx.plan(C2R, dims_t)
        od = _i32(out_dims)
        _lib.check(self.ctx.lib.sstc_fft_convolve_dev(plan._h, ctypes.c_void_p(c_im.ptr), ctypes.c_void_p(ker_f.ptr),
                                                      ctypes.c_void_p(out.ptr), _pi(od), self.ctx.stream))
        return out
