"""Host-side mirror of the reference's FftService SPI for the CUDA provider.

`CudaFftService` has the method names, argument meaning and error behaviour of
org.shared.fft.FftService (src/org/shared/fft/FftService.java:38-106) as implemented by
JavaFftService (src/org/shared/fft/JavaFftService.java:44-104) and FftwService
(src_extensions/org/sharedx/fftw/FftwService.java:54-239): `dims` are the logical dims (no trailing 2; the
REAL dims for rfft/rifft), arrays are flat float64 (interleaved re/im for complex), `out` is written in place
and `in_` is preserved. It is the Python twin of shared_b200/java/org/sharedx/cuda/CudaFftService.java; both
sit on the same C ABI (include/sst_cuda.h). `FftPlan` is the device tier used by bench.py and sharded.py.
"""
import ctypes

import numpy as np

from . import _lib

R2C, C2R, FORWARD, BACKWARD = 0, 1, 2, 3  # Plan.java:73-88


def _host_array(a, name, writable=False):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags.c_contiguous:
        raise TypeError("%s must be a C-contiguous float64 numpy array (a Java double[])" % name)
    if writable and not a.flags.writeable:
        raise TypeError("%s must be writable" % name)
    return a


class CudaFftService:
    """Drop-in for JavaFftService / FftwService behind ModalFftService (ModalFftService.java:39-111)."""

    def __init__(self, device=0):
        # NativeArrayKernel.java:46-49 checks Library.isInitialized(); here: a usable sm_100 device.
        _lib.init(device)
        self._hints = {}

    def _transform(self, kind, dims, in_, out):
        in_ = _host_array(in_, "in")
        out = _host_array(out, "out", writable=True)
        d = _lib.i32_array(dims)
        _lib.check(_lib.lib().sstc_fft(kind, len(dims), d, in_.ctypes.data, in_.size, out.ctypes.data, out.size))

    def rfft(self, dims, in_, out):
        """FftService.rfft (FftService.java:50): real -> half-complex on the last axis."""
        self._transform(R2C, dims, in_, out)

    def rifft(self, dims, in_, out):
        """FftService.rifft (FftService.java:62): half-complex -> real, scaled by 1/N."""
        self._transform(C2R, dims, in_, out)

    def fft(self, dims, in_, out):
        """FftService.fft (FftService.java:74): forward c2c, unscaled."""
        self._transform(FORWARD, dims, in_, out)

    def ifft(self, dims, in_, out):
        """FftService.ifft (FftService.java:86): backward c2c, scaled by 1/N."""
        self._transform(BACKWARD, dims, in_, out)

    # hint name -> (validator / action, getter); FftService.setHint / getHint (FftService.java:96-105)
    _MODES = ("estimate", "measure", "patient", "exhaustive")

    def setHint(self, name, value):
        """FftService.setHint (FftService.java:96). FFTW's "mode" / "wisdom" hints (FftwService.java:93-122) are accepted
        for compatibility (plans here are deterministic); "devices" = comma-separated CUDA device ordinals: with two or
        more, eligible 3-D c2c transforms run slab-sharded over all of them (sstc_fft_set_devices). Anything else ->
        "Unknown hint" (JavaFftService.java:96-99 raises IllegalArgumentException)."""
        if name == "mode":
            if value not in self._MODES:
                raise ValueError("Invalid execution mode")
            self._hints["mode"] = value
        elif name == "wisdom":
            self._hints["wisdom"] = ""
        elif name == "devices":
            try:
                devs = [int(v) for v in value.split(",") if v.strip() != ""]
            except ValueError:
                raise ValueError("Invalid device list") from None
            _lib.check(_lib.lib().sstc_fft_set_devices(len(devs), _lib.i32_array(devs)))
            self._hints["devices"] = ",".join(str(d) for d in devs)
        else:
            raise ValueError("Unknown hint")

    def getHint(self, name):
        """FftService.getHint (FftService.java:105)."""
        if name == "mode":
            return self._hints.get("mode", "estimate")
        if name == "wisdom":
            return ""
        if name == "devices":
            return self._hints.get("devices", "")
        raise ValueError("Unknown hint")


def transform_lengths(kind, dims):
    """(in_len, out_len) in doubles; Plan::getTransformParameters (native/src/sharedx/Plan.cpp:230-320)."""
    a, b = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(_lib.lib().sstc_fft_lengths(kind, len(dims), _lib.i32_array(dims), ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


class FftPlan:
    """Device-tier plan: executes on device pointers (e.g. torch tensor .data_ptr()) on a CUDA stream."""

    def __init__(self, kind, dims, device=0):
        _lib.init(device)
        self.kind, self.dims = kind, tuple(int(d) for d in dims)
        self.in_len, self.out_len = transform_lengths(kind, self.dims)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().sstc_fft_plan_create(ctypes.byref(h), kind, len(self.dims), _lib.i32_array(self.dims)))
        self._h = h

    def execute(self, d_in, d_out, stream=0):
        _lib.check(_lib.lib().sstc_fft_plan_execute(self._h, ctypes.c_void_p(int(d_in)), ctypes.c_void_p(int(d_out)),
                                                    ctypes.c_void_p(int(stream))))

    def set_hint(self, name, value):
        _lib.check(_lib.lib().sstc_fft_plan_set_hint(self._h, name.encode(), int(value)))

    @property
    def launches(self):
        return _lib.lib().sstc_fft_plan_launches(self._h)

    @property
    def workspace_bytes(self):
        return _lib.lib().sstc_fft_plan_workspace(self._h)

    def close(self):
        if self._h is not None and self._h.value:
            _lib.lib().sstc_fft_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fft_axis(d_in, d_out, outer, length, inner, inverse=False, scale=1.0, stream=0):
    """One axis pass on device pointers (sstc_fft_axis_dev)."""
    _lib.check(_lib.lib().sstc_fft_axis_dev(ctypes.c_void_p(int(d_in)), ctypes.c_void_p(int(d_out)), int(outer),
                                            int(length), int(inner), int(bool(inverse)), float(scale),
                                            ctypes.c_void_p(int(stream))))


def fft_axis_strided(d_in, d_out, outer, length, inner, in_plane, in_kstride, out_plane, out_kstride, inverse=False,
                     scale=1.0, stream=0):
    """One axis pass with explicit input/output layouts (sstc_fft_axis_strided_dev)."""
    _lib.check(_lib.lib().sstc_fft_axis_strided_dev(
        ctypes.c_void_p(int(d_in)), ctypes.c_void_p(int(d_out)), int(outer), int(length), int(inner), int(in_plane),
        int(in_kstride), int(out_plane), int(out_kstride), int(bool(inverse)), float(scale),
        ctypes.c_void_p(int(stream))))
