"""shared_b200 -- the B200 (sm_100a) provider for the FftService and ArrayKernel SPIs of carsomyr/Shared.

Everything numerical lives in libsst_cuda.so (hand-written CUDA behind the C ABI of include/sst_cuda.h);
this package is the thin host-side mirror of the reference's provider classes, used by the parity tests and
bench.py in place of the Java classes under shared_b200/java (no JVM exists in the build image).
"""
from ._lib import SstCudaError, init, lib  # noqa: F401
from .fft_service import (BACKWARD, C2R, FORWARD, R2C, CudaFftService, FftPlan, fft_axis,  # noqa: F401
                          fft_axis_strided, transform_lengths)

__all__ = ["CudaFftService", "FftPlan", "fft_axis", "fft_axis_strided", "transform_lengths", "SstCudaError", "init", "lib",
           "R2C", "C2R", "FORWARD", "BACKWARD"]
