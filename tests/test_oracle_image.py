"""Pins the ImageKernel oracle -- the reference's own NativeImageKernel.cpp compiled from /root/reference against
the shim (oracle/_ref/libsst_ref.so) -- to the reference's IntegralImageTest / IntegralHistogramTest, and to the
closed form (zero border + n-D cumulative sum). CPU only."""
import numpy as np
import pytest

import oracle
from image_cases import as_nd, replay_integral_histogram, replay_integral_image, strides

ref = oracle.ref()
pytestmark = pytest.mark.skipif(ref is None or not hasattr(ref, "createIntegralImage"),
                                reason="oracle/_ref/libsst_ref.so not built (needs /root/reference)")


def test_integral_image_test_replay():
    replay_integral_image(ref, np.random.default_rng(5), base=24, trials=2, queries=48)  # IntegralImageTest.java:55-99


def test_integral_histogram_test_replay():
    replay_integral_histogram(ref, np.random.default_rng(6), base=12, trials=2, queries=16)  # IntegralHistogramTest.java:58


def test_closed_form():
    rng = np.random.default_rng(7)
    for dims, order in (([5], "FAR"), ([4, 7], "FAR"), ([4, 7], "NEAR"), ([3, 5, 4], "FAR"), ([3, 5, 4], "NEAR")):
        src = rng.integers(-4, 5, int(np.prod(dims))).astype(np.float64)  # small integers: every order of sums is exact
        ddims = [d + 1 for d in dims]
        dst = np.zeros(int(np.prod(ddims)))
        ref.createIntegralImage(src, dims, strides(dims, order), dst, ddims, strides(ddims, order))
        want = np.zeros(ddims)
        want[tuple(slice(1, None) for _ in dims)] = as_nd(src, dims, order)
        for ax in range(len(dims)):
            want = np.cumsum(want, axis=ax)
        assert np.array_equal(as_nd(dst, ddims, order), want)


def test_error_texts():
    E = oracle.refkernel.RefKernelError
    with pytest.raises(E, match="Dimension mismatch"):
        ref.createIntegralImage(np.zeros(6), [2, 3], [3, 1], np.zeros(12), [3, 4][::-1], [3, 1])
    with pytest.raises(E, match="Invalid membership index"):
        ref.createIntegralHistogram(np.ones(2), [2], [1], np.array([0, 5], dtype=np.int32), np.zeros(6), [3, 2], [2, 1])
