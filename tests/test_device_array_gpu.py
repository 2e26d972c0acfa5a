"""Device-resident arrays and the fused convolution against the reference's ConvolutionCacheTest
known answers and the CPU oracle (oracle.rfft / rifft = restated JavaFftService) on seeded inputs.
Tolerance: the reference's 1e-8 for its KATs; relative L2 <= 1e-12*log2(N) against the oracle; exact for data
movement (pad, subarray, fftShift)."""
import math

import numpy as np
import pytest

import oracle
from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import shared_b200.device_array as m
    return m


def rel(a, b):
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def tol(n):
    return 1e-12 * max(1.0, math.log2(max(n, 2)))


def oracle_convolve(dims, image, kdims, kernel):
    """ConvolutionCache.convolve (ConvolutionCache.java:99-138) through the oracle's rfft / rifft."""
    padded = np.zeros(dims)
    padded[tuple(slice(0, k) for k in kdims)] = kernel.reshape(kdims)
    kf = oracle.rfft(dims, padded.ravel()).reshape(-1, 2)
    kf[:, 1] *= -1.0
    im = oracle.rfft(dims, image).reshape(-1, 2)
    prod = np.empty_like(im)
    prod[:, 0] = im[:, 0] * kf[:, 0] - im[:, 1] * kf[:, 1]
    prod[:, 1] = im[:, 0] * kf[:, 1] + im[:, 1] * kf[:, 0]
    full = oracle.rifft(dims, prod.ravel()).reshape(dims)
    return full[tuple(slice(0, d - k + 1) for d, k in zip(dims, kdims))].ravel()


def test_kat_convolve(da):
    a, b, exp = golden("ConvolutionCacheTest")["testConvolve"]  # ConvolutionCacheTest.java:56-91
    cache = da.DeviceConvolutionCache()
    im = da.DeviceRealArray.from_host(a["values"], a["dims"])
    ker = da.DeviceRealArray.from_host(b["values"], b["dims"])
    for fused in (True, False):
        res = cache.convolve(im.rfft(), ker, fused=fused)
        assert res.dims() == exp["dims"]
        assert np.max(np.abs(res.values() - np.array(exp["values"]))) < 1e-8
    assert cache.misses == 1 and cache.hits == 1  # the kernel spectrum is computed once (FftCache.java:83-105)


def test_kat_pad(da):
    recs = golden("ConvolutionCacheTest")["testPad"]  # ConvolutionCacheTest.java:96-175 (Arrays.equals: exact)
    for (src, exp), margins in zip(((recs[0], recs[1]), (recs[2], recs[3])), ((1, 1, 1), (2, 1))):
        got = da.pad(da.DeviceRealArray.from_host(src["values"], src["dims"]), *margins)
        assert got.dims() == exp["dims"]
        assert np.array_equal(got.values(), np.array(exp["values"], dtype=np.float64))


CONV = [([64], [5]), ([63], [4]), ([32, 48], [3, 7]), ([30, 45], [30, 1]), ([16, 31], [2, 2]), ([8, 16, 12], [3, 1, 5]),
        ([2, 3, 4, 5, 6], [1, 2, 1, 2, 3]), ([256, 256], [31, 31]), ([6, 4096], [2, 33]), ([4096, 6], [9, 2]),
        # an axis the tile kernels do not hold: the unfused fallback inside sstc_fft_convolve_dev
        ([3, 8192], [2, 9]), ([10007], [11])]


@pytest.mark.parametrize("dims,kdims", CONV, ids=lambda d: "x".join(map(str, d)))
def test_convolve_matches_oracle(da, dims, kdims):
    rng = np.random.default_rng(sum(dims) + len(kdims))
    image, kernel = rng.uniform(-0.5, 0.5, int(np.prod(dims))), rng.uniform(-1, 1, int(np.prod(kdims)))
    want = oracle_convolve(dims, image, kdims, kernel)
    cache = da.DeviceConvolutionCache()
    c_im = da.DeviceRealArray.from_host(image, dims).rfft()
    keep = c_im.values()
    ker = da.DeviceRealArray.from_host(kernel, kdims)
    fused = cache.convolve(c_im, ker)
    plain = cache.convolve(c_im, ker, fused=False)
    n = int(np.prod(dims))
    assert fused.dims() == [d - k + 1 for d, k in zip(dims, kdims)]
    assert rel(fused.values(), want) < tol(n)
    assert rel(plain.values(), want) < tol(n)
    assert np.array_equal(c_im.values(), keep)  # the image spectrum is an input, not scratch


def test_array_methods(da):
    rng = np.random.default_rng(3)
    for dims in ([7], [6, 5], [4, 6, 5], [16, 9]):
        n = int(np.prod(dims))
        x, y = rng.uniform(-1, 1, n), rng.uniform(0.5, 1.5, n)
        a, b = da.DeviceRealArray.from_host(x, dims), da.DeviceRealArray.from_host(y, dims)
        assert np.array_equal(a.eMul(b).values(), x * y) and np.array_equal(a.eDiv(b).values(), x / y)
        assert np.array_equal(a.eAdd(b).uMul(2.0).values(), (x + y) * 2.0)
        assert abs(a.aSum() - x.sum()) < 1e-12 * n
        assert rel(a.rSum(len(dims) - 1).values(), x.reshape(dims).sum(-1).ravel()) < tol(n)
        spec = a.rfft()
        assert spec.dims() == dims[:-1] + [dims[-1] // 2 + 1, 2] and spec.parity == dims[-1] % 2
        assert rel(spec.values(), oracle.rfft(dims, x)) < tol(n)
        assert spec.rifft().dims() == dims and rel(spec.rifft().values(), x) < tol(n)
        z = rng.uniform(-1, 1, 2 * n)
        c = da.DeviceComplexArray.from_host(z, dims + [2])
        assert rel(c.fft().values(), oracle.fft(dims, z)) < tol(n)
        assert rel(c.fft().ifft().values(), z) < tol(n)
        zc = z.reshape(-1, 2) * np.array([1.0, -1.0])
        assert np.array_equal(c.uConj().values(), zc.ravel())
        zz = (z[0::2] + 1j * z[1::2]).reshape(dims)
        axes = tuple(range(len(dims)))
        for got, want in ((c.fftShift(), np.fft.fftshift(zz, axes)), (c.ifftShift(), np.fft.ifftshift(zz, axes)),
                          (c.fftShift().ifftShift(), zz)):
            g = got.values()
            assert np.array_equal(g[0::2] + 1j * g[1::2], want.ravel())
        lo = [int(rng.integers(0, d)) for d in dims]
        hi = [l + 1 + int(rng.integers(0, d - l)) for l, d in zip(lo, dims)]
        bounds = [v for pair in zip(lo, hi) for v in pair]
        assert np.array_equal(a.subarray(*bounds).values(),
                              x.reshape(dims)[tuple(slice(l, h) for l, h in zip(lo, hi))].ravel())


def test_errors(da):
    import shared_b200 as s
    cache = da.DeviceConvolutionCache()
    im = da.DeviceRealArray.from_host(np.zeros(16), [4, 4]).rfft()
    with pytest.raises(s.SstCudaError, match="Invalid kernel size"):
        cache.convolve(im, da.DeviceRealArray.from_host(np.zeros(10), [5, 2]))
    with pytest.raises(s.SstCudaError, match="Parity not set"):
        da.DeviceComplexArray([4, 3, 2]).rifft()
    with pytest.raises(s.SstCudaError, match="Dimensionality mismatch"):
        da.pad(da.DeviceRealArray([3, 3]), 1)


def test_c2_shape_fused_vs_plain(da):
    """BASELINE configs[1]: 4096 x 4096 image, 31 x 31 kernel -- fused and three-step paths agree."""
    import torch
    rng = np.random.default_rng(9)
    n = 4096
    image = da.DeviceRealArray.from_host(rng.uniform(-0.5, 0.5, n * n), [n, n])
    ker = da.DeviceRealArray.from_host(rng.uniform(-1, 1, 31 * 31), [31, 31])
    cache = da.DeviceConvolutionCache()
    c_im = image.rfft()
    a, b = cache.convolve(c_im, ker).values(), cache.convolve(c_im, ker, fused=False).values()
    assert a.size == 4066 * 4066 and rel(a, b) < 1e-13
    # spot-check a few outputs against the definition: out[i, j] = sum_{u, v} image[i + u, j + v] * ker[u, v]
    im_h, k_h = image.values().reshape(n, n), ker.values().reshape(31, 31)
    out = a.reshape(4066, 4066)
    for (i, j) in ((0, 0), (17, 4000), (4065, 4065), (2048, 1)):
        assert abs(out[i, j] - (im_h[i:i + 31, j:j + 31] * k_h).sum()) < 1e-9
    del torch
