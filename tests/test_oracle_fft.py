"""Pins the CPU oracle (oracle/fftops_oracle.c) against the reference's own FFT known-answer tests
(src_test/org/shared/test/fft/ArrayFftTest.java, ConvolutionCacheTest.java; tolerance 1e-8 as
src_test/org/shared/test/Tests.java:110-123) and against numpy's pocketfft. CPU only."""
import numpy as np
import pytest

import oracle

TOL = 1e-8  # Tests.java:117


def cmul(a, b):
    """ComplexArray.eMul on interleaved flats (ElementOps.java CE_MUL)."""
    a = a.view(np.complex128)
    b = b.view(np.complex128)
    return (a * b).view(np.float64)


def conj(a):
    return np.conj(a.view(np.complex128)).view(np.float64)


def test_factors_quirk():
    # FftOps.java:111-150: even trial divisors only -> odd lengths are one radix-n DFT; 12 -> 2,2,3.
    assert oracle.factors(512) == [2, 256, 2, 128, 2, 64, 2, 32, 2, 16, 2, 8, 2, 4, 2, 2, 2, 1]
    assert oracle.factors(27) == [27, 1]
    assert oracle.factors(12) == [2, 6, 2, 3, 3, 1]
    assert oracle.factors(1) == [1, 1] or oracle.factors(1)[1] <= 1


def test_kat_1d_c2c(fft_kats):
    # ArrayFftTest.java:61-80
    a, b, exp = [np.array(r["values"]) for r in fft_kats["testOneDimensionalFft"][:3]]
    got = oracle.ifft([3], cmul(oracle.fft([3], a), oracle.fft([3], b)))
    assert np.max(np.abs(got - exp)) < TOL


def test_kat_1d_r2c(fft_kats):
    # ArrayFftTest.java:82-107
    a, b, exp = [np.array(r["values"]) for r in fft_kats["testOneDimensionalFft"][3:6]]
    got = oracle.rifft([27], cmul(oracle.rfft([27], a), oracle.rfft([27], b)))
    assert np.max(np.abs(got - exp)) < TOL


def test_kat_2d(fft_kats):
    # ArrayFftTest.java:116-160
    a, b, exp = [np.array(r["values"]) for r in fft_kats["testTwoDimensionalFft"]]
    d = [7, 7]
    got = oracle.rifft(d, cmul(oracle.rfft(d, a), conj(oracle.rfft(d, b))))
    assert np.max(np.abs(got - exp)) < TOL
    ac = np.stack([a, np.zeros_like(a)], -1).ravel()
    bc = np.stack([b, np.zeros_like(b)], -1).ravel()
    got = oracle.ifft(d, cmul(oracle.fft(d, ac), conj(oracle.fft(d, bc))))[0::2]
    assert np.max(np.abs(got - exp)) < TOL


def test_kat_3d(fft_kats):
    # ArrayFftTest.java:169-224
    a, b, exp = [np.array(r["values"]) for r in fft_kats["testThreeDimensionalFft"]]
    d = [3, 3, 3]
    ac = np.stack([a, np.zeros_like(a)], -1).ravel()
    bc = np.stack([b, np.zeros_like(b)], -1).ravel()
    got = oracle.ifft(d, cmul(oracle.fft(d, ac), oracle.fft(d, bc)))[0::2]
    assert np.max(np.abs(got - exp)) < TOL
    got = oracle.rifft(d, cmul(oracle.rfft(d, a), oracle.rfft(d, b)))
    assert np.max(np.abs(got - exp)) < TOL


def test_kat_convolve(conv_kats):
    # ConvolutionCacheTest.java:56-91 via ConvolutionCache.java:99-138: valid-region cross-correlation.
    a, b, exp = [np.array(r["values"]) for r in conv_kats["testConvolve"]]
    d = [6, 6]
    ker = np.zeros((6, 6))
    ker[:3, :3] = b.reshape(3, 3)
    kerf = conj(oracle.rfft(d, ker.ravel()))
    res = oracle.rifft(d, cmul(oracle.rfft(d, a), kerf)).reshape(6, 6)[:4, :4]
    assert np.max(np.abs(res.ravel() - exp)) < TOL


@pytest.mark.parametrize("dims", [[8], [9], [6, 5], [7, 7], [3, 3, 3], [4, 6, 5], [16, 8], [256, 256], [5, 6, 5, 6, 5]])
def test_against_numpy(dims):
    rng = np.random.default_rng(20261019)
    n = int(np.prod(dims))
    x = rng.uniform(-0.5, 0.5, 2 * n)
    want = np.fft.fftn(x.view(np.complex128).reshape(dims)).ravel().view(np.float64)
    got = oracle.fft(dims, x)
    tol = 1e-12 * max(1.0, np.log2(n))
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol
    back = oracle.ifft(dims, got)
    assert np.linalg.norm(back - x) / np.linalg.norm(x) < tol
    r = rng.uniform(-0.5, 0.5, n)
    want = np.fft.rfftn(r.reshape(dims)).ravel().view(np.float64)
    got = oracle.rfft(dims, r)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol
    back = oracle.rifft(dims, got)
    assert np.linalg.norm(back - r) / np.linalg.norm(r) < tol
