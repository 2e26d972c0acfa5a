"""Parity of the CUDA ImageKernel provider (through the C ABI host tier) with the reference's own
NativeImageKernel (oracle.ref()) on seeded inputs, plus a replay of the reference's IntegralImageTest /
IntegralHistogramTest. Bar: exact on integer-valued inputs (every summation order is exact there), relative L2 <=
1e-12*log2(n) otherwise (the CUDA scans add in a tree order, the reference left to right)."""
import math

import numpy as np
import pytest

import oracle
from image_cases import replay_integral_histogram, replay_integral_image, strides

pytestmark = pytest.mark.gpu
ref = oracle.ref()
needs_ref = pytest.mark.skipif(ref is None or not hasattr(ref, "createIntegralImage"), reason="oracle/_ref missing")


@pytest.fixture(scope="module")
def k():
    from shared_b200.image_kernel import CudaImageKernel
    return CudaImageKernel()


def rel(a, b):
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def test_replay_reference_tests(k):
    replay_integral_image(k, np.random.default_rng(11))           # IntegralImageTest.java:55-99
    replay_integral_histogram(k, np.random.default_rng(12))       # IntegralHistogramTest.java:58-135


CASES = [([1], "FAR"), ([7], "FAR"), ([300], "FAR"), ([5, 9], "FAR"), ([5, 9], "NEAR"), ([64, 33], "FAR"),
         ([257, 300], "NEAR"), ([3, 4, 5], "FAR"), ([3, 4, 5], "NEAR"), ([20, 31, 17], "FAR"), ([2, 3, 2, 4], "NEAR"),
         ([1, 6], "FAR"), ([6, 1], "NEAR"),
         # enough lines off dimension 0 for the fused copy + first scan
         ([70, 1500], "FAR"), ([40, 30, 50], "FAR"), ([300, 40, 33], "FAR"), ([3, 2000], "FAR"), ([1500, 70], "NEAR"),
         # 2-D images large enough for the single-pass tile scan (array_scan2d.cu): ragged tiles, both orders, one thin
         ([1000, 1030], "FAR"), ([513, 2049], "NEAR"), ([63, 1100], "FAR"), ([4097, 65], "FAR")]


@needs_ref
@pytest.mark.parametrize("dims,order", CASES)
def test_integral_image_matches_reference(k, dims, order):
    rng = np.random.default_rng(sum(dims))
    n = int(np.prod(dims))
    ddims = [d + 1 for d in dims]
    nd_ = int(np.prod(ddims))
    for integer in (True, False):
        src = rng.integers(-3, 4, n).astype(np.float64) if integer else rng.uniform(-1, 1, n)
        # dst arrives non-zero: the border takes part in the sums exactly as the reference leaves it
        fill = rng.integers(-2, 3, nd_).astype(np.float64) if integer else rng.uniform(-1, 1, nd_)
        a, b = fill.copy(), fill.copy()
        k.createIntegralImage(src, dims, strides(dims, order), a, ddims, strides(ddims, order))
        ref.createIntegralImage(src, dims, strides(dims, order), b, ddims, strides(ddims, order))
        if integer:
            assert np.array_equal(a, b), (dims, order)
        else:
            assert rel(a, b) < 1e-12 * max(1.0, math.log2(nd_)), (dims, order)


@needs_ref
@pytest.mark.parametrize("dims,order", CASES)
def test_integral_histogram_matches_reference(k, dims, order):
    rng = np.random.default_rng(100 + sum(dims))
    n = int(np.prod(dims))
    nbins = 1 + int(rng.integers(6))
    ddims = [d + 1 for d in dims] + [nbins]
    nd_ = int(np.prod(ddims))
    src = rng.integers(-3, 4, n).astype(np.float64)
    mem = rng.integers(nbins, size=n).astype(np.int32)
    fill = rng.integers(-2, 3, nd_).astype(np.float64)
    a, b = fill.copy(), fill.copy()
    k.createIntegralHistogram(src, dims, strides(dims, order), mem, a, ddims, strides(ddims, order))
    ref.createIntegralHistogram(src, dims, strides(dims, order), mem, b, ddims, strides(ddims, order))
    assert np.array_equal(a, b), (dims, order, nbins)


def test_error_messages(k):
    import shared_b200 as s
    E = s.SstCudaError
    with pytest.raises(E, match="Dimension mismatch"):
        k.createIntegralImage(np.zeros(6), [2, 3], [3, 1], np.zeros(12), [4, 3], [3, 1])
    with pytest.raises(E, match="Invalid dimensions and/or strides"):
        k.createIntegralImage(np.zeros(6), [2, 3], [1, 1], np.zeros(12), [3, 4], [4, 1])
    with pytest.raises(E, match="Invalid arguments"):
        k.createIntegralImage(np.zeros(6), [2, 3], [3, 1], np.zeros(12), [3, 4], [4])
    dst = np.full(6, 9.0)
    with pytest.raises(E, match="Invalid membership index"):
        k.createIntegralHistogram(np.ones(2), [2], [1], np.array([0, 5], dtype=np.int32), dst, [3, 2], [2, 1])
    with pytest.raises(E, match="Invalid arguments"):
        k.createIntegralHistogram(np.ones(2), [2], [1], np.array([0], dtype=np.int32), dst, [3, 2], [2, 1])
    # empty source: validated, then a no-op (NativeImageKernel.cpp:82-84)
    d1 = np.full(3, 5.0)
    k.createIntegralImage(np.zeros(0), [0, 2], [2, 1], d1, [1, 3], [3, 1])
    assert np.array_equal(d1, np.full(3, 5.0))


def test_device_tier_8192_square():
    """BASELINE configs[3] shape through the device tier: totals and random region queries against torch sums."""
    import torch
    from shared_b200.image_kernel import DeviceImageKernel, query
    n = 8192
    x = torch.randint(-3, 4, (n, n), device="cuda").to(torch.float64)
    out = torch.zeros((n + 1) * (n + 1), dtype=torch.float64, device="cuda")
    dk = DeviceImageKernel(0, torch.cuda.current_stream().cuda_stream)
    dk.createIntegralImage(x.data_ptr(), n * n, [n, n], [n, 1], out.data_ptr(), out.numel(), [n + 1, n + 1], [n + 1, 1])
    torch.cuda.synchronize()
    ii = out.view(n + 1, n + 1)
    assert torch.all(ii[0] == 0) and torch.all(ii[:, 0] == 0)
    assert ii[n, n].item() == x.sum().item()
    want = torch.cumsum(torch.cumsum(x, 0), 1)
    assert torch.equal(ii[1:, 1:], want)
    host = out.cpu().numpy()
    rng = np.random.default_rng(3)
    for _ in range(16):
        r0, c0 = int(rng.integers(n)), int(rng.integers(n))
        r1, c1 = r0 + 1 + int(rng.integers(n - r0)), c0 + 1 + int(rng.integers(n - c0))
        assert query(host, [n + 1, 1], [r0, r1, c0, c1]) == x[r0:r1, c0:c1].sum().item()
