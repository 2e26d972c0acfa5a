"""Shared by the CPU (oracle) and GPU (CUDA provider) ImageKernel tests: a replay of the reference's
IntegralImageTest.java:55-99 and IntegralHistogramTest.java:58-135 against any object with the ImageKernel methods
(region sums looked up through the integral structure must equal naive sums of the region, abs 1e-8)."""
import numpy as np

from shared_b200.image_kernel import query


def strides(dims, order):
    n = len(dims)
    s = [1] * n
    if order == "FAR":
        for i in range(n - 2, -1, -1):
            s[i] = s[i + 1] * dims[i + 1]
    else:
        for i in range(1, n):
            s[i] = s[i - 1] * dims[i - 1]
    return s


def as_nd(flat, dims, order):
    """Logical n-D view of a flat FAR / NEAR array."""
    return flat.reshape(dims) if order == "FAR" else flat.reshape(dims[::-1]).transpose(range(len(dims) - 1, -1, -1))


def replay_integral_image(kernel, rng, base=64, trials=3, queries=128):
    for _ in range(trials):
        for nd in (1, 2, 3):
            dims = [base + int(rng.integers(base)) for _ in range(nd)]
            order = "FAR" if rng.integers(2) == 0 else "NEAR"
            n = int(np.prod(dims))
            mat = np.arange(n, dtype=np.float64) * (1.0 / n)  # Arithmetic.doubleRange(n).uMul(1/n)
            ddims = [d + 1 for d in dims]
            dst = np.zeros(int(np.prod(ddims)))
            ds = strides(ddims, order)
            kernel.createIntegralImage(mat, dims, strides(dims, order), dst, ddims, ds)
            view = as_nd(mat, dims, order)
            for _q in range(queries):
                b = []
                for d in range(nd):
                    lo = int(rng.integers(dims[d]))
                    b += [lo, lo + int(rng.integers(dims[d] - lo)) + 1]
                want = view[tuple(slice(b[2 * d], b[2 * d + 1]) for d in range(nd))].sum()
                assert abs(query(dst, ds, b) - want) < 1e-8, (dims, order, b)


def replay_integral_histogram(kernel, rng, base=32, trials=3, queries=64):
    for _ in range(trials):
        for nd in (1, 2, 3):
            dims = [base + int(rng.integers(base)) for _ in range(nd)]
            order = "FAR" if rng.integers(2) == 0 else "NEAR"
            n = int(np.prod(dims))
            nbins = 4 + int(rng.integers(4))
            mat = np.arange(n, dtype=np.float64) * (1.0 / n)
            mem = rng.integers(nbins, size=n).astype(np.int32)
            ddims = [d + 1 for d in dims] + [nbins]
            ds = strides(ddims, order)
            dst = np.zeros(int(np.prod(ddims)))
            kernel.createIntegralHistogram(mat, dims, strides(dims, order), mem, dst, ddims, ds)
            view, mview = as_nd(mat, dims, order), as_nd(mem, dims, order)
            for _q in range(queries):
                b = []
                for d in range(nd):
                    lo = int(rng.integers(dims[d]))
                    b += [lo, lo + int(rng.integers(dims[d] - lo)) + 1]
                sl = tuple(slice(b[2 * d], b[2 * d + 1]) for d in range(nd))
                for bin_ in range(nbins):
                    want = view[sl][mview[sl] == bin_].sum()
                    got = query(dst, ds, b, bin_offset=bin_ * ds[nd])
                    assert abs(got - want) < 1e-8, (dims, order, b, bin_)
