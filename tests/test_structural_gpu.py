"""The dedicated structural entry points (sstc_shift_dev, sstc_fftshift_dev, sstc_transpose_dev, sstc_pad_dev; SURVEY.md
8 f-3): each must equal what the reference's JVM-side caller produces with ONE map / slice call -- checked against numpy's
definition of the same operation (bit-exact: pure data movement), against the reference's own native kernel driven through
shared_b200.proto_array.ProtoArray with the reference's bounds arithmetic, and against the literal fftShift vectors of
ArrayFftTest (src_test/org/shared/test/fft/ArrayFftTest.java)."""
import ctypes

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ref = oracle.ref()


@pytest.fixture(scope="module")
def dev():
    import torch
    import shared_b200
    shared_b200.init(0)
    return torch


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def call(fn, *args):
    from shared_b200 import _lib
    _lib.check(getattr(_lib.lib(), fn)(*args))


def i32(v):
    from shared_b200 import _lib
    return _lib.i32_array(v)


@pytest.mark.parametrize("dims,shifts", [([7], [3]), ([5, 9], [2, -4]), ([4, 6, 5], [-1, 7, 2]), ([300, 257], [150, 128]),
                                           ([2048, 1024], [1024, 512])])
@pytest.mark.parametrize("near", [0, 1])
def test_shift(dev, dims, shifts, near):
    torch = dev
    x = np.random.default_rng(1).uniform(-1, 1, dims)
    flat = x.ravel(order="F" if near else "C")
    d_src = torch.from_numpy(flat.copy()).cuda()
    d_dst = torch.empty_like(d_src)
    call("sstc_shift_dev", 8, ptr(d_src), ptr(d_dst), i32(dims), len(dims), near, i32(shifts), None)
    want = np.roll(x, shifts, axis=tuple(range(len(dims)))).ravel(order="F" if near else "C")
    assert np.array_equal(d_dst.cpu().numpy(), want)
    if ref is not None:  # the reference's own kernel, driven with the reference's bounds arithmetic
        from shared_b200.proto_array import ProtoArray
        a = ProtoArray(ref, flat.copy(), dims, "NEAR" if near else "FAR")
        assert np.array_equal(a.shift(*shifts).values, want)


def test_fftshift(dev):
    """fftShift = shift by dims/2 of every dimension but the complex pair (AbstractComplexArray.java:67-97); the literal
    vectors of ArrayFftTest.testFftShift are replayed through the same map kernel by tests/test_proto_array_gpu.py."""
    torch = dev
    rng = np.random.default_rng(2)
    for dims in ([8, 2], [5, 6, 2], [4, 3, 5, 2], [512, 256, 2]):
        x = rng.uniform(-1, 1, dims)
        d_src = torch.from_numpy(x.ravel().copy()).cuda()
        d_dst = torch.empty_like(d_src)
        back = torch.empty_like(d_src)
        call("sstc_fftshift_dev", ptr(d_src), ptr(d_dst), i32(dims), len(dims), 0, None)
        axes = tuple(range(len(dims) - 1))
        assert np.array_equal(d_dst.cpu().numpy(), np.fft.fftshift(x, axes=axes).ravel())
        call("sstc_fftshift_dev", ptr(d_dst), ptr(back), i32(dims), len(dims), 1, None)
        assert np.array_equal(back.cpu().numpy(), x.ravel())


@pytest.mark.parametrize("dims,perm", [([3, 4], [1, 0]), ([3, 4, 5], [2, 0, 1]), ([2, 3, 4, 5], [3, 1, 0, 2]),
                                         ([1000, 700], [1, 0]), ([64, 100, 36], [1, 2, 0])])
@pytest.mark.parametrize("near", [0, 1])
def test_transpose(dev, dims, perm, near):
    torch = dev
    order = "F" if near else "C"
    x = np.random.default_rng(3).integers(-1000, 1000, dims).astype(np.int32)
    d_src = torch.from_numpy(x.ravel(order=order).copy()).cuda()
    d_dst = torch.empty_like(d_src)
    call("sstc_transpose_dev", 4, ptr(d_src), ptr(d_dst), i32(dims), len(dims), near, i32(perm), None)
    # source dimension d becomes destination dimension perm[d]
    want = np.transpose(x, np.argsort(perm)).ravel(order=order)
    assert np.array_equal(d_dst.cpu().numpy(), want)
    if ref is not None:
        from shared_b200.proto_array import ProtoArray
        a = ProtoArray(ref, x.ravel(order=order).copy(), dims, "NEAR" if near else "FAR", dtype=np.int32)
        assert np.array_equal(a.transpose(*perm).values, want)
    from shared_b200 import _lib
    rc = _lib.lib().sstc_transpose_dev(4, ptr(d_src), ptr(d_dst), i32(dims), len(dims), near, i32([0] * len(dims)), None)
    assert (rc != 0 and b"Invalid permutation" in _lib.lib().sstc_last_error()) or len(dims) == 1


@pytest.mark.parametrize("dims,margins", [([5], [2]), ([4, 6], [1, 3]), ([3, 4, 5], [3, 0, 2]), ([1000, 900], [15, 15])])
def test_pad(dev, dims, margins):
    torch = dev
    x = np.random.default_rng(4).uniform(-1, 1, dims)
    d_src = torch.from_numpy(x.ravel().copy()).cuda()
    out_dims = [s + 2 * m for s, m in zip(dims, margins)]
    d_dst = torch.full((int(np.prod(out_dims)),), float("nan"), dtype=torch.float64, device="cuda")
    call("sstc_pad_dev", ptr(d_src), ptr(d_dst), i32(dims), len(dims), i32(margins), None)
    # ConvolutionCache.pad mirrors about the first / last sample, the edge sample included ("symmetric")
    want = np.pad(x, [(m, m) for m in margins], mode="symmetric")
    assert np.array_equal(d_dst.cpu().numpy().reshape(out_dims), want)
