"""The staging copy of the FftService host tier (shared_b200/csrc/fft_host.cu: copy threads, non-temporal AVX2 stores with
an aligned main loop and memcpy head / tail) needs no GPU: every alignment of source and destination, sizes around the
vector width and the task grain, pitched rows, and several callers at once -- against numpy."""
import ctypes
import threading

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    from shared_b200 import _lib
    return _lib.lib()


def copy(lib, dst, doff, dpitch, src, soff, spitch, width, rows):
    rc = lib.sstc_exp_host_copy(ctypes.c_void_p(dst.ctypes.data + doff), dpitch, ctypes.c_void_p(src.ctypes.data + soff), spitch,
                                width, rows)
    assert rc == 0, lib.sstc_last_error()


def test_every_alignment_and_ragged_sizes(lib):
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, 1 << 16, dtype=np.uint8)
    for size in (0, 1, 15, 63, 64, 127, 128, 129, 4095, 4096, 4097, 9999, 40000):
        for soff in (0, 1, 8, 31, 63):
            for doff in (0, 3, 16, 33, 63):
                dst = np.full(size + 200, 0xAB, dtype=np.uint8)
                copy(lib, dst, doff, size, src, soff, size, size, 1)
                assert np.array_equal(dst[doff:doff + size], src[soff:soff + size]), (size, soff, doff)
                assert (dst[:doff] == 0xAB).all() and (dst[doff + size:] == 0xAB).all(), "wrote outside the destination"


def test_large_copy_spread_over_threads(lib):
    rng = np.random.default_rng(1)
    n = (37 << 20) + 12345  # many 2 MiB tasks and a ragged tail
    src = rng.integers(0, 256, n, dtype=np.uint8)
    dst = np.zeros(n + 64, dtype=np.uint8)
    copy(lib, dst, 7, n, src, 0, n, n, 1)
    assert np.array_equal(dst[7:7 + n], src)


def test_pitched_rows(lib):
    rng = np.random.default_rng(2)
    for rows, width, spitch, dpitch in ((512, 65536, 65536, 4 << 20), (4096, 8192, 8192, 65536), (300, 1000, 1024, 1536), (7, 100000, 100008, 100000)):
        src = rng.integers(0, 256, rows * spitch, dtype=np.uint8)
        dst = np.full(rows * dpitch, 0xCD, dtype=np.uint8)
        copy(lib, dst, 0, dpitch, src, 0, spitch, width, rows)
        d2, s2 = dst.reshape(rows, dpitch), src.reshape(rows, spitch)
        assert np.array_equal(d2[:, :width], s2[:, :width])
        assert (d2[:, width:] == 0xCD).all(), "wrote into the padding between rows"


def test_concurrent_callers_share_the_pool(lib):
    rng = np.random.default_rng(3)
    n = 9 << 20
    srcs = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(4)]
    dsts = [np.zeros(n, dtype=np.uint8) for _ in range(4)]

    def work(i):
        for _ in range(5):
            copy(lib, dsts[i], 0, n, srcs[i], 0, n, n, 1)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(4):
        assert np.array_equal(dsts[i], srcs[i])


def test_bad_arguments(lib):
    a = np.zeros(16, dtype=np.uint8)
    assert lib.sstc_exp_host_copy(None, 16, ctypes.c_void_p(a.ctypes.data), 16, 16, 1) != 0
    assert lib.sstc_exp_host_copy(ctypes.c_void_p(a.ctypes.data), 8, ctypes.c_void_p(a.ctypes.data), 16, 16, 1) != 0
    assert b"Invalid arguments" in lib.sstc_last_error()
