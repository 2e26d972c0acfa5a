"""The host tier of the FftService provider (sstc_fft, shared_b200/csrc/fft_host.cu): the chunked H2D / inner axes /
outermost axis / D2H pipeline, pageable arrays staged through the pinned ring, pinned arrays read by the DMA engines
directly, and concurrent callers on separate slots. Every variant must give the SAME BITS as the unpipelined schedule
(same kernels, same arithmetic) and match the CPU oracle within 1e-12 * log2 N (BASELINE.json north_star)."""
import ctypes
import math
import threading

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def rel_tol(n):
    return 1e-12 * max(1.0, math.log2(max(2, n)))


def rel_l2(got, want):
    return np.linalg.norm(got - want) / np.linalg.norm(want)


@pytest.fixture(scope="module")
def s():
    import shared_b200
    shared_b200.init(0)
    return shared_b200


def option(s, name, value):
    from shared_b200 import _lib
    _lib.check(_lib.lib().sstc_host_option(name.encode(), int(value)))


def pinned(n):
    from shared_b200 import _lib
    p = ctypes.c_void_p()
    _lib.check(_lib.lib().sstc_host_alloc(ctypes.byref(p), n * 8))
    return np.frombuffer((ctypes.c_double * n).from_address(p.value), dtype=np.float64), p


def transform(s, kind, dims, x, out=None):
    from shared_b200 import _lib
    _, out_len = s.transform_lengths(kind, dims)
    if out is None:
        out = np.full(out_len, np.nan)
    _lib.check(_lib.lib().sstc_fft(kind, len(dims), _lib.i32_array(dims), x.ctypes.data, x.size, out.ctypes.data, out.size))
    return out


# sizes above the 8 MiB pipelining threshold: pow2 and odd outer axes, a 2-D case with many thin planes, rank 4,
# a chunk size that does not divide the planes
CASES = [
    ("fft", [64, 128, 128]), ("ifft", [64, 128, 128]), ("fft", [6, 256, 512]), ("fft", [2048, 512]), ("ifft", [27, 40, 512]),
    ("fft", [5, 4, 64, 1024]), ("rfft", [64, 128, 256]), ("rfft", [12, 100, 1000]), ("rifft", [64, 128, 256]),
    ("fft", [3000, 700]),
]


@pytest.mark.parametrize("name,dims", CASES, ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else v)
def test_pipeline_is_bit_identical_and_matches_oracle(s, name, dims):
    kind = {"fft": s.FORWARD, "ifft": s.BACKWARD, "rfft": s.R2C, "rifft": s.C2R}[name]
    in_len, _ = s.transform_lengths(kind, dims)
    n = int(np.prod(dims))
    rng = np.random.default_rng(11)
    if name == "rifft":
        x = oracle.rfft(dims, rng.uniform(-0.5, 0.5, n))  # a Hermitian-consistent half-spectrum
    else:
        x = rng.uniform(-0.5, 0.5, in_len)
    keep = x.copy()
    try:
        option(s, "pipeline", 0)
        plain = transform(s, kind, dims, x)
        option(s, "pipeline", 1)
        option(s, "chunk_mib", 3)   # many chunks and groups, ragged against the plane size
        piped_small = transform(s, kind, dims, x)
        option(s, "chunk_mib", 32)
        piped = transform(s, kind, dims, x)
    finally:
        option(s, "pipeline", 1)
        option(s, "chunk_mib", 32)
    assert np.array_equal(x, keep), "input must be preserved"
    assert np.array_equal(piped, plain) and np.array_equal(piped_small, plain)
    want = getattr(oracle, name)(dims, x)
    assert rel_l2(piped, want) < rel_tol(n)
    # pinned arrays: the same bits again, without the staging ring
    hx, px = pinned(in_len)
    ho, po = pinned(plain.size)
    try:
        hx[:] = x
        ho[:] = np.nan
        transform(s, kind, dims, hx, ho)
        assert np.array_equal(ho, plain)
    finally:
        from shared_b200 import _lib
        _lib.lib().sstc_host_free(px)
        _lib.lib().sstc_host_free(po)


def test_concurrent_callers_share_the_bus(s):
    """Four threads, two slots: every result equals the one computed alone (slots are handed on, never shared)."""
    dims = [64, 128, 128]
    n = int(np.prod(dims))
    rng = np.random.default_rng(12)
    xs = [rng.uniform(-0.5, 0.5, 2 * n) for _ in range(4)]
    alone = [transform(s, s.FORWARD, dims, x) for x in xs]
    got = [None] * 4
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = transform(s, s.FORWARD, dims, xs[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(4):
        assert np.array_equal(got[i], alone[i])


def test_256_cube_against_oracle(s):
    """The largest cube the CPU oracle finishes in seconds (~7 s): every kernel of the 512^3 path at half the length,
    through the pipelined host tier."""
    dims = [256, 256, 256]
    n = 256 ** 3
    x = np.random.default_rng(13).uniform(-0.5, 0.5, 2 * n)
    got = transform(s, s.FORWARD, dims, x)
    want = oracle.fft(dims, x)
    assert rel_l2(got, want) < rel_tol(n)
    back = transform(s, s.BACKWARD, dims, got)
    assert rel_l2(back, x) < rel_tol(n)


def test_errors_leave_the_tier_usable(s):
    from shared_b200 import _lib
    x = np.zeros(16)
    out = np.zeros(16)
    rc = _lib.lib().sstc_fft(s.FORWARD, 1, _lib.i32_array([7]), x.ctypes.data, x.size, out.ctypes.data, out.size)
    assert rc != 0 and b"expected sizes" in _lib.lib().sstc_last_error()
    rc = _lib.lib().sstc_host_option(b"no_such_option", 1)
    assert rc != 0 and b"Unknown option" in _lib.lib().sstc_last_error()
    x = np.arange(16, dtype=np.float64)
    got = transform(s, s.FORWARD, [8], x)
    assert rel_l2(got, oracle.fft([8], x)) < 1e-14
