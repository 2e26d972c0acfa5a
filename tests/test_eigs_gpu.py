"""ArrayKernel.eigs through the C ABI against the reference's own native
LinearAlgebraOps::eigs (LinearAlgebraOpsEigs.cpp:26-615). Bar: BIT-IDENTICAL eigenvalues and eigenvectors -- every
matrix element is updated in the reference's operation order -- plus the reference test's own criterion
(MatrixTest.java:208-221: sum |A V - V D| < 1e-8 at size 128)."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from eigs_cases import cases, non_converging, residual

pytestmark = pytest.mark.gpu
ref = oracle.ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref/libsst_ref.so missing")
GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eigs_golden.npz"))


@pytest.fixture(scope="module")
def k():
    from shared_b200.array_kernel import CudaArrayKernel
    return CudaArrayKernel()


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64))


def test_eigs_bit_identical_to_reference_fixture(k):
    for name, a in cases():
        n = a.shape[0]
        src = a.ravel().copy()
        vec, val = np.full(n * n, 7.0), np.full(2 * n, 7.0)
        k.eigs(src, vec, val, n)
        assert np.array_equal(src, a.ravel()), name  # srcV is read-only (pinned READ_ONLY in the reference)
        assert same_bits(val, GOLDEN[name + "/val"]), name
        assert same_bits(vec, GOLDEN[name + "/vec"]), name


@needs_ref
def test_eigs_matches_live_reference_and_its_own_test(k):
    rng = np.random.default_rng(41)
    for n, trials in ((128, 4), (150, 1), (160, 1), (257, 1)):  # 128: MatrixTest.java:210; >= 159: working matrix in HBM
        for _ in range(trials):
            a = rng.uniform(0, 1, (n, n))  # RealArray(size, size).uRnd(1.0)
            vec, val = np.zeros(n * n), np.zeros(2 * n)
            want_vec, want_val = np.zeros(n * n), np.zeros(2 * n)
            k.eigs(a.ravel().copy(), vec, val, n)
            ref.eigs(a.ravel().copy(), want_vec, want_val, n)
            assert same_bits(val, want_val), n
            assert same_bits(vec, want_vec), n
            assert residual(a, vec, val) < 1e-8, (n, residual(a, vec, val))
            got = np.sort_complex(val[0::2] + 1j * val[1::2])
            assert np.allclose(got, np.sort_complex(np.linalg.eigvals(a)), rtol=0, atol=1e-9 * n), n


def test_eigs_errors_and_empty(k):
    import shared_b200 as s
    with pytest.raises(s.SstCudaError, match="Invalid arguments"):
        k.eigs(np.zeros(9), np.zeros(9), np.zeros(5), 3)
    with pytest.raises(s.SstCudaError, match="Invalid arguments"):
        k.eigs(np.zeros(8), np.zeros(9), np.zeros(6), 3)
    k.eigs(np.zeros(0), np.zeros(0), np.zeros(0), 0)
    # where the reference's uncapped QR loop never returns, the kernel gives up with an error instead of hanging
    for name, a in non_converging():
        n = a.shape[0]
        with pytest.raises(s.SstCudaError, match="did not converge"):
            k.eigs(a.ravel().copy(), np.zeros(n * n), np.zeros(2 * n), n)
    # ... and the library is still usable afterwards
    a = dict(cases())["uniform8"]
    vec, val = np.zeros(64), np.zeros(16)
    k.eigs(a.ravel().copy(), vec, val, 8)
    assert same_bits(vec, GOLDEN["uniform8/vec"])


def test_eigs_device_tier_on_a_stream(k):
    import torch

    import shared_b200
    lib = shared_b200.lib()
    a = dict(cases())["uniform33"]
    n = a.shape[0]
    dev = torch.device("cuda:0")
    src = torch.from_numpy(a.ravel().copy()).to(dev)
    vec = torch.zeros(n * n, dtype=torch.float64, device=dev)
    val = torch.zeros(2 * n, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    rc = lib.sstc_eigs_dev(ctypes.c_void_p(src.data_ptr()), n * n, ctypes.c_void_p(vec.data_ptr()), n * n,
                           ctypes.c_void_p(val.data_ptr()), 2 * n, n, ctypes.c_void_p(stream.cuda_stream))
    assert rc == 0, lib.sstc_last_error()
    stream.synchronize()
    assert same_bits(vec.cpu().numpy(), GOLDEN["uniform33/vec"])
    assert same_bits(val.cpu().numpy(), GOLDEN["uniform33/val"])
