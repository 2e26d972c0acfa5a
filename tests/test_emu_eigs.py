"""CPU check of the eigs kernel bodies (shared_b200/csrc/eigs_core.cuh -- the code array_eigs.cu launches): the same
source runs on host threads standing in for the CTA's threads (tests/emu), once under ThreadSanitizer, where a barrier
missing between two phases is a reported data race, and its output must be bit-identical to the reference's own
LinearAlgebraOps::eigs (golden fixture generated from oracle/_ref by tests/golden/gen_eigs_golden.py, and the live
reference build where present)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from eigs_cases import cases, non_converging, residual

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
GOLDEN = np.load(os.path.join(HERE, "golden", "eigs_golden.npz"))


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None or shutil.which("make") is None:
        pytest.skip("no g++ / make")
    res = subprocess.run(["make", "-s", "-k", "-C", EMU, "all"], capture_output=True, text=True)
    fast = os.path.join(EMU, "_build", "eigs_emu_fast")
    tsan = os.path.join(EMU, "_build", "eigs_emu_tsan")
    if not os.path.exists(fast):
        pytest.fail("emulation build failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
    return {"fast": fast, "tsan": tsan if os.path.exists(tsan) else None}


def run(exe, a, nthreads, tmp_path, caps=()):
    """caps = (log_cap, chunk_cap): small values force the log-full path and the chunked staging of the rotation log."""
    n = a.shape[0]
    fin, fout = str(tmp_path / "in.f64"), str(tmp_path / "out.f64")
    a.tofile(fin)
    res = subprocess.run([exe, str(n), str(nthreads), fin, fout] + [str(c) for c in caps], capture_output=True, text=True,
                         timeout=300)
    out = np.fromfile(fout) if res.returncode == 0 else np.zeros(n * n + 2 * n)
    return res, out[:n * n], out[n * n:]


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64))


def test_emulation_bit_identical_to_reference_fixture(emu, tmp_path):
    for name, a in cases():
        for nthreads in (1, 7):
            res, vec, val = run(emu["fast"], a, nthreads, tmp_path)
            assert res.returncode == 0, (name, res.stderr[-500:])
            assert same_bits(val, GOLDEN[name + "/val"]), (name, nthreads)
            assert same_bits(vec, GOLDEN[name + "/vec"]), (name, nthreads)
        if not name.startswith(("repeated", "zeros")):  # defective / null input: vectors are not a full basis
            assert residual(a, vec, val) < 1e-8 * max(1.0, np.abs(a).sum()), name  # MatrixTest.java:216-219


def test_emulation_has_no_data_race(emu, tmp_path):
    if emu["tsan"] is None:
        pytest.skip("ThreadSanitizer runtime not installed")
    for name, a in cases(sizes=(1, 2, 3, 5, 8, 16, 33)):
        for caps in ((), (7, 3)):
            res, vec, val = run(emu["tsan"], a, 6, tmp_path, caps)
            assert "ThreadSanitizer" not in res.stderr, (name, caps, res.stderr[:3000])
            assert res.returncode == 0, (name, caps)
            assert same_bits(vec, GOLDEN[name + "/vec"]) and same_bits(val, GOLDEN[name + "/val"]), (name, caps)


def test_rotation_log_overflow_and_chunking(emu, tmp_path):
    """The QR steps are recorded and applied to V afterwards; a full log is applied on the way. Any capacity must give
    the same bits."""
    for name, a in cases(sizes=(16, 64)):
        for caps in ((1, 1), (5, 2), (64, 64), (1000, 7)):
            res, vec, val = run(emu["fast"], a, 5, tmp_path, caps)
            assert res.returncode == 0, (name, caps)
            assert same_bits(vec, GOLDEN[name + "/vec"]) and same_bits(val, GOLDEN[name + "/val"]), (name, caps)


def test_emulation_gives_up_where_the_reference_spins(emu, tmp_path):
    for name, a in non_converging():
        res, _, _ = run(emu["fast"], a, 4, tmp_path)
        assert res.returncode == 9, name  # EIGS_FAILED -> "Eigenvalue iteration did not converge"


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libsst_ref.so missing")
def test_emulation_matches_live_reference_at_the_reference_tests_size(emu, tmp_path):
    rng = np.random.default_rng(7)
    ref = oracle.ref()
    for n, nthreads in ((128, 8), (160, 8)):  # MatrixTest.java:210 uses 128; 160 is past the shared-memory layout
        a = rng.uniform(0, 1, (n, n))
        want_vec, want_val = np.zeros(n * n), np.zeros(2 * n)
        ref.eigs(a.ravel().copy(), want_vec, want_val, n)
        res, vec, val = run(emu["fast"], a, nthreads, tmp_path)
        assert res.returncode == 0
        assert same_bits(vec, want_vec) and same_bits(val, want_val), n
        assert residual(a, vec, val) < 1e-8, n


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libsst_ref.so missing")
def test_emulation_fuzz_structured_matrices(emu, tmp_path):
    """Small matrices with exact zeros / ties / integer entries walk the rare branches of hqr2 (zero Householder columns,
    p == 0, s == 0, NaN-producing 2 x 2 zero blocks). The reference is only called where the emulation converged -- same
    bits, same path -- because the reference has no iteration limit."""
    rng = np.random.default_rng(20261024)
    ref = oracle.ref()
    compared = 0
    for _ in range(250):
        size = int(rng.integers(1, 13))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            a = rng.integers(-2, 3, (size, size)).astype(np.float64)
        elif kind == 1:
            a = rng.uniform(-1, 1, (size, size)) * (rng.random((size, size)) < 0.4)
        elif kind == 2:
            a = np.triu(rng.integers(-3, 4, (size, size)).astype(np.float64), -1)
        elif kind == 3:
            a = rng.standard_normal((size, size))
            a = a + a.T
        else:
            a = rng.integers(0, 2, (size, size)).astype(np.float64)
        a = np.ascontiguousarray(a)
        res, vec, val = run(emu["fast"], a, int(rng.integers(1, 9)), tmp_path)
        if res.returncode == 9:
            continue
        assert res.returncode == 0
        want_vec, want_val = np.zeros(size * size), np.zeros(2 * size)
        ref.eigs(a.ravel().copy(), want_vec, want_val, size)
        assert np.array_equal(vec, want_vec, equal_nan=True) and np.array_equal(val, want_val, equal_nan=True), a
        compared += 1
    assert compared >= 200
