"""CPU check of the sparse-array pipeline (shared_b200/csrc/sparse_core.cuh -- the orchestration and per-element code
array_sparse.cu runs on the GPU) under the serial host executor of tests/emu, bit-exact against the reference's own
SparseOps (golden fixture from oracle/_ref, and the live reference build where present); and of the three phases of the
executor's scan on 256 host threads per CTA under ThreadSanitizer."""
import os
import subprocess

import pytest

import emu_sparse
import oracle
import sparse_checks as sc


@pytest.fixture(scope="module")
def build_dir():
    try:
        path, log = emu_sparse.build()
    except RuntimeError as e:
        pytest.skip(str(e))
    if not os.path.exists(os.path.join(path, "libsparse_emu.so")):
        pytest.fail("emulation build failed:\n" + log)
    return path


@pytest.fixture(scope="module")
def k(build_dir):
    return emu_sparse.EmuSparseKernel(os.path.join(build_dir, "libsparse_emu.so"))


def test_insert_bit_exact(k):
    assert sc.check_insert_cases(k, oracle.ref()) >= 20


def test_slice_bit_exact(k):
    assert sc.check_slice_cases(k, oracle.ref()) >= 20


def test_slice_matches_dense_slicing(k):
    sc.check_dense_property(k)


def test_large_against_numpy_model(k):
    sc.check_large(k, dims=(48, 40, 32), n_old=20_000, n_new=30_000)


def test_errors(k):
    sc.check_errors(k, emu_sparse.EmuError)


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref/libsst_ref.so missing")
def test_fuzz_against_live_reference(k):
    """Random ranks 1-4, dimensions 1-6 (size-1 dimensions, empty arrays), slice triples in random order with repeated
    source and destination indices, dimensions without slices; insert with overlapping and disjoint new indices."""
    import numpy as np

    from sparse_cases import run_insert, run_slice
    rng = np.random.default_rng(20261025)
    ref = oracle.ref()
    for _ in range(1500):
        nd = int(rng.integers(1, 5))
        sdims = [int(rng.integers(1, 7)) for _ in range(nd)]
        ddims = [int(rng.integers(1, 7)) for _ in range(nd)]
        dtype = [np.float64, np.int32][int(rng.integers(0, 2))]
        stot, dtot = int(np.prod(sdims)), int(np.prod(ddims))
        sphys = np.sort(rng.choice(stot, size=int(rng.integers(0, stot + 1)), replace=False)).astype(np.int32)
        dphys = np.sort(rng.choice(dtot, size=int(rng.integers(0, dtot + 1)), replace=False)).astype(np.int32)
        mk = (lambda n: rng.uniform(-1, 1, n)) if dtype == np.float64 else (lambda n: rng.integers(-99, 99, n))
        triples = []
        for d in range(nd):
            count = 0 if rng.random() < 0.15 else int(rng.integers(0, 2 * max(sdims[d], ddims[d]) + 1))
            triples += [[int(rng.integers(0, sdims[d])), int(rng.integers(0, ddims[d])), d] for _ in range(count)]
        triples = np.array(triples, dtype=np.int32).reshape(-1, 3)
        rng.shuffle(triples)
        c = dict(sdims=sdims, ddims=ddims, sphys=sphys, svals=mk(sphys.size).astype(dtype), dphys=dphys,
                 dvals=mk(dphys.size).astype(dtype), slices=triples.ravel().copy())
        for a, b in zip(run_slice(k, c), run_slice(ref, c)):
            assert sc.same(a, b), c
        new_phys = rng.permutation(stot)[:int(rng.integers(0, stot + 1))].astype(np.int32)
        ci = dict(dims=sdims, old_phys=sphys, old_vals=c["svals"], new_phys=new_phys, new_vals=mk(new_phys.size).astype(dtype))
        for a, b in zip(run_insert(k, ci), run_insert(ref, ci)):
            assert sc.same(a, b), ci


def test_scan_phases_on_emulated_ctas(build_dir):
    fast, tsan = os.path.join(build_dir, "scan_emu_fast"), os.path.join(build_dir, "scan_emu_tsan")
    for n in (0, 1, 7, 2047, 2048, 2049, 4096, 100_000, 1_500_000):
        assert subprocess.run([fast, str(n), "11"], capture_output=True, timeout=300).returncode == 0, n
    if not os.path.exists(tsan):
        pytest.skip("ThreadSanitizer runtime not installed")
    for n in (5, 2049, 9000):
        res = subprocess.run([tsan, str(n), "3"], capture_output=True, text=True, timeout=600)
        assert "ThreadSanitizer" not in res.stderr, res.stderr[:3000]
        assert res.returncode == 0, n
