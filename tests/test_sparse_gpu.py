"""ArrayKernel.insertSparse / sliceSparse through the C ABI, bit-exact against
the reference's own native SparseOps (SparseOpsInsert.cpp, SparseOpsSlice.cpp, SparseOps.cpp): golden fixture, the live
reference build, the reference test's property (sparse slicing == dense slicing, SparseArrayTest.java:198-233), a numpy
model at sizes past one scan chunk / sort tile, and the reference's error texts."""
import numpy as np
import pytest

import oracle
import sparse_checks as sc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def k():
    from shared_b200.array_kernel import CudaArrayKernel
    return CudaArrayKernel()


def test_insert_bit_exact(k):
    assert sc.check_insert_cases(k, oracle.ref()) >= 20


def test_slice_bit_exact(k):
    assert sc.check_slice_cases(k, oracle.ref()) >= 20


def test_slice_matches_dense_slicing(k):
    sc.check_dense_property(k)


def test_large_against_numpy_model(k):
    sc.check_large(k)


def test_errors(k):
    import shared_b200 as s
    sc.check_errors(k, s.SstCudaError)
    with pytest.raises(s.SstCudaError, match="Invalid array types"):
        k.insertSparse(np.zeros(0), [3], [1], [0, 4], np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int32),
                       np.zeros(1, dtype=np.int32))
