"""The reference's structural-operation KATs through ProtoArray over the CUDA
provider. Same checks as tests/test_proto_array.py, which runs them over the reference's native kernel on the CPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import proto_array_checks as pc  # noqa: E402
from conftest import golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def k():
    from shared_b200.array_kernel import CudaArrayKernel
    return CudaArrayKernel()


@pytest.mark.parametrize("fixture,dtype", [("RealArrayTest", np.float64), ("IntegerArrayTest", np.int32)])
def test_map(k, fixture, dtype):
    pc.check_map(k, fixture, dtype)


def test_structural_operations(k):
    pc.check_slice_variants(k)
    pc.check_integer_slice(k)
    pc.check_tile_transpose_shift_subarray(k)
    pc.check_reshape_reverse_concat_reverse_order(k)
    pc.check_fft_shift_is_a_shift_of_the_complex_array(k, golden("ArrayFftTest"))
    pc.check_integer_find(k)
    pc.check_argument_errors(k)
