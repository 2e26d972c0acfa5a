"""The reference's structural-operation tests (RealArrayTest.java:58-720, IntegerArrayTest.java:52-222, ArrayFftTest's
fftShift) replayed through shared_b200.proto_array.ProtoArray -- the host-side mirror of ProtoArray's callers of
ArrayKernel.map / slice -- with the reference's own native kernel as the provider. What is pinned here, on a box without
a GPU, is the bounds / slices arithmetic of tile, transpose, shift, subarray, reverse, concat, reverseOrder and the slice
variants against the literal known-answer arrays of the reference's tests (all `Arrays.equals`: exact). The same class
over CudaArrayKernel issues the same kernel calls (checks shared through tests/proto_array_checks.py)."""
import numpy as np
import pytest

import oracle
import proto_array_checks as pc

ref = oracle.ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref/libsst_ref.so missing")


@pytest.mark.parametrize("fixture,dtype", [("RealArrayTest", np.float64), ("IntegerArrayTest", np.int32)])
def test_map(fixture, dtype):
    pc.check_map(ref, fixture, dtype)


def test_slice_variants():
    pc.check_slice_variants(ref)


def test_integer_slice():
    pc.check_integer_slice(ref)


def test_tile_transpose_shift_subarray():
    pc.check_tile_transpose_shift_subarray(ref)


def test_reshape_reverse_concat_reverse_order():
    pc.check_reshape_reverse_concat_reverse_order(ref)


def test_fft_shift_is_a_shift_of_the_complex_array(fft_kats):
    pc.check_fft_shift_is_a_shift_of_the_complex_array(ref, fft_kats)


def test_integer_find():
    pc.check_integer_find(ref)


def test_argument_errors():
    pc.check_argument_errors(ref)
