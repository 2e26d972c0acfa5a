"""The reference's structural-operation tests (RealArrayTest.java:58-720, IntegerArrayTest.java:52-222, ArrayFftTest's
fftShift) as checks over ANY ArrayKernel provider `k`, run through shared_b200.proto_array.ProtoArray. tests/
test_proto_array.py hands in the reference's native kernel (CPU); the GPU replay over CudaArrayKernel is
tests/test_proto_array_gpu.py (over CudaArrayKernel, on a B200)."""
import numpy as np
import pytest

from conftest import golden
from shared_b200.proto_array import FAR, NEAR, ProtoArray


def arr(k, rec, dtype=np.float64):
    return ProtoArray(k, np.array(rec["values"], dtype=dtype), rec["dims"], rec["order"] or FAR)


def same(a, rec):
    return np.array_equal(a.values, np.array(rec["values"], dtype=a.values.dtype))


def check_map(k, fixture, dtype):
    a = arr(k, golden(fixture)["testMap"][0], dtype)  # RealArrayTest.java:64-93, IntegerArrayTest.java:58-86
    bounds = (1, 0, 2, 0, 1, 2, 1, 0, 3)
    expected = a.map(ProtoArray(k, None, [3, 3, 4], NEAR, dtype=dtype), *bounds)
    got = a.map(ProtoArray(k, None, [3, 3, 4], FAR, dtype=dtype), *bounds)
    assert np.array_equal(got.reverseOrder().values, expected.values) and got.values.any()


def check_slice_variants(k):
    r = golden("RealArrayTest")["testSlice"]  # RealArrayTest.java:100-229
    original = arr(k, r[0])
    a = original.splice(ProtoArray(k, None, [5, 5]), 1, 0, 0, 3, 2, 0, 5, 4, 0, 0, 0, 1, 2, 1, 1, 4, 3, 1, 6, 4, 1)
    assert same(a, r[1])
    a = original.slice_into([r[2]["values"], r[3]["values"]], ProtoArray(k, None, [5, 5]), [r[4]["values"], r[5]["values"]])
    assert same(a, r[1])
    a = arr(k, r[6]).slice_to(ProtoArray(k, None, [5, 5]), r[8]["values"], r[9]["values"])
    assert same(a, r[7])
    a = arr(k, r[10]).slice_fill(0.0, r[12]["values"], r[13]["values"])
    assert same(a, r[11])
    a = original.slice(r[14]["values"], r[15]["values"])
    assert a.dims == [3, 4] and same(a, r[16])


def check_integer_slice(k):
    r = golden("IntegerArrayTest")["testSlice"]  # IntegerArrayTest.java:94-178
    a = arr(k, r[0], np.int32).splice(ProtoArray(k, None, [2, 3, 4], NEAR, dtype=np.int32),
                                   1, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 1, 0, 2, 2, 1, 2, 2, 2, 2, 3, 3, 2)
    assert same(a.reverseOrder(), r[1])
    a = arr(k, r[2], np.int32).splice(ProtoArray(k, None, [3, 3], dtype=np.int32), 1, 0, 0, 2, 1, 0, 3, 2, 0, 0, 0, 1, 2, 1, 1,
                                   4, 2, 1)
    assert same(a, r[3])


def check_tile_transpose_shift_subarray(k):
    r = golden("RealArrayTest")
    t = r["testTile"]  # RealArrayTest.java:238-300
    assert same(arr(k, t[0]).tile(2, 2, 2), t[1]) and same(arr(k, t[2]).tile(1, 5), t[3])
    t = r["testTranspose"]  # :309-372
    assert same(arr(k, t[0]).transpose(1, 0), t[1])
    got = arr(k, t[2]).transpose(2, 0, 1)
    assert got.dims == [4, 5, 3] and same(got, t[3])
    t = r["testShift"]  # :390-443
    assert same(arr(k, t[0]).shift(-3, -1, 1), t[1])
    t = r["testSubarray"]  # :471-506
    got = arr(k, t[0]).subarray(1, 3, 1, 4, 1, 3)
    assert got.dims == [2, 3, 2] and same(got, t[1])


def check_reshape_reverse_concat_reverse_order(k):
    r = golden("RealArrayTest")
    t = r["testReshape"]  # RealArrayTest.java:520-575
    a = arr(k, t[0])
    assert same(a.mMul(a.reshape(6, 3)), t[1])
    assert same(arr(k, t[2]).reshape(9, 3).tile(1, 2), t[3])
    t = r["testReverse"]  # :587-630
    assert same(arr(k, t[0]).reverse(2, 0), t[1])
    t = r["testConcat"]  # :638-682
    a, b = arr(k, t[0]), arr(k, t[1])
    got = a.concat(1, b, a)
    assert got.dims == [2, 7, 2] and same(got, t[2])
    t = r["testReverseOrder"]  # :695-720
    got = arr(k, t[0]).reverseOrder().reverseOrder().reverseOrder()
    assert got.order == NEAR and same(got, t[1])


def check_fft_shift_is_a_shift_of_the_complex_array(k, fft_kats):
    """AbstractComplexArray.fftShift (AbstractComplexArray.java:85-97): shift by dims / 2 over every dimension but the
    trailing re / im pair. ArrayFftTest.java:232-284."""
    recs = fft_kats["testFftShift"]
    a = arr(k, recs[0])
    assert a.dims[-1] == 2
    shifts = [d // 2 for d in a.dims[:-1]] + [0]
    assert same(a.shift(*shifts), recs[1])
    back = [-(d // 2) for d in a.dims[:-1]] + [0]
    assert np.array_equal(a.shift(*shifts).shift(*back).values, a.values)


def check_integer_find(k):
    """IntegerArray.find (IntegerArray.java) -> ArrayKernel.find on the positions riOp wrote. IntegerArrayTest.java:185-217."""
    r = golden("IntegerArrayTest")["testFind"]
    a = arr(k, r[0], np.int32)
    n = 1
    for i in range(3):
        for j in range(4):
            got = k.find(a.values, a.dims, a.strides, [i, j, -1])
            assert np.array_equal(got, np.array(r[n]["values"], dtype=np.int32)), (i, j)
            n += 1


def check_argument_errors(k):
    a = ProtoArray(k, np.arange(6.0), [2, 3])
    with pytest.raises(ValueError, match="Invalid permutation"):
        a.transpose(1, 1)
    with pytest.raises(ValueError, match="Dimensionality mismatch"):
        a.tile(2)
    with pytest.raises(ValueError, match="Invalid subarray bounds"):
        a.subarray(0, 1)
    with pytest.raises(ValueError, match="Cardinality mismatch"):
        a.reshape(4, 2)
    with pytest.raises(ValueError, match="Source and destination cannot be the same"):
        a.map(a, 0, 0, 2, 0, 0, 3)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        a.concat(0, ProtoArray(k, np.arange(4.0), [2, 2]))
