"""The checks both sparse test files run: against tests/emu (host executor, CPU) and against the CUDA provider (GPU).
`k` has ArrayKernel's insertSparse / sliceSparse; results are (values, indices, indirectionOffsets, indirections)."""
import os

import numpy as np
import pytest

from sparse_cases import (dense_slice_expectation, dim_offsets, empty_state, insert_cases, row_major, run_insert,
                          run_slice, slice_cases)

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_golden.npz"))
FIELDS = ("values", "indices", "offsets", "indirections")


def same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def check_insert_cases(k, ref=None):
    for n, c in enumerate(insert_cases()):
        got = run_insert(k, c)
        for name, a in zip(FIELDS + ("newI",), got):
            assert same(a, GOLDEN["insert%d/%s" % (n, name)]), (n, c["dims"], name)
        if ref is not None:
            for name, a, b in zip(FIELDS + ("newI",), got, run_insert(ref, c)):
                assert same(a, b), (n, c["dims"], name)
    return n + 1


def check_slice_cases(k, ref=None):
    for n, c in enumerate(slice_cases()):
        got = run_slice(k, c)
        for name, a in zip(FIELDS, got):
            assert same(a, GOLDEN["slice%d/%s" % (n, name)]), (n, c["sdims"], c["ddims"], name)
        if ref is not None:
            for name, a, b in zip(FIELDS, got, run_slice(ref, c)):
                assert same(a, b), (n, name)
    return n + 1


def check_dense_property(k):
    """SparseArrayTest.java:198-233: slicing sparse arrays gives what slicing the dense arrays gives."""
    checked = 0
    for c in slice_cases():
        per_dim_dst = [[int(c["slices"][i + 1]) for i in range(0, c["slices"].size, 3) if c["slices"][i + 2] == d]
                       for d in range(len(c["sdims"]))]
        if any(len(set(p)) != len(p) for p in per_dim_dst):
            continue  # repeated destination indices: dense assignment order is not what the sparse rule keeps
        values, indices, _, _ = run_slice(k, c)
        want = dense_slice_expectation(c)
        dense = np.full(want.size, None, dtype=object)
        dense[indices] = list(values)
        assert all((a is None and b is None) or (a is not None and b is not None and a == b) for a, b in zip(dense, want))
        checked += 1
    assert checked >= 10


def metadata(dims, indices):
    """indirectionOffsets / indirections as SparseOps.cpp:198-248 defines them, with numpy."""
    strides = row_major(dims)
    count = indices.size
    offsets, indirections = [], []
    acc = indices.astype(np.int64)
    for d, (size, stride) in enumerate(zip(dims, strides)):
        logical = acc // stride
        acc = acc % stride
        indirections.append(np.argsort(logical, kind="stable").astype(np.int32))
        offsets.append(np.concatenate([np.searchsorted(np.sort(logical), np.arange(size)), [count]]).astype(np.int32))
    return np.concatenate(offsets), (np.concatenate(indirections) if count else np.zeros(0, dtype=np.int32))


def check_large(k, dims=(128, 128, 96), n_old=300_000, n_new=500_000):
    """Sizes past one scan chunk / one sort tile, against a numpy model of the semantics."""
    rng = np.random.default_rng(99)
    dims = list(dims)
    total = int(np.prod(dims))
    old_phys = np.sort(rng.choice(total, n_old, replace=False)).astype(np.int32)
    new_phys = rng.choice(total, n_new, replace=False).astype(np.int32)
    old_vals, new_vals = rng.uniform(-1, 1, n_old), rng.uniform(-1, 1, n_new)
    old = k.insertSparse(*empty_state(dims, np.float64)[:1], dims, row_major(dims), dim_offsets(dims),
                         np.zeros(0, dtype=np.int32), old_vals.copy(), old_phys.copy())
    assert np.array_equal(old[1], old_phys) and np.array_equal(old[0], old_vals)
    new_i = new_phys.copy()
    got = k.insertSparse(old[0], dims, row_major(dims), dim_offsets(dims), old[1], new_vals.copy(), new_i)
    assert np.array_equal(new_i, np.sort(new_phys))
    dense = np.full(total, np.nan)
    dense[old_phys] = old_vals
    dense[new_phys] = new_vals  # new wins
    want_i = np.flatnonzero(~np.isnan(dense)).astype(np.int32)
    assert np.array_equal(got[1], want_i)
    assert np.array_equal(got[0], dense[want_i])
    want_off, want_ind = metadata(dims, want_i)
    assert np.array_equal(got[2], want_off)
    assert np.array_equal(got[3], want_ind)
    # slice the whole of it into an empty array of the same shape with every dimension reversed
    slices = np.array([v for d, size in enumerate(dims) for j in range(size) for v in (j, size - 1 - j, d)], dtype=np.int32)
    e = empty_state(dims, np.float64)
    out = k.sliceSparse(slices, got[0], dims, row_major(dims), dim_offsets(dims), got[1], got[2], got[3], e[0], dims,
                        row_major(dims), dim_offsets(dims), e[1], e[2], e[3])
    rev = dense.reshape(dims)[::-1, ::-1, ::-1].ravel()
    rev_i = np.flatnonzero(~np.isnan(rev)).astype(np.int32)
    assert np.array_equal(out[1], rev_i) and np.array_equal(out[0], rev[rev_i])
    want_off, want_ind = metadata(dims, rev_i)
    assert np.array_equal(out[2], want_off) and np.array_equal(out[3], want_ind)


def check_errors(k, err):
    dims = [4, 5]
    S, Do = row_major(dims), dim_offsets(dims)
    e = empty_state(dims, np.float64)
    vals = np.array([1.0, 2.0, 3.0])

    def ins(D=dims, St=S, Dof=Do, new_i=(3, 7, 11), oldV=e[0], oldI=e[1], newV=vals):
        return k.insertSparse(oldV, D, St, Dof, oldI, newV.copy(), np.array(new_i, dtype=np.int32))

    with pytest.raises(err, match="Duplicate values are not allowed"):  # SparseOpsInsert.cpp:130-135
        ins(new_i=(3, 7, 3))
    with pytest.raises(err, match="Invalid physical index"):  # SparseOps.cpp:206-208
        ins(new_i=(3, 7, 20))
    with pytest.raises(err, match="Invalid physical index"):
        ins(new_i=(-1, 7, 11))
    with pytest.raises(err, match="Invalid dimensions and/or strides"):  # MappingOps.cpp:73-91
        ins(St=[4, 1])
    with pytest.raises(err, match="Invalid arguments"):  # SparseOpsInsert.cpp:103-108
        ins(Dof=[0, 5, 10])
    state = ins()
    assert np.array_equal(state[1], [3, 7, 11]) and np.array_equal(state[0], vals)

    def sl(slices, src=state, dst=e, sD=dims, dD=dims):
        return k.sliceSparse(np.array(slices, dtype=np.int32), src[0], sD, row_major(sD), dim_offsets(sD), src[1], src[2],
                             src[3], dst[0], dD, row_major(dD), dim_offsets(dD), dst[1], dst[2], dst[3])

    with pytest.raises(err, match="Invalid dimension"):  # SparseOpsSlice.cpp:146-160
        sl([0, 0, 2])
    with pytest.raises(err, match="Invalid index"):
        sl([4, 0, 0])
    with pytest.raises(err, match="Invalid index"):
        sl([0, 5, 1])
    with pytest.raises(err, match="Invalid arguments"):
        sl([0, 0])
    broken = (state[0], state[1], state[2].copy(), state[3])
    broken[2][1] = 99  # an indirection range past the array (SparseOpsSlice.cpp:482-497)
    with pytest.raises(err, match="Invalid arguments"):
        sl([0, 0, 0, 0, 0, 1], src=broken)
    # no slices at all: the destination comes back as it is
    out = sl([], dst=state)
    for a, b in zip(out, state):
        assert np.array_equal(a, b)
    # int values
    ei = empty_state(dims, np.int32)
    si = k.insertSparse(ei[0], dims, S, Do, ei[1], np.array([5, 6], dtype=np.int32), np.array([19, 0], dtype=np.int32))
    assert si[0].dtype == np.int32 and np.array_equal(si[0], [6, 5]) and np.array_equal(si[1], [0, 19])
