"""Seeded inputs for the insertSparse / sliceSparse parity tests (GPU, host emulation, golden fixture).

A case is a dict of the exact arguments of ArrayKernel.insertSparse / sliceSparse (ArrayKernel.java:628-674). Sparse
states are built the way ProtoSparseArray does it: by inserting into the empty state of the given dimensions
(ProtoSparseArray.java:190-210), through whatever kernel `build` is handed (the reference for fixtures)."""
import numpy as np


def row_major(dims):
    s = [1] * len(dims)
    for i in range(len(dims) - 2, -1, -1):
        s[i] = s[i + 1] * dims[i + 1]
    return s


def dim_offsets(dims):
    out = [0]
    for d in dims:
        out.append(out[-1] + d + 1)
    return out


def empty_state(dims, dtype):
    nd = len(dims)
    return (np.zeros(0, dtype=dtype), np.zeros(0, dtype=np.int32), np.zeros(int(np.sum(dims)) + nd, dtype=np.int32),
            np.zeros(0, dtype=np.int32))


def build_state(kernel, dims, physical, values):
    v0, i0, _, _ = empty_state(dims, values.dtype)
    return kernel.insertSparse(v0, dims, row_major(dims), dim_offsets(dims), i0, values.copy(),
                               np.ascontiguousarray(physical, dtype=np.int32).copy())


def insert_cases():
    rng = np.random.default_rng(20261022)
    shapes = [([7], 3, 4), ([4, 5], 6, 9), ([3, 4, 5], 20, 25), ([16, 16, 16], 900, 1500), ([2, 3, 2, 3], 10, 30),
              ([5, 1, 6], 0, 7), ([6, 6], 12, 0), ([3, 3], 0, 0), ([40, 50], 2000, 1), ([9, 9, 9], 729, 729),
              ([64, 64], 1000, 3000)]
    cases = []
    for dims, n_old, n_new in shapes:
        for dtype in (np.float64, np.int32):
            total = int(np.prod(dims))
            old_phys = np.sort(rng.choice(total, size=min(n_old, total), replace=False)).astype(np.int32)
            new_phys = rng.choice(total, size=min(n_new, total), replace=False).astype(np.int32)  # unsorted, overlaps old
            mk = (lambda n: rng.uniform(-1, 1, n)) if dtype == np.float64 else (lambda n: rng.integers(-999, 999, n).astype(np.int32))
            cases.append(dict(dims=dims, old_phys=old_phys, old_vals=mk(old_phys.size).astype(dtype), new_phys=new_phys,
                              new_vals=mk(new_phys.size).astype(dtype)))
    return cases


def slice_cases():
    """(src dims, dst dims, fill fractions, slice style). Styles: 'perm' = distinct indices per dimension as
    SparseArrayTest.java:189-196 draws them; 'dup' = repeated source and destination indices (one source index feeding
    several destination indices and several source indices landing on one destination index); 'none' = a dimension
    without any slice (nothing selected); 'full' = every index of every dimension."""
    rng = np.random.default_rng(20261023)
    specs = [([8], [8], 0.6, 0.6, "perm"), ([6, 7], [7, 6], 0.5, 0.5, "perm"), ([16, 16, 16], [16, 16, 16], 0.3, 0.3, "perm"),
             ([5, 6], [6, 5], 0.7, 0.4, "dup"), ([4, 4, 4], [5, 5, 5], 0.5, 0.5, "dup"), ([6, 6], [6, 6], 0.5, 0.5, "none"),
             ([5, 4, 3], [5, 4, 3], 0.6, 0.6, "full"), ([7, 7], [7, 7], 0.0, 0.5, "perm"), ([7, 7], [7, 7], 0.5, 0.0, "perm"),
             ([3, 3], [3, 3], 0.0, 0.0, "perm"), ([12, 10, 8], [9, 11, 10], 0.4, 0.6, "dup"), ([30, 40], [35, 45], 0.5, 0.5, "perm"),
             ([2, 3, 2, 3], [3, 2, 3, 2], 0.8, 0.8, "dup")]
    cases = []
    for sdims, ddims, sfill, dfill, style in specs:
        for dtype in (np.float64, np.int32):
            nd = len(sdims)
            stot, dtot = int(np.prod(sdims)), int(np.prod(ddims))
            sphys = np.sort(rng.choice(stot, size=int(sfill * stot), replace=False)).astype(np.int32)
            dphys = np.sort(rng.choice(dtot, size=int(dfill * dtot), replace=False)).astype(np.int32)
            mk = (lambda n: rng.uniform(-1, 1, n)) if dtype == np.float64 else (lambda n: rng.integers(-999, 999, n).astype(np.int32))
            slices = []
            for d in range(nd):
                if style == "none" and d == nd - 1:
                    continue
                lim = min(sdims[d], ddims[d])
                if style == "full":
                    si, di = np.arange(lim), np.arange(lim)
                elif style == "dup":
                    k = int(rng.integers(1, 2 * lim + 1))
                    si, di = rng.integers(0, sdims[d], k), rng.integers(0, ddims[d], k)
                else:
                    k = int(rng.integers(0, lim))
                    si, di = rng.permutation(sdims[d])[:k], rng.permutation(ddims[d])[:k]
                for a, b in zip(si, di):
                    slices += [int(a), int(b), d]
            cases.append(dict(sdims=sdims, ddims=ddims, sphys=sphys, svals=mk(sphys.size).astype(dtype), dphys=dphys,
                              dvals=mk(dphys.size).astype(dtype), slices=np.array(slices, dtype=np.int32)))
    return cases


def run_insert(kernel, c):
    """-> (values, indices, indirectionOffsets, indirections, newI as left by the call)"""
    dims = c["dims"]
    old = build_state(kernel, dims, c["old_phys"], c["old_vals"])
    new_i = c["new_phys"].copy()
    out = kernel.insertSparse(old[0], dims, row_major(dims), dim_offsets(dims), old[1], c["new_vals"].copy(), new_i)
    return tuple(out) + (new_i,)


def run_slice(kernel, c):
    src = build_state(kernel, c["sdims"], c["sphys"], c["svals"])
    dst = build_state(kernel, c["ddims"], c["dphys"], c["dvals"])
    return kernel.sliceSparse(c["slices"], src[0], c["sdims"], row_major(c["sdims"]), dim_offsets(c["sdims"]), src[1],
                              src[2], src[3], dst[0], c["ddims"], row_major(c["ddims"]), dim_offsets(c["ddims"]), dst[1],
                              dst[2], dst[3])


def dense_slice_expectation(c):
    """What the sparse result must look like as a dense array: SparseArrayTest.java:198-233 compares sparse slicing
    with dense slicing. Only meaningful for slice specifications without repeated destination indices."""
    sdims, ddims = c["sdims"], c["ddims"]
    nd = len(sdims)
    NONE = object()
    src = np.full(int(np.prod(sdims)), None, dtype=object)
    dst = np.full(int(np.prod(ddims)), None, dtype=object)
    src[c["sphys"]] = list(c["svals"])
    dst[c["dphys"]] = list(c["dvals"])
    src, dst = src.reshape(sdims), dst.reshape(ddims)
    per_dim = [[(int(c["slices"][i]), int(c["slices"][i + 1])) for i in range(0, c["slices"].size, 3)
                if c["slices"][i + 2] == d] for d in range(nd)]
    if any(len(p) == 0 for p in per_dim):
        return dst.ravel()
    out = dst.copy()
    import itertools
    for combo in itertools.product(*per_dim):
        out[tuple(p[1] for p in combo)] = NONE
    for combo in itertools.product(*per_dim):
        s = src[tuple(p[0] for p in combo)]
        if s is not None:
            out[tuple(p[1] for p in combo)] = s
    out = out.ravel()
    for i in range(out.size):
        if out[i] is NONE:
            out[i] = None
    return out
