"""Parity of the CUDA ArrayKernel provider (through the C ABI host tier) with
  (1) the reference's own known-answer tests (ArrayKernelTest / RealArrayTest / MatrixTest literals), and
  (2) the reference's own native kernel (oracle.ref() = /root/reference C++ built against a shim) on seeded inputs.
Bar: bit-exact for integer / index / data-movement results and for IEEE-exact maps; relative L2 <= 1e-12*log2(n)
for floating-point reductions, scans and products (BASELINE.json north_star)."""
import math

import numpy as np
import pytest

import oracle
from conftest import golden
from kernel_kats import OPCODES as OP, check_all

pytestmark = pytest.mark.gpu
ref = oracle.ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref/libsst_ref.so missing")


@pytest.fixture(scope="module")
def k():
    from shared_b200.array_kernel import CudaArrayKernel
    return CudaArrayKernel()


def strides(dims, order="FAR"):
    n = len(dims)
    s = [1] * n
    if order in (None, "FAR"):
        for i in range(n - 2, -1, -1):
            s[i] = s[i + 1] * dims[i + 1]
    else:  # NEAR = column-major (Array.java:240-287)
        for i in range(1, n):
            s[i] = s[i - 1] * dims[i - 1]
    return s


def rel(got, want):
    d = np.linalg.norm(want)
    return np.linalg.norm(got - want) / (d if d > 0 else 1.0)


def tol(n):
    return 1e-12 * max(1.0, math.log2(max(n, 2)))


# ---- (1) the reference's KATs ---------------------------------------------------------------------------

def test_call_level_kats(k):
    assert check_all(k) >= 50  # ArrayKernelTest.java:61-378


def _reduce(k, op, rec, exp, *opdims):
    src = np.array(rec["values"], dtype=np.float64)
    dst = np.full(len(exp["values"]), np.nan)
    k.rrOp(op, src, rec["dims"], strides(rec["dims"], rec["order"]), dst, exp["dims"],
           strides(exp["dims"], exp["order"]), *opdims)
    return dst, np.array(exp["values"], dtype=np.float64)


def test_kat_rrops(k):
    r = golden("RealArrayTest")["testRrOps"]  # RealArrayTest.java:724-938
    for rec, exp, op, dims in [(r[0], r[1], "RR_SUM", (1,)), (r[2], r[3], "RR_SUM", (1,)), (r[2], r[4], "RR_SUM", (1, 0)),
                               (r[2], r[5], "RR_SUM", (1, 2)), (r[2], r[6], "RR_SUM", (1, 0, 2)),
                               (r[7], r[8], "RR_PROD", (1,)), (r[9], r[10], "RR_MIN", (0,)), (r[11], r[12], "RR_MAX", (1,)),
                               (r[13], r[14], "RR_VAR", (0,))]:
        got, want = _reduce(k, OP[op], rec, exp, *dims)
        assert np.array_equal(got, want), (op, dims, got, want)  # the reference asserts exact equality here


def _index(k, op, rec, dim):
    src = np.array(rec["values"], dtype=np.float64)
    dst = np.full(src.size, -7, dtype=np.int32)
    k.riOp(op, src, rec["dims"], strides(rec["dims"], rec["order"]), dst, dim)
    return src, dst


def test_kat_riops(k):
    r = golden("RealArrayTest")["testRiOps"]  # RealArrayTest.java:944-1275
    for a, e, op, dim in [(r[0], r[1], "RI_MAX", 1), (r[2], r[3], "RI_ZERO", 2), (r[4], r[5], "RI_GZERO", 2),
                          (r[9], r[10], "RI_MAX", -1), (r[11], r[12], "RI_MIN", -1), (r[13], r[14], "RI_ZERO", -1),
                          (r[15], r[16], "RI_GZERO", -1)]:
        _, got = _index(k, OP[op], a, dim)
        assert np.array_equal(got, np.array(e["values"], dtype=np.int32)), (op, dim)
    for a, e, v, dim in [(r[6], r[7], r[8], 1), (r[17], r[18], r[19], -1)]:  # iSort: indices and sorted source
        src, got = _index(k, OP["RI_SORT"], a, dim)
        assert np.array_equal(got, np.array(e["values"], dtype=np.int32))
        assert np.array_equal(src, np.array(v["values"]))


def test_kat_rdops(k):
    r = golden("RealArrayTest")["testRdOps"]  # RealArrayTest.java:1281-1361
    for a, e, op, dims in [(r[0], r[1], "RD_SUM", (0, 2)), (r[2], r[3], "RD_PROD", (1,))]:
        src = np.array(a["values"], dtype=np.float64)
        dst = np.full(src.size, np.nan)
        k.rdOp(OP[op], src, a["dims"], strides(a["dims"], a["order"]), dst, *dims)
        assert np.array_equal(dst, np.array(e["values"]))


def test_kat_matrix(k):
    r = golden("MatrixTest")["testMMul"]  # MatrixTest.java:67-123
    a = np.array(r[0]["values"])
    at = np.ascontiguousarray(a.reshape(4, 3).T).ravel()
    dst = np.zeros(16)
    k.mul(a, at, 4, 4, dst, False)
    assert np.max(np.abs(dst - np.array(r[1]["values"]))) < 1e-8
    dst = np.zeros(18)
    k.mul(np.array(r[2]["values"]), np.array(r[3]["values"]), 3, 3, dst, True)
    assert np.max(np.abs(dst - np.array(r[4]["values"]))) < 1e-8
    d = golden("MatrixTest")["testMDiag"]  # MatrixTest.java:129-180
    dst = np.zeros(5)
    k.diag(np.array(d[0]["values"]), dst, 5, False)
    assert np.array_equal(dst, np.array(d[1]["values"]))
    dst = np.zeros(10)
    k.diag(np.array(d[2]["values"]), dst, 5, True)
    assert np.array_equal(dst, np.array(d[3]["values"]))


# ---- (2) seeded parity against the reference's native kernel --------------------------------------------

SHAPES = [([7], "FAR"), ([3, 4, 5], "FAR"), ([3, 4, 5], "NEAR"), ([6, 1, 5], "FAR"), ([2, 3, 4, 5, 6], "FAR"),
          ([2, 3, 4, 5, 6], "NEAR"), ([300, 70], "FAR"), ([70, 300], "NEAR"), ([5, 2100], "FAR"), ([2100, 5], "FAR"),
          ([17, 33, 9], "FAR"), ([1, 1, 1], "FAR"), ([4096], "FAR"),
          # many long lines off the fastest dimension: the strip scan / thread-per-line folds
          ([130, 5000], "FAR"), ([5000, 200], "NEAR"), ([70, 40, 130], "FAR"), ([5000, 300], "FAR")]


@needs_ref
@pytest.mark.parametrize("dims,order", SHAPES, ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else v)
def test_rrop_matches_reference(k, dims, order):
    rng = np.random.default_rng(11)
    n = int(np.prod(dims))
    st = strides(dims, order)
    src = rng.uniform(-1, 1, n)
    nd = len(dims)
    choices = [(d,) for d in range(nd)] + ([tuple(range(nd))] if nd > 1 else []) + ([(nd - 1, 0)] if nd > 2 else [])
    for op in ("RR_SUM", "RR_PROD", "RR_MAX", "RR_MIN", "RR_VAR"):
        for opdims in choices:
            ddims = [1 if i in opdims else dims[i] for i in range(nd)]
            dst_a, dst_b = np.zeros(int(np.prod(ddims))), np.zeros(int(np.prod(ddims)))
            x = src if op != "RR_PROD" else 1.0 + 0.01 * src
            k.rrOp(OP[op], x.copy(), dims, st, dst_a, ddims, strides(ddims, order), *opdims)
            ref.rrOp(OP[op], x.copy(), dims, st, dst_b, ddims, strides(ddims, order), *opdims)
            if op in ("RR_MAX", "RR_MIN"):
                assert np.array_equal(dst_a, dst_b), (op, opdims)
            else:
                assert rel(dst_a, dst_b) < tol(n), (op, opdims, rel(dst_a, dst_b))


@needs_ref
@pytest.mark.parametrize("dims,order", SHAPES, ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else v)
def test_riop_matches_reference(k, dims, order):
    rng = np.random.default_rng(12)
    n = int(np.prod(dims))
    st = strides(dims, order)
    src = rng.integers(-2, 3, n).astype(np.float64)  # small integers: ties and zeros occur
    for op in ("RI_MAX", "RI_MIN", "RI_ZERO", "RI_GZERO", "RI_LZERO"):
        for dim in list(range(len(dims))) + [-1]:
            a, b = np.full(n, -9, dtype=np.int32), np.full(n, -9, dtype=np.int32)
            k.riOp(OP[op], src.copy(), dims, st, a, dim)
            ref.riOp(OP[op], src.copy(), dims, st, b, dim)
            assert np.array_equal(a, b), (op, dim)
    # sort: distinct values -> identical to the reference; ties -> stable order (the Java provider's contract)
    distinct = rng.permutation(n).astype(np.float64) - n / 3.0
    for dim in list(range(len(dims))) + [-1]:
        sa, sb = distinct.copy(), distinct.copy()
        a, b = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        k.riOp(OP["RI_SORT"], sa, dims, st, a, dim)
        ref.riOp(OP["RI_SORT"], sb, dims, st, b, dim)
        assert np.array_equal(sa, sb) and np.array_equal(a, b), dim
    ties = src.copy()
    perm = np.zeros(n, dtype=np.int32)
    k.riOp(OP["RI_SORT"], ties, [n], [1], perm, 0)
    assert np.array_equal(perm, np.argsort(src, kind="stable").astype(np.int32))
    assert np.array_equal(ties, np.sort(src))


def test_rrop_max_min_nan_and_zero_semantics(k):
    """Math.max / Math.min (DimensionOps.java:96-123): a NaN anywhere along the line gives NaN; the folds start from
    -/+Double.MAX_VALUE, so a line of -infinity reduces to -MAX_VALUE under MAX."""
    rng = np.random.default_rng(22)
    for dims in ([6, 300], [300, 6], [5000, 70], [6, 4096]):
        n = int(np.prod(dims))
        src = rng.uniform(-1, 1, n)
        m = src.reshape(dims)
        m[1, 3] = np.nan
        m[2, :] = -np.inf
        m[3, :] = -0.0           # signed zeros: Math.max(-0.0, +0.0) = +0.0, Math.min = -0.0 (compared bit for bit below)
        m[3, 5] = 0.0
        m[4, :] = 0.0
        m[4, 1] = -0.0
        for dim in (0, 1):
            ddims = [1 if i == dim else dims[i] for i in range(2)]
            for op, red, start in (("RR_MAX", np.max, -np.finfo(np.float64).max), ("RR_MIN", np.min, np.finfo(np.float64).max)):
                dst = np.zeros(int(np.prod(ddims)))
                k.rrOp(OP[op], src.copy(), dims, strides(dims), dst, ddims, strides(ddims), dim)
                want = red(np.concatenate([m, np.full(ddims, start)], axis=dim), axis=dim).ravel()
                assert np.array_equal(dst, want, equal_nan=True), (dims, dim, op)
                if dim == 1:  # rows 3 and 4 reduce over both zeros
                    zero = 0.0 if op == "RR_MAX" else -0.0
                    assert np.signbit(dst[3]) == np.signbit(zero) and np.signbit(dst[4]) == np.signbit(zero), (dims, op)


def test_riop_many_lines_edge_cases(k):
    """The warp-per-line index kernel (many lines along the fastest dimension): unique extreme (one pass), ties
    (second pass), NaN (Math.max gives NaN, nothing equals it), a line of -infinity (the fold starts from
    -Double.MAX_VALUE: DimensionOps.java:181), against the definition."""
    rng = np.random.default_rng(21)
    rows, n = 5000, 96
    src = rng.uniform(-1, 1, rows * n)
    m = src.reshape(rows, n)
    m[1, :] = np.round(m[1, :] * 2)          # ties
    m[2, 7] = np.nan
    m[3, :] = -np.inf
    m[4, :] = 0.25                           # every position matches
    m[5, 40] = np.inf
    for op, red, start in (("RI_MAX", np.max, -np.finfo(np.float64).max), ("RI_MIN", np.min, np.finfo(np.float64).max)):
        if op == "RI_MIN":
            m[3, :] = np.inf
            m[5, 40] = -np.inf
        out = np.full(rows * n, -9, dtype=np.int32)
        k.riOp(OP[op], src.copy(), [rows, n], [n, 1], out, 1)
        out = out.reshape(rows, n)
        for r in range(rows):
            line = m[r]
            if np.isnan(line).any():
                want = []
            else:
                ext = red(np.append(line, start))
                want = list(np.nonzero(line == ext)[0])
            exp = np.full(n, -1, dtype=np.int32)
            exp[:len(want)] = want
            assert np.array_equal(out[r], exp), (op, r)


@needs_ref
@pytest.mark.parametrize("dims,order", SHAPES, ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else v)
def test_rdop_matches_reference(k, dims, order):
    rng = np.random.default_rng(13)
    n = int(np.prod(dims))
    st = strides(dims, order)
    nd = len(dims)
    for op, src in (("RD_SUM", rng.uniform(-1, 1, n)), ("RD_PROD", 1.0 + 0.01 * rng.uniform(-1, 1, n))):
        for opdims in [(d,) for d in range(nd)] + [tuple(range(nd)), ()]:
            a, b = np.zeros(n), np.zeros(n)
            k.rdOp(OP[op], src.copy(), dims, st, a, *opdims)
            ref.rdOp(OP[op], src.copy(), dims, st, b, *opdims)
            assert rel(a, b) < tol(n), (op, opdims)
    ints = rng.integers(-3, 4, n).astype(np.float64)  # integers: any summation order is exact
    a, b = np.zeros(n), np.zeros(n)
    k.rdOp(OP["RD_SUM"], ints, dims, st, a, *range(nd))
    ref.rdOp(OP["RD_SUM"], ints, dims, st, b, *range(nd))
    assert np.array_equal(a, b)


@needs_ref
def test_elementwise_matches_reference(k):
    rng = np.random.default_rng(14)
    for n in (1, 2, 7, 1000, 4097):
        x = rng.uniform(0.1, 2.0, n)
        for op, a in [("RU_ADD", 0.3), ("RU_MUL", -1.7), ("RU_ABS", 0), ("RU_SQRT", 0), ("RU_SQR", 0), ("RU_INV", 2.5),
                      ("RU_FILL", 4.25)]:  # IEEE-exact maps: bit-identical
            va, vb = (x - 1.0).copy() if op == "RU_ABS" else x.copy(), None
            vb = va.copy()
            k.ruOp(OP[op], a, va)
            ref.ruOp(OP[op], a, vb)
            assert np.array_equal(va, vb), op
        for op, a in [("RU_POW", 1.0 / 3), ("RU_EXP", 0), ("RU_LOG", 0), ("RU_COS", 0), ("RU_SIN", 0), ("RU_ATAN", 0)]:
            va, vb = x.copy(), x.copy()
            k.ruOp(OP[op], a, va)
            ref.ruOp(OP[op], a, vb)
            assert np.max(np.abs(va - vb) / np.maximum(np.abs(vb), 1e-300)) < 4e-16, op  # libm vs CUDA: <= 2 ulp
        z = rng.uniform(-1, 1, 2 * n)
        for op in ("CU_ADD", "CU_MUL", "CU_CONJ", "CU_FILL"):
            va, vb = z.copy(), z.copy()
            k.cuOp(OP[op], 0.5, -0.25, va)
            ref.cuOp(OP[op], 0.5, -0.25, vb)
            assert np.array_equal(va, vb), op
        for op in ("CU_EXP", "CU_COS", "CU_SIN"):
            va, vb = z.copy(), z.copy()
            k.cuOp(OP[op], 0, 0, va)
            ref.cuOp(OP[op], 0, 0, vb)
            assert rel(va, vb) < 1e-15, op
        y = rng.uniform(0.5, 2.0, n)
        for op in ("RE_ADD", "RE_SUB", "RE_MUL", "RE_DIV", "RE_MAX", "RE_MIN"):
            da, db = np.zeros(n), np.zeros(n)
            k.eOp(OP[op], x, y, da, False)
            ref.eOp(OP[op], x, y, db, False)
            assert np.array_equal(da, db), op
        w = rng.uniform(-1, 1, 2 * n)
        for op in ("CE_ADD", "CE_SUB", "CE_MUL", "CE_DIV"):
            da, db = np.zeros(2 * n), np.zeros(2 * n)
            k.eOp(OP[op], z, w, da, True)
            ref.eOp(OP[op], z, w, db, True)
            assert np.array_equal(da, db), op
        ia = rng.integers(-2 ** 31, 2 ** 31 - 1, n).astype(np.int32)
        ib = rng.integers(-2 ** 31, 2 ** 31 - 1, n).astype(np.int32)
        for op in ("IE_ADD", "IE_SUB", "IE_MUL", "IE_MAX", "IE_MIN"):  # Java int wrap-around
            da, db = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
            k.eOp(OP[op], ia, ib, da, False)
            ref.eOp(OP[op], ia, ib, db, False)
            assert np.array_equal(da, db), op
        for op, a in [("IU_ADD", 2 ** 31 - 5), ("IU_MUL", 65537), ("IU_FILL", -3)]:
            va, vb = ia.copy(), ia.copy()
            k.iuOp(OP[op], a, va)
            ref.iuOp(OP[op], a, vb)
            assert np.array_equal(va, vb), op
        # in-place binary (the l* operations pass dst = lhs)
        la, lb = x.copy(), x.copy()
        k.eOp(OP["RE_MUL"], la, y, la, False)
        ref.eOp(OP["RE_MUL"], lb, y, lb, False)
        assert np.array_equal(la, lb)
        for op, sc, dc, src_, nd in [("R_TO_C_RE", False, True, x, 2 * n), ("R_TO_C_IM", False, True, x, 2 * n),
                                     ("C_TO_R_ABS", True, False, z, n), ("C_TO_R_RE", True, False, z, n),
                                     ("C_TO_R_IM", True, False, z, n), ("I_TO_R", False, False, ia, n)]:
            da, db = np.zeros(nd), np.zeros(nd)
            k.convert(OP[op], src_, sc, da, dc)
            ref.convert(OP[op], src_, sc, db, dc)
            assert np.array_equal(da, db), op


@needs_ref
def test_accumulators_match_reference(k):
    rng = np.random.default_rng(15)
    for n in (1, 3, 1000, 100003, 1 << 20):
        x = rng.uniform(0.0, 1.0, n)
        for op in ("RA_SUM", "RA_VAR", "RA_MAX", "RA_MIN", "RA_ENT"):
            a, b = k.raOp(OP[op], x), ref.raOp(OP[op], x)
            if op in ("RA_MAX", "RA_MIN"):
                assert a == b
            else:
                assert abs(a - b) <= tol(n) * abs(b), (op, n, a, b)
        p = 1.0 + 1e-3 * rng.uniform(-1, 1, n)
        a, b = k.raOp(OP["RA_PROD"], p), ref.raOp(OP["RA_PROD"], p)
        assert abs(a - b) <= tol(n) * abs(b)
        z = rng.uniform(-1, 1, 2 * n)
        assert rel(k.caOp(OP["CA_SUM"], z), ref.caOp(OP["CA_SUM"], z)) < tol(n) * 10
        zp = (np.exp(1j * rng.uniform(-3, 3, n)) * (1.0 + 1e-4 * rng.uniform(-1, 1, n))).view(np.float64)
        assert rel(k.caOp(OP["CA_PROD"], zp), ref.caOp(OP["CA_PROD"], zp)) < tol(n) * 10
    assert k.raOp(OP["RA_MAX"], np.zeros(0)) == -np.finfo(np.float64).max  # Arithmetic.java:54 initial value
    assert k.raOp(OP["RA_SUM"], np.zeros(0)) == 0.0


@needs_ref
def test_map_slice_match_reference(k):
    rng = np.random.default_rng(16)
    for dtype in (np.float64, np.int32):
        for sdims, ddims, order in [([5, 7], [5, 7], "FAR"), ([4, 6, 3], [9, 5, 4], "FAR"), ([4, 6, 3], [9, 5, 4], "NEAR"),
                                    ([13], [29], "FAR"), ([3, 3, 3, 3], [2, 4, 3, 5], "FAR")]:
            ns, ndd = int(np.prod(sdims)), int(np.prod(ddims))
            src = (rng.uniform(-1, 1, ns) * 1000).astype(dtype)
            for trial in range(6):
                bounds = []
                for d in range(len(sdims)):
                    # sizes beyond either dimension (tiling), negative starts (wrap): MappingOps.java:237-249
                    bounds += [int(rng.integers(-9, 9)), int(rng.integers(-9, 9)), int(rng.integers(0, 2 * ddims[d] + 1))]
                a, b = np.full(ndd, 77, dtype=dtype), np.full(ndd, 77, dtype=dtype)
                k.map(bounds, src, sdims, strides(sdims, order), a, ddims, strides(ddims, order))
                ref.map(bounds, src, sdims, strides(sdims, order), b, ddims, strides(ddims, order))
                assert np.array_equal(a, b), (sdims, ddims, bounds)
                slices = []
                for d in range(len(sdims)):
                    for _ in range(int(rng.integers(0, 5)) if trial else 2):
                        slices += [int(rng.integers(0, sdims[d])), int(rng.integers(0, ddims[d])), d]
                a, b = np.full(ndd, 55, dtype=dtype), np.full(ndd, 55, dtype=dtype)
                k.slice(slices, src, sdims, strides(sdims, order), a, ddims, strides(ddims, order))
                ref.slice(slices, src, sdims, strides(sdims, order), b, ddims, strides(ddims, order))
                assert np.array_equal(a, b), (sdims, ddims, slices)
    # long last dimension (row-walking kernel) and mixed storage orders (tile-transposing kernel), with wrap / tiling
    for sdims, ddims, so, do in [([3, 700], [5, 1100], "FAR", "FAR"), ([2, 3, 1500], [3, 2, 900], "FAR", "FAR"),
                                 ([33, 47], [50, 41], "FAR", "NEAR"), ([40, 70, 3], [64, 64, 5], "NEAR", "FAR")]:
        ns, ndd = int(np.prod(sdims)), int(np.prod(ddims))
        src = rng.uniform(-1, 1, ns)
        for trial in range(6):
            bounds = []
            for d in range(len(sdims)):
                bounds += [int(rng.integers(-9, 9)), int(rng.integers(-9, 9)), int(rng.integers(0, 2 * ddims[d] + 1))]
            a, b = np.full(ndd, 77.0), np.full(ndd, 77.0)
            k.map(bounds, src, sdims, strides(sdims, so), a, ddims, strides(ddims, do))
            ref.map(bounds, src, sdims, strides(sdims, so), b, ddims, strides(ddims, do))
            assert np.array_equal(a, b), (sdims, ddims, bounds)
    # transpose expressed as permuted destination strides (ProtoArray.java:303-309)
    src = np.arange(24, dtype=np.float64)
    a, b = np.zeros(24), np.zeros(24)
    args = ([0, 0, 2, 0, 0, 3, 0, 0, 4], src, [2, 3, 4], [12, 4, 1])
    k.map(*args, a, [2, 3, 4], [1, 8, 2])
    ref.map(*args, b, [2, 3, 4], [1, 8, 2])
    assert np.array_equal(a, b) and np.array_equal(a.reshape(3, 4, 2), src.reshape(2, 3, 4).transpose(1, 2, 0))


@needs_ref
def test_mul_matches_reference(k):
    rng = np.random.default_rng(17)
    for (m, kk, n) in [(1, 1, 1), (3, 5, 4), (17, 33, 9), (128, 128, 128), (130, 67, 259), (257, 300, 131),
                       (96, 18, 70), (129, 16, 2), (64, 2, 300), (200, 48, 256), (2, 1000, 2)]:
        a, b = rng.uniform(-1, 1, m * kk), rng.uniform(-1, 1, kk * n)
        ca, cb = np.zeros(m * n), np.zeros(m * n)
        k.mul(a, b, m, n, ca, False)
        ref.mul(a, b, m, n, cb, False)
        assert rel(ca, cb) < tol(kk), (m, kk, n)
        za, zb = rng.uniform(-1, 1, 2 * m * kk), rng.uniform(-1, 1, 2 * kk * n)
        ca, cb = np.zeros(2 * m * n), np.zeros(2 * m * n)
        k.mul(za, zb, m, n, ca, True)
        ref.mul(za, zb, m, n, cb, True)
        assert rel(ca, cb) < tol(kk), (m, kk, n)
    c = np.ones(0)
    k.mul(np.zeros(0), np.zeros(0), 0, 0, c, False)  # lr == 0 / rc == 0 are legal (MatrixOps.java:46-55)


@needs_ref
def test_invert_matches_reference(k):
    rng = np.random.default_rng(31)
    for n in (1, 2, 3, 17, 64, 128, 200):
        a = rng.uniform(0, 1, (n, n)) + 2 * np.eye(n)  # MatrixTest.java:227-242 builds its inputs like this
        if n >= 3:
            a[[0, n - 1]] = a[[n - 1, 0]]              # force row exchanges
        got, want = np.full(n * n, 7.0), np.zeros(n * n)
        k.invert(a.ravel().copy(), got, n)
        ref.invert(a.ravel().copy(), want, n)
        assert np.abs(a @ got.reshape(n, n) - np.eye(n)).sum() < 1e-8, n  # the reference test's own criterion
        assert rel(got, want) < 1e-11, (n, rel(got, want))
    import shared_b200 as s
    with pytest.raises(s.SstCudaError, match="Matrix is singular"):
        k.invert(np.zeros(9), np.zeros(9), 3)
    with pytest.raises(s.SstCudaError, match="Invalid arguments"):
        k.invert(np.zeros(8), np.zeros(9), 3)
    k.invert(np.zeros(0), np.zeros(0), 0)


@needs_ref
def test_svd_matches_reference(k):
    """MatrixTest.java:186-203's criterion (|A - U S V^T| summed < 1e-8 over the divisor shapes of 120), singular values
    against the reference's JAMA port, orthonormal factors, and the singular vectors up to the sign of each pair."""
    rng = np.random.default_rng(33)
    shapes = [(120 // i, i) for i in range(1, 121) if 120 % i == 0 and 120 // i >= i] + [(128, 128), (300, 7), (9, 9)]
    for m, n in shapes:
        a = rng.uniform(0, 1, (m, n))
        for transposed in (False, True):
            if transposed:   # the same matrix handed over as the transpose of a row-major n x m array
                src, rs, cs = np.ascontiguousarray(a.T).ravel(), 1, m
            else:
                src, rs, cs = a.ravel().copy(), n, 1
            u, sv, v = np.zeros(m * n), np.zeros(n), np.zeros(n * n)
            ru, rsv, rv = np.zeros(m * n), np.zeros(n), np.zeros(n * n)
            k.svd(src, rs, cs, u, sv, v, m, n)
            ref.svd(src.copy(), rs, cs, ru, rsv, rv, m, n)
            U, V, RU, RV = u.reshape(m, n), v.reshape(n, n), ru.reshape(m, n), rv.reshape(n, n)
            assert np.abs(a - U @ np.diag(sv) @ V.T).sum() < 1e-8, (m, n, transposed)
            assert np.all(np.diff(sv) <= 0) and np.max(np.abs(sv - rsv)) < 1e-11 * max(1.0, rsv[0]), (m, n)
            assert np.max(np.abs(U.T @ U - np.eye(n))) < 1e-12 and np.max(np.abs(V.T @ V - np.eye(n))) < 1e-12
            gaps = np.abs(np.diff(np.concatenate([rsv, [0.0]])))
            for j in range(n):  # well-separated singular values: vectors agree up to a common sign
                if gaps[j] > 1e-2 * rsv[0] and (j == 0 or gaps[j - 1] > 1e-2 * rsv[0]):
                    sign = np.sign(U[:, j] @ RU[:, j])
                    assert np.max(np.abs(sign * U[:, j] - RU[:, j])) < 1e-8 and np.max(np.abs(sign * V[:, j] - RV[:, j])) < 1e-8
    import shared_b200 as s
    with pytest.raises(s.SstCudaError, match="Invalid arguments"):
        k.svd(np.zeros(6), 3, 1, np.zeros(6), np.zeros(3), np.zeros(9), 2, 3)      # wide matrices are refused
    with pytest.raises(s.SstCudaError, match="Invalid arguments"):
        k.svd(np.zeros(6), 2, 2, np.zeros(6), np.zeros(2), np.zeros(4), 3, 2)      # neither layout
    z = np.zeros(6)
    u, sv, v = np.ones(6), np.ones(2), np.ones(4)
    k.svd(z, 2, 1, u, sv, v, 3, 2)                                                 # the zero matrix: S = 0, A = U S V^T
    assert np.all(sv == 0) and np.all(np.isfinite(u)) and np.allclose(v.reshape(2, 2).T @ v.reshape(2, 2), np.eye(2))


@needs_ref
def test_find_matches_reference(k):
    rng = np.random.default_rng(32)
    for dims, order in (([7], "FAR"), ([5, 9], "FAR"), ([5, 9], "NEAR"), ([3, 4, 300], "FAR"), ([3, 4, 5], "NEAR")):
        n = int(np.prod(dims))
        src = rng.uniform(-1, 1, n)
        idx = np.zeros(n, dtype=np.int32)
        st = strides(dims, order)
        for dim in range(len(dims)):
            k.riOp(OP["RI_GZERO"], src.copy(), dims, st, idx, dim)  # positions of the positive entries, then -1
            for _ in range(6):
                logical = [int(rng.integers(0, d)) for d in dims]
                logical[dim] = -1
                got, want = k.find(idx, dims, st, logical), ref.find(idx, dims, st, logical)
                assert got.dtype == np.int32 and np.array_equal(got, want), (dims, order, logical)
    import shared_b200 as s
    v = np.arange(8, dtype=np.int32)
    with pytest.raises(s.SstCudaError, match="Invalid index"):
        k.find(v, [2, 4], [4, 1], [2, -1])
    with pytest.raises(s.SstCudaError, match="Invalid arguments"):
        k.find(v, [2, 4], [4, 1], [-1, -1])
    with pytest.raises(s.SstCudaError, match="Invalid dimensions and/or strides"):
        k.find(v, [2, 4], [1, 1], [0, -1])


def test_error_messages(k):
    import shared_b200 as s
    E = s.SstCudaError
    with pytest.raises(E, match="Invalid dimensions and/or strides"):
        k.rrOp(0, np.zeros(6), [2, 3], [1, 1], np.zeros(2), [2, 1], [1, 1], 1)
    with pytest.raises(E, match="Operation type not recognized"):
        k.ruOp(99, 0.0, np.zeros(3))
    with pytest.raises(E, match="Duplicate operating dimensions are not allowed"):
        k.rrOp(0, np.zeros(6), [2, 3], [3, 1], np.zeros(1), [1, 1], [1, 1], 1, 1)
    with pytest.raises(E, match="Operating dimensions must have singleton or zero length"):
        k.rrOp(0, np.zeros(6), [2, 3], [3, 1], np.zeros(6), [2, 3], [3, 1], 1)
    with pytest.raises(E, match="Invalid dimension"):
        k.riOp(0, np.zeros(6), [2, 3], [3, 1], np.zeros(6, dtype=np.int32), 2)
    with pytest.raises(E, match="Invalid index"):
        k.slice([0, 5, 0], np.zeros(4), [4], [1], np.zeros(4), [4], [1])
    with pytest.raises(E, match="Invalid mapping parameters"):
        k.map([0, 0, -1], np.zeros(4), [4], [1], np.zeros(4), [4], [1])
    with pytest.raises(E, match="Invalid array lengths"):
        k.mul(np.zeros(6), np.zeros(8), 2, 2, np.zeros(4), False)
    with pytest.raises(E, match="Invalid array lengths"):
        k.eOp(0, np.zeros(3), np.zeros(4), np.zeros(3), False)
    with pytest.raises(E, match="Invalid array types"):  # ObjectArrayTest.java:289-305
        k.map([0, 0, 1], np.zeros(1), [1], [1], np.zeros(1, dtype=np.int32), [1], [1])


def test_rng_ops_statistics(k):
    # ArrayKernelTest.java:151-172,219-251: RND is uniform (quartile bins within 0.005), SHUFFLE is a permutation
    k.derandomize()
    n = 1 << 16
    v = np.zeros(n)
    k.ruOp(OP["RU_RND"], math.pi, v)
    assert v.min() >= 0 and v.max() < math.pi
    hist, _ = np.histogram(v, bins=4, range=(0, math.pi))
    assert np.max(np.abs(hist / n - 0.25)) < 0.005 * 4
    z = np.zeros(2 * n)
    k.cuOp(OP["CU_RND"], math.e, math.pi, z)
    assert 0 <= z[0::2].min() and z[0::2].max() < math.e and z[1::2].max() < math.pi
    for arr, op, fn in [(np.arange(1000, dtype=np.float64), "RU_SHUFFLE", lambda a: k.ruOp(OP["RU_SHUFFLE"], 0, a)),
                        (np.arange(1000, dtype=np.int32), "IU_SHUFFLE", lambda a: k.iuOp(OP["IU_SHUFFLE"], 0, a))]:
        before = arr.copy()
        fn(arr)
        assert not np.array_equal(arr, before) and np.array_equal(np.sort(arr), before), op
    c = np.arange(2000, dtype=np.float64)
    k.cuOp(OP["CU_SHUFFLE"], 0, 0, c)
    pairs = c.reshape(-1, 2)
    assert np.all(pairs[:, 1] == pairs[:, 0] + 1) and np.array_equal(np.sort(pairs[:, 0]), np.arange(0, 2000, 2))


# ---- (3) BASELINE.json configs[3] sizes on device-resident data: size-independent properties ----------------

def test_c4_device_tier_properties():
    import torch
    from shared_b200.array_kernel import REAL, DeviceArrayKernel
    n = 8192
    dk = DeviceArrayKernel(0, torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(20261019)
    x = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    xm = x.view(n, n)
    far = [n, 1]
    # rrOp SUM / MAX / VAR over dim 1 (contiguous) and dim 0 (strided) vs torch
    for dim, ddims in ((1, [n, 1]), (0, [1, n])):
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        dk.rrOp(OP["RR_SUM"], x.data_ptr(), n * n, [n, n], far, out.data_ptr(), n, ddims, [ddims[1], 1], dim)
        assert (torch.linalg.norm(out - xm.sum(dim)) / torch.linalg.norm(xm.sum(dim))).item() < 1e-11
        dk.rrOp(OP["RR_MAX"], x.data_ptr(), n * n, [n, n], far, out.data_ptr(), n, ddims, [ddims[1], 1], dim)
        assert torch.equal(out, xm.max(dim).values)
        dk.rrOp(OP["RR_VAR"], x.data_ptr(), n * n, [n, n], far, out.data_ptr(), n, ddims, [ddims[1], 1], dim)
        want = xm.var(dim, unbiased=False)
        assert (torch.linalg.norm(out - want) / torch.linalg.norm(want)).item() < 1e-11
    # rdOp SUM over dim 1, dim 0, both: last element of a scan = the reduction; difference of neighbours = input
    y = torch.empty_like(x)
    for dims in ((1,), (0,), (0, 1)):
        dk.rdOp(OP["RD_SUM"], x.data_ptr(), n * n, [n, n], far, y.data_ptr(), n * n, *dims)
        want = xm
        for d in dims:
            want = want.cumsum(d)
        assert (torch.linalg.norm(y.view(n, n) - want) / torch.linalg.norm(want)).item() < 1e-11
    # riOp MAX over dim 1: first slot = argmax, remaining slots -1 (ties have probability ~0 here)
    idx = torch.empty(n * n, dtype=torch.int32, device="cuda")
    dk.riOp(OP["RI_MAX"], x.data_ptr(), n * n, [n, n], far, idx.data_ptr(), n * n, 1)
    im = idx.view(n, n)
    assert torch.equal(im[:, 0].long(), xm.argmax(1)) and bool((im[:, 1:] == -1).all())
    # ruOp / eOp round trips
    z = x.clone()
    dk.ruOp(OP["RU_MUL"], 2.0, z.data_ptr(), n * n)
    assert torch.equal(z, x * 2.0)
    dk.eOp(OP["RE_ADD"], REAL, x.data_ptr(), z.data_ptr(), y.data_ptr(), n * n)
    assert torch.equal(y, x + z)
    dk.ruOp(OP["RU_EXP"], 0.0, z.data_ptr(), n * n)
    assert (torch.linalg.norm(z - torch.exp(2.0 * x)) / torch.linalg.norm(z)).item() < 1e-15
    # mul 4096^2 vs cuBLAS (yardstick)
    m = 4096
    a = torch.rand(m * m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    b = torch.rand(m * m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    c = torch.empty(m * m, dtype=torch.float64, device="cuda")
    dk.mul(a.data_ptr(), m * m, b.data_ptr(), m * m, m, m, c.data_ptr(), m * m, False)
    want = a.view(m, m) @ b.view(m, m)
    assert (torch.linalg.norm(c.view(m, m) - want) / torch.linalg.norm(want)).item() < 1e-12 * 12
    torch.cuda.synchronize()


@needs_ref
def test_c4_full_size_against_reference(k):
    """BASELINE.json configs[3]'s own size, 8192 x 8192, against the reference's compiled C++ on the same input (host
    tier): riOp and the reductions' index / extremum results bit-exact, the scans and sums within 1e-12 * log2 N. The
    integer-valued input makes every partial sum exact, so the scan must agree BIT FOR BIT whatever the summation order."""
    n = 8192
    dims, st = [n, n], [n, 1]
    rng = np.random.default_rng(8192)
    x = rng.integers(-1000, 1000, n * n).astype(np.float64)
    # riOp MAX / MIN along dim 1 (bit-exact: positions of the extrema, then -1 padding; ties are frequent here)
    for op in ("RI_MAX", "RI_MIN"):
        want, got = np.zeros(n * n, dtype=np.int32), np.zeros(n * n, dtype=np.int32)
        ref.riOp(OP[op], x.copy(), dims, st, want, 1)
        k.riOp(OP[op], x.copy(), dims, st, got, 1)
        assert np.array_equal(got, want), op
    # rdOp SUM / PROD-free scans along dim 1, dim 0 and both: exact on integers
    for opdims in ((1,), (0,), (0, 1)):
        want, got = np.zeros(n * n), np.zeros(n * n)
        ref.rdOp(OP["RD_SUM"], x, dims, st, want, *opdims)
        k.rdOp(OP["RD_SUM"], x, dims, st, got, *opdims)
        assert np.array_equal(got, want), opdims
    # rrOp SUM / MAX / MIN over each dimension: exact on integers
    for dim, dd, ds in ((1, [n, 1], [1, 1]), (0, [1, n], [n, 1])):
        for op in ("RR_SUM", "RR_MAX", "RR_MIN"):
            want, got = np.zeros(n), np.zeros(n)
            ref.rrOp(OP[op], x, dims, st, want, dd, ds, dim)
            k.rrOp(OP[op], x, dims, st, got, dd, ds, dim)
            assert np.array_equal(got, want), (op, dim)
    # on reals: the scan and the variance within the floating-point tolerance
    y = rng.uniform(-0.5, 0.5, n * n)
    want, got = np.zeros(n * n), np.zeros(n * n)
    ref.rdOp(OP["RD_SUM"], y, dims, st, want, 0, 1)
    k.rdOp(OP["RD_SUM"], y, dims, st, got, 0, 1)
    assert rel(got, want) < tol(n * n)
    want, got = np.zeros(n), np.zeros(n)
    ref.rrOp(OP["RR_VAR"], y, dims, st, want, [n, 1], [1, 1], 1)
    k.rrOp(OP["RR_VAR"], y, dims, st, got, [n, 1], [1, 1], 1)
    assert rel(got, want) < tol(n * n)
    # ruOp / eOp: IEEE-exact maps, bit for bit
    a, b = y.copy(), y.copy()
    ref.ruOp(OP["RU_MUL"], 1.0000001, a)
    k.ruOp(OP["RU_MUL"], 1.0000001, b)
    assert np.array_equal(a, b)
    want, got = np.zeros(n * n), np.zeros(n * n)
    ref.eOp(OP["RE_ADD"], x, y, want, False)
    k.eOp(OP["RE_ADD"], x, y, got, False)
    assert np.array_equal(got, want)
