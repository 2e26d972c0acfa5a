"""Parity of the CUDA FftService provider (through the C ABI) with the CPU oracle and with the reference's
own known-answer tests. Tolerance: relative L2 <= 1e-12 * log2(N) (BASELINE.json north_star); the
reference's KATs additionally at its own absolute 1e-8 (src_test/org/shared/test/Tests.java:117)."""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

KAT_TOL = 1e-8


def rel_tol(n):
    return 1e-12 * max(1.0, math.log2(max(2, n)))


def rel_l2(got, want):
    d = np.linalg.norm(want)
    return np.linalg.norm(got - want) / (d if d > 0 else 1.0)


@pytest.fixture(scope="module")
def svc():
    import shared_b200
    return shared_b200.CudaFftService()


def cmul(a, b):
    return (a.view(np.complex128) * b.view(np.complex128)).view(np.float64)


def conj(a):
    return np.conj(a.view(np.complex128)).view(np.float64)


def run(svc, name, dims, x):
    import shared_b200 as s
    kind = {"fft": s.FORWARD, "ifft": s.BACKWARD, "rfft": s.R2C, "rifft": s.C2R}[name]
    _, out_len = s.transform_lengths(kind, dims)
    out = np.full(out_len, np.nan)
    x = np.ascontiguousarray(x, dtype=np.float64)
    keep = x.copy()
    getattr(svc, name)(dims, x, out)
    assert np.array_equal(x, keep), "input must be preserved"
    return out


# ---- the reference's own KATs, replayed through the CUDA provider ------------------------------------

def test_kat_1d(svc, fft_kats):
    recs = fft_kats["testOneDimensionalFft"]
    a, b, exp = [np.array(r["values"]) for r in recs[:3]]          # ArrayFftTest.java:61-80
    got = run(svc, "ifft", [3], cmul(run(svc, "fft", [3], a), run(svc, "fft", [3], b)))
    assert np.max(np.abs(got - exp)) < KAT_TOL
    a, b, exp = [np.array(r["values"]) for r in recs[3:6]]         # ArrayFftTest.java:82-107
    got = run(svc, "rifft", [27], cmul(run(svc, "rfft", [27], a), run(svc, "rfft", [27], b)))
    assert np.max(np.abs(got - exp)) < KAT_TOL


def test_kat_2d(svc, fft_kats):
    a, b, exp = [np.array(r["values"]) for r in fft_kats["testTwoDimensionalFft"]]  # ArrayFftTest.java:116-160
    d = [7, 7]
    got = run(svc, "rifft", d, cmul(run(svc, "rfft", d, a), conj(run(svc, "rfft", d, b))))
    assert np.max(np.abs(got - exp)) < KAT_TOL
    ac = np.stack([a, np.zeros_like(a)], -1).ravel()
    bc = np.stack([b, np.zeros_like(b)], -1).ravel()
    got = run(svc, "ifft", d, cmul(run(svc, "fft", d, ac), conj(run(svc, "fft", d, bc))))[0::2]
    assert np.max(np.abs(got - exp)) < KAT_TOL


def test_kat_3d(svc, fft_kats):
    a, b, exp = [np.array(r["values"]) for r in fft_kats["testThreeDimensionalFft"]]  # ArrayFftTest.java:169-224
    d = [3, 3, 3]
    ac = np.stack([a, np.zeros_like(a)], -1).ravel()
    bc = np.stack([b, np.zeros_like(b)], -1).ravel()
    got = run(svc, "ifft", d, cmul(run(svc, "fft", d, ac), run(svc, "fft", d, bc)))[0::2]
    assert np.max(np.abs(got - exp)) < KAT_TOL
    got = run(svc, "rifft", d, cmul(run(svc, "rfft", d, a), run(svc, "rfft", d, b)))
    assert np.max(np.abs(got - exp)) < KAT_TOL


def test_kat_convolve(svc, conv_kats):
    # ConvolutionCacheTest.java:56-91 through ConvolutionCache.java:99-138
    a, b, exp = [np.array(r["values"]) for r in conv_kats["testConvolve"]]
    d = [6, 6]
    ker = np.zeros((6, 6))
    ker[:3, :3] = b.reshape(3, 3)
    kerf = conj(run(svc, "rfft", d, ker.ravel()))
    res = run(svc, "rifft", d, cmul(run(svc, "rfft", d, a), kerf)).reshape(6, 6)[:4, :4]
    assert np.max(np.abs(res.ravel() - exp)) < KAT_TOL


# ---- seeded parity against the oracle -----------------------------------------------------------------

DIMS = [
    [1], [2], [3], [4], [5], [7], [8], [9], [16], [27], [32], [64], [100], [128], [256], [512], [1000], [1024],
    [2048], [4096], [4099],
    [1, 1], [1, 8], [8, 1], [6, 5], [7, 7], [16, 8], [17, 64], [64, 17], [256, 256], [4, 1024], [1024, 4],
    [12, 4096], [2048, 6], [96, 80],
    [3, 3, 3], [4, 6, 5], [8, 8, 8], [64, 64, 64], [128, 32, 16], [5, 512, 3], [2, 4, 1, 8],
    [5, 6, 5, 6, 5], [2, 3, 4, 5, 6],
    # axes longer than the tile kernels hold: four-step (powers of two) and Bluestein (anything else)
    [8192], [65536], [1 << 20], [16384, 3], [3, 8192], [10007], [9000], [7, 12000], [12000, 5], [2, 7200, 3],
]


@pytest.mark.parametrize("dims", DIMS, ids=lambda d: "x".join(map(str, d)))
def test_c2c_matches_oracle(svc, dims):
    rng = np.random.default_rng(20261019)
    n = int(np.prod(dims))
    x = rng.uniform(-0.5, 0.5, 2 * n)
    want = oracle.fft(dims, x)
    got = run(svc, "fft", dims, x)
    assert rel_l2(got, want) < rel_tol(n)
    wantb = oracle.ifft(dims, want)
    gotb = run(svc, "ifft", dims, want)
    assert rel_l2(gotb, wantb) < rel_tol(n)
    assert rel_l2(gotb, x) < rel_tol(n)


@pytest.mark.parametrize("dims", DIMS, ids=lambda d: "x".join(map(str, d)))
def test_r2c_c2r_match_oracle(svc, dims):
    rng = np.random.default_rng(7)
    n = int(np.prod(dims))
    x = rng.uniform(-0.5, 0.5, n)
    want = oracle.rfft(dims, x)
    got = run(svc, "rfft", dims, x)
    assert rel_l2(got, want) < rel_tol(n)
    wantb = oracle.rifft(dims, want)
    gotb = run(svc, "rifft", dims, want)
    assert rel_l2(gotb, wantb) < rel_tol(n)
    assert rel_l2(gotb, x) < rel_tol(n)


def test_c1_config_roundtrip(svc):
    # BASELINE.json configs[0]: ComplexArray 256x256 2-D fft/ifft round trip
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, 2 * 256 * 256)
    f = run(svc, "fft", [256, 256], x)
    assert rel_l2(f, oracle.fft([256, 256], x)) < rel_tol(65536)
    assert rel_l2(run(svc, "ifft", [256, 256], f), x) < rel_tol(65536)


def test_length_errors(svc):
    import shared_b200 as s
    with pytest.raises(s.SstCudaError, match="Input and/or output arrays do not have expected sizes"):
        svc.fft([4, 4], np.zeros(32), np.zeros(30))  # Plan.cpp:88-90
    with pytest.raises(s.SstCudaError, match="Invalid dimensions"):
        svc.fft([4, 0], np.zeros(0), np.zeros(0))
    with pytest.raises(ValueError, match="Unknown hint"):
        svc.setHint("nope", "x")  # JavaFftService.java:96-99
    with pytest.raises(ValueError, match="Unknown hint"):
        svc.getHint("nope")


# ---- full-size properties on device-resident data (BASELINE.json configs[2]: 512^3) ----------------------

def _dev_plan(kind, dims):
    import shared_b200 as s
    return s.FftPlan(kind, dims)


@pytest.mark.parametrize("dims", [[128, 128, 128], [512, 512, 512]], ids=["128c", "512c"])
def test_3d_device_properties(dims):
    import torch
    import shared_b200 as s
    n = int(np.prod(dims))
    tol = rel_tol(n)
    g = torch.Generator(device="cuda").manual_seed(20261019)
    x = torch.rand(n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    fwd, bwd = _dev_plan(s.FORWARD, dims), _dev_plan(s.BACKWARD, dims)
    stream = torch.cuda.current_stream().cuda_stream
    x0 = x.clone()
    fwd.execute(x.data_ptr(), y.data_ptr(), stream)
    assert torch.equal(x, x0), "input must be preserved"
    # (1) independent library on the same box (cuFFT via torch) -- a yardstick, not the oracle
    ref = torch.view_as_real(torch.fft.fftn(torch.view_as_complex(x).reshape(dims))).reshape(n, 2)
    err = (torch.linalg.norm(y - ref) / torch.linalg.norm(ref)).item()
    assert err < tol, err
    del ref
    # (2) Parseval: sum |X|^2 = N sum |x|^2
    ex, ey = (x * x).sum().item(), (y * y).sum().item()
    assert abs(ey / (n * ex) - 1.0) < 1e-12
    # (3) DC bin = sum of the input
    assert abs(y[0, 0].item() - x[:, 0].sum().item()) < 1e-9 * n ** 0.5
    # (4) round trip: ifft(fft(x)) = x
    bwd.execute(y.data_ptr(), z.data_ptr(), stream)
    err = (torch.linalg.norm(z - x) / torch.linalg.norm(x)).item()
    assert err < tol, err
    # (5) the transposed-pass schedule and the in-place schedule are the same arithmetic: bit-identical
    fwd.set_hint("transposed_passes", 0)
    assert fwd.workspace_bytes == 0
    fwd.execute(x.data_ptr(), z.data_ptr(), stream)
    assert torch.equal(z, y)
    fwd.set_hint("transposed_passes", 1)
    assert fwd.workspace_bytes == 16 * n
    z.zero_()
    fwd.execute(x.data_ptr(), z.data_ptr(), stream)
    assert torch.equal(z, y)
    # (6) a unit impulse at flat index j transforms to a pure phase of modulus 1
    x.zero_()
    x[12345, 0] = 1.0
    fwd.execute(x.data_ptr(), y.data_ptr(), stream)
    mod = (y * y).sum(dim=1)
    assert (mod - 1.0).abs().max().item() < 1e-12
    torch.cuda.synchronize()


def test_1024_cube_single_gpu():
    """BASELINE configs[4]'s size on ONE GPU (2^30 points, 16 GiB per array: more than a Java array can hold, so it is
    reachable through the device tier only): round trip, Parseval, the DC bin and an impulse -- 64-bit indexing."""
    import torch
    import shared_b200 as s
    free, _ = torch.cuda.mem_get_info()
    if free < 56 * (1 << 30):
        pytest.skip("needs 3 x 16 GiB of device memory")
    dims = [1024, 1024, 1024]
    n = 1 << 30
    tol = rel_tol(n)
    x = torch.empty(n, 2, dtype=torch.float64, device="cuda")
    x.uniform_(-0.5, 0.5, generator=torch.Generator(device="cuda").manual_seed(7))
    y, z = torch.empty_like(x), torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    fwd, bwd = _dev_plan(s.FORWARD, dims), _dev_plan(s.BACKWARD, dims)
    fwd.execute(x.data_ptr(), y.data_ptr(), st)
    ex = sum((x[i:i + (1 << 27)] ** 2).sum().item() for i in range(0, n, 1 << 27))
    ey = sum((y[i:i + (1 << 27)] ** 2).sum().item() for i in range(0, n, 1 << 27))
    assert abs(ey / (n * ex) - 1.0) < 1e-12
    assert abs(y[0, 0].item() - x[:, 0].sum().item()) < 1e-9 * n ** 0.5
    bwd.execute(y.data_ptr(), z.data_ptr(), st)
    z.sub_(x)
    err = sum((z[i:i + (1 << 27)] ** 2).sum().item() for i in range(0, n, 1 << 27)) ** 0.5 / ex ** 0.5
    assert err < tol, err
    x.zero_()
    x[(1 << 30) - 12345, 0] = 1.0     # an impulse far beyond 2^31 bytes
    fwd.execute(x.data_ptr(), y.data_ptr(), st)
    worst = max(((y[i:i + (1 << 27)] ** 2).sum(dim=1) - 1.0).abs().max().item() for i in range(0, n, 1 << 27))
    assert worst < 1e-12
    torch.cuda.synchronize()


def test_3d_against_oracle_64(svc):
    # the largest 3-D cube the CPU oracle finishes in about a second
    dims = [64, 64, 64]
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, 2 * 64 ** 3)
    assert rel_l2(run(svc, "fft", dims, x), oracle.fft(dims, x)) < rel_tol(64 ** 3)


def test_axis_pass_matches_numpy():
    import torch
    import shared_b200 as s
    s.init(0)
    rng = np.random.default_rng(5)
    for (outer, length, inner) in [(3, 512, 24), (1, 64, 1000), (7, 256, 1), (2, 12, 5), (5, 1024, 9)]:
        x = rng.uniform(-0.5, 0.5, (outer, length, inner, 2))
        want = np.fft.fft(x.view(np.complex128)[..., 0], axis=1)
        d = torch.from_numpy(x).cuda()
        o = torch.empty_like(d)
        s.fft_axis(d.data_ptr(), o.data_ptr(), outer, length, inner, False, 1.0,
                   torch.cuda.current_stream().cuda_stream)
        got = o.cpu().numpy().view(np.complex128)[..., 0]
        assert rel_l2(got, want) < rel_tol(outer * length * inner)
        s.fft_axis(o.data_ptr(), o.data_ptr(), outer, length, inner, True, 1.0 / length,
                   torch.cuda.current_stream().cuda_stream)  # in place
        back = o.cpu().numpy()
        assert rel_l2(back, x) < rel_tol(outer * length * inner)
