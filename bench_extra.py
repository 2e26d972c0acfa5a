"""bench_extra.py -- the other BASELINE.json configs, measured next to the headline so that the driver's BENCH / SCALE
records carry them (bench.py puts the list under "extra_configs" on its one JSON line).

  C1  configs[0]  ComplexArray 256x256 fft/ifft round trip                      device tier + host tier; CPU: the oracle at 256^2
  C1b the reference's own benchmark loop: 2048 x [eMul -> 2-D ifft] at 256^2    (native/src/sharedx/Benchmark.cpp:59-116,
                                                                                src_test/org/sharedx/test/BenchmarkSpecification.java:43-53)
  C2  configs[1]  RealArray 4096x4096 rfft / rifft / FFT-based 2-D correlation  parity against the oracle AT FULL SIZE
  C4  configs[3]  RealArray 8192x8192 ArrayKernel families + 4096^2 mul          CPU: the reference's own C++ (oracle/_ref)
  C5  configs[4]  ComplexArray 1024^3 fft/ifft round trip on 8 B200              (N = 8 only)

Every entry: the workload, CUDA-event time, the metric, a roofline fraction against MEASURED_PEAKS.json where the
kernel is HBM-bound (algorithmic bytes: SURVEY.md 8d), a parity figure, the clocks sampled while it ran, and a
cpu_baseline timed on this box's host ("reference" = the reference's compiled C++, "port" = the C restatement of its
Java; both single-threaded, as the reference is). The oracle is used here ONLY as the CPU baseline and as the checker.
"""
import ctypes
import math
import os
import time

import numpy as np


def _ev(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def _rel(torch, got, want):
    return (torch.linalg.norm(got - want) / torch.linalg.norm(want)).item()


def _timed_cpu(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


class Window:
    """Clock record of one config: the sampler rows between enter and exit."""

    def __init__(self, sampler):
        self.sampler = sampler

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *exc):
        self.t1 = time.time()

    def clocks(self):
        return self.sampler.summary(self.t0, self.t1) if self.sampler else None


def config_c1(torch, s, peak, sampler, with_cpu):
    import oracle
    st = torch.cuda.current_stream().cuda_stream
    dims = [256, 256]
    n = 256 * 256
    flops = 2 * 5.0 * n * math.log2(n)  # fft + ifft
    with Window(sampler) as w:
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.rand(n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5
        y, z = torch.empty_like(x), torch.empty_like(x)
        f, b = s.FftPlan(s.FORWARD, dims), s.FftPlan(s.BACKWARD, dims)
        ms = _ev(torch, lambda: (f.execute(x.data_ptr(), y.data_ptr(), st), b.execute(y.data_ptr(), z.data_ptr(), st)), 200)
        rt = _rel(torch, z, x)
        # through the provider call a JVM caller makes (pageable arrays, H2D + D2H inside)
        svc = s.CudaFftService(0)
        hx = x.cpu().numpy().reshape(-1).copy()
        hy, hz = np.empty_like(hx), np.empty_like(hx)
        svc.fft(dims, hx, hy)
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            svc.fft(dims, hx, hy)
            svc.ifft(dims, hy, hz)
        e2e_ms = (time.perf_counter() - t0) / reps * 1e3
        want = oracle.fft(dims, hx)
        parity = float(np.linalg.norm(hy - want) / np.linalg.norm(want))
    out = {"config": "C1: ComplexArray 256x256 complex128 fft + ifft round trip (BASELINE configs[0])",
           "metric": "GFLOP/s (5N*log2N, both transforms)", "value": flops / ms / 1e6, "unit": "GFLOP/s", "ms": ms,
           "e2e": {"ms": e2e_ms, "value": flops / e2e_ms / 1e6, "unit": "GFLOP/s", "api": "CudaFftService.fft + .ifft on "
                   "pageable host arrays (1 MiB each way per call)"},
           "roofline": None, "bound": "launch latency (4 launches of ~4 us; 2 MiB of traffic)",
           "parity": {"rel_l2_vs_oracle": parity, "roundtrip_rel_l2": rt, "tolerance": 1e-12 * math.log2(n)},
           "clocks": w.clocks()}
    if with_cpu:
        oracle.fft(dims, hx)
        secs, _ = _timed_cpu(lambda: [(oracle.fft(dims, hx), oracle.ifft(dims, hy)) for _ in range(10)])
        secs /= 10
        out["cpu_baseline"] = {"value": flops / secs / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port", "seconds": secs,
                               "sample": "the same 256x256 fft + ifft (the reference's own CPU-runnable case), oracle = C "
                                         "restatement of FftOps.java"}
    return out


def config_c1b(torch, s, peak, sampler, with_cpu):
    """2048 x [complex eMul -> 2-D ifft (1/N inside)] on 256 x 256 complex128, after 64 fft/ifft warm-ups."""
    import oracle
    from shared_b200.array_kernel import COMPLEX, DeviceArrayKernel
    st = torch.cuda.current_stream().cuda_stream
    dk = DeviceArrayKernel(0, st)
    dims = [256, 256]
    n = 256 * 256
    reps = 2048
    with Window(sampler) as w:
        g = torch.Generator(device="cuda").manual_seed(2)
        a = torch.rand(n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5
        b = torch.rand(n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5
        prod, out = torch.empty_like(a), torch.empty_like(a)
        f, inv = s.FftPlan(s.FORWARD, dims), s.FftPlan(s.BACKWARD, dims)
        for _ in range(64):
            f.execute(a.data_ptr(), out.data_ptr(), st)
            inv.execute(out.data_ptr(), prod.data_ptr(), st)

        def body():
            dk.eOp(2, COMPLEX, a.data_ptr(), b.data_ptr(), prod.data_ptr(), 2 * n)  # CE_MUL
            inv.execute(prod.data_ptr(), out.data_ptr(), st)

        body()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            body()
        e1.record()
        torch.cuda.synchronize()
        eager_ms = e0.elapsed_time(e1)
        # the same loop recorded once and replayed (what a device-resident caller would do): launch cost off the host
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            dk_side = DeviceArrayKernel(0, side.cuda_stream)
            with torch.cuda.graph(graph, stream=side):
                for _ in range(64):
                    dk_side.eOp(2, COMPLEX, a.data_ptr(), b.data_ptr(), prod.data_ptr(), 2 * n)
                    inv.execute(prod.data_ptr(), out.data_ptr(), side.cuda_stream)
        graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps // 64):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1)
        ha, hb = a.cpu().numpy().reshape(-1), b.cpu().numpy().reshape(-1)
        hp = (ha.view(np.complex128) * hb.view(np.complex128)).view(np.float64)
        want = oracle.ifft(dims, hp)
        got = out.cpu().numpy().reshape(-1)
        parity = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    res = {"config": "C1b: the reference's benchmark loop, 2048 x [eMul -> 2-D ifft] at 256x256 complex128 "
                     "(Benchmark.cpp:59-116, BenchmarkSpecification.java:43-53), device-resident",
           "metric": "milliseconds for the 2048 repetitions", "value": graph_ms, "unit": "ms", "higher_is_better": False,
           "ms": graph_ms, "eager_ms": eager_ms, "us_per_repetition": graph_ms / reps * 1e3,
           "roofline": None, "bound": "launch / dependency latency: 3 kernels of ~2-3 us each per repetition, 3 MiB of traffic",
           "parity": {"rel_l2_vs_oracle": parity, "tolerance": 1e-12 * math.log2(n)}, "clocks": w.clocks()}
    if with_cpu:
        cp = np.empty_like(hp)

        def cpu_loop(k):
            for _ in range(k):
                np.multiply(ha.view(np.complex128), hb.view(np.complex128), out=cp.view(np.complex128))
                oracle.ifft(dims, cp)
        cpu_loop(4)
        k = 128
        secs, _ = _timed_cpu(lambda: cpu_loop(k))
        res["cpu_baseline"] = {"value": secs / k * reps * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                               "sample": "%d of the 2048 repetitions, scaled; oracle ifft = C restatement of FftOps.java "
                                         "(BenchmarkJava's pure-Java arm), numpy complex multiply for eMul" % k}
    return res


def config_c2(torch, s, peak, sampler, with_cpu):
    import oracle
    from shared_b200 import _lib
    st = torch.cuda.current_stream().cuda_stream
    d = [4096, 4096]
    n = 4096 * 4096
    h = 4096 * 2049
    flops = 2.5 * n * math.log2(n)
    with Window(sampler) as w:
        g = torch.Generator(device="cuda").manual_seed(3)
        r = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
        spec = torch.empty(h, 2, dtype=torch.float64, device="cuda")
        back = torch.empty(n, dtype=torch.float64, device="cuda")
        rf, ri = s.FftPlan(s.R2C, d), s.FftPlan(s.C2R, d)
        ms_f = _ev(torch, lambda: rf.execute(r.data_ptr(), spec.data_ptr(), st), 20)
        ms_i = _ev(torch, lambda: ri.execute(spec.data_ptr(), back.data_ptr(), st), 20)
        rt = _rel(torch, back, r)
        # FFT-based valid-region correlation with a 31 x 31 kernel (ConvolutionCache.convolve, ConvolutionCache.java:99-138)
        ker = torch.zeros(4096, 4096, dtype=torch.float64, device="cuda")
        ker[:31, :31] = torch.rand(31, 31, dtype=torch.float64, device="cuda", generator=g)
        kerf = torch.empty(h, 2, dtype=torch.float64, device="cuda")
        rf.execute(ker.data_ptr(), kerf.data_ptr(), st)
        kerf[:, 1] *= -1.0  # uConj: the cached kernel spectrum
        crop = torch.empty(4066 * 4066, dtype=torch.float64, device="cuda")
        od = np.array([4066, 4066], dtype=np.int32)

        def convolve():
            _lib.check(_lib.lib().sstc_fft_convolve_dev(ri._h, ctypes.c_void_p(spec.data_ptr()), ctypes.c_void_p(kerf.data_ptr()),
                                                        ctypes.c_void_p(crop.data_ptr()), od.ctypes.data_as(_lib.c_i32_p),
                                                        ctypes.c_void_p(st)))
        ms_c = _ev(torch, convolve, 20)
        bytes_model = 8.0 * n + 16.0 * h + 2 * 16.0 * h  # r2c pass (read reals, write half-spectrum) + one c2c pass in place
        hr = r.cpu().numpy()
    res = {"config": "C2: RealArray 4096x4096 rfft / rifft / FFT-based 2-D correlation, 1 B200 (BASELINE configs[1])",
           "metric": "rfft GFLOP/s (2.5N*log2N)", "value": flops / ms_f / 1e6, "unit": "GFLOP/s", "ms": ms_f,
           "rifft_ms": ms_i, "correlation_31x31_fused_ms": ms_c,
           "roofline": {"bound": "hbm", "achieved": bytes_model / ms_f / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": bytes_model / ms_f / 1e6 / peak, "algorithmic_bytes": bytes_model,
                        "model": "two passes: r2c (8 B read + ~8 B written per real) and the column c2c (16 + 16 B per bin)"},
           "clocks": w.clocks()}
    # parity at FULL size against the oracle (the CPU leg below is the same call, so it costs nothing extra)
    if with_cpu:
        secs, want = _timed_cpu(lambda: oracle.rfft(d, hr))
        got = spec.cpu().numpy().reshape(-1)
        res["parity"] = {"rel_l2_vs_oracle_full_size": float(np.linalg.norm(got - want) / np.linalg.norm(want)),
                         "roundtrip_rel_l2": rt, "tolerance": 1e-12 * math.log2(n)}
        res["cpu_baseline"] = {"value": flops / secs / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port", "seconds": secs,
                               "sample": "the same 4096x4096 rfft, whole (no sampling), oracle = C restatement of "
                                         "JavaFftService.rfft / FftOps.java"}
    else:
        res["parity"] = {"roundtrip_rel_l2": rt, "tolerance": 1e-12 * math.log2(n)}
    return res


def _i32(v):
    from shared_b200 import _lib
    return _lib.i32_array(v)


def _structural(name, *args):
    from shared_b200 import _lib
    _lib.check(getattr(_lib.lib(), name)(*args))


OP = dict(RU_MUL=1, RU_EXP=4, RU_SQRT=7, RE_ADD=0, RE_MUL=2, RR_SUM=0, RR_MAX=2, RR_VAR=4, RD_SUM=0, RI_MAX=0, RA_SUM=0)


def config_c4(torch, s, peak, sampler, with_cpu):
    import oracle
    from shared_b200.array_kernel import REAL, DeviceArrayKernel
    from shared_b200.image_kernel import DeviceImageKernel
    st = torch.cuda.current_stream().cuda_stream
    dk, ik = DeviceArrayKernel(0, st), DeviceImageKernel(0, st)
    n = 8192
    N = n * n
    far = [n, 1]
    with Window(sampler) as w:
        g = torch.Generator(device="cuda").manual_seed(4)
        x = torch.rand(N, dtype=torch.float64, device="cuda", generator=g) + 0.5
        y = torch.rand(N, dtype=torch.float64, device="cuda", generator=g) + 0.5
        z = x.clone()
        row = torch.empty(n, dtype=torch.float64, device="cuda")
        idx = torch.empty(N, dtype=torch.int32, device="cuda")
        one = torch.empty(2, dtype=torch.float64, device="cuda")
        ii = torch.zeros((n + 1) * (n + 1), dtype=torch.float64, device="cuda")
        P = x.data_ptr()
        cases = [
            ("ruOp MUL", 16, lambda: dk.ruOp(OP["RU_MUL"], 1.0000001, z.data_ptr(), N)),
            ("ruOp SQRT", 16, lambda: dk.ruOp(OP["RU_SQRT"], 0, z.data_ptr(), N)),
            ("ruOp EXP", 16, lambda: dk.ruOp(OP["RU_EXP"], 0, z.data_ptr(), N)),
            ("eOp ADD", 24, lambda: dk.eOp(OP["RE_ADD"], REAL, P, y.data_ptr(), z.data_ptr(), N)),
            ("raOp SUM", 8, lambda: dk.raOp(OP["RA_SUM"], P, N, one.data_ptr())),
            ("rrOp SUM dim1", 8, lambda: dk.rrOp(OP["RR_SUM"], P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1)),
            ("rrOp SUM dim0", 8, lambda: dk.rrOp(OP["RR_SUM"], P, N, [n, n], far, row.data_ptr(), n, [1, n], [n, 1], 0)),
            ("rrOp MAX dim1", 8, lambda: dk.rrOp(OP["RR_MAX"], P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1)),
            ("rrOp MAX dim0", 8, lambda: dk.rrOp(OP["RR_MAX"], P, N, [n, n], far, row.data_ptr(), n, [1, n], [n, 1], 0)),
            ("rrOp VAR dim1", 16, lambda: dk.rrOp(OP["RR_VAR"], P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1)),
            ("rdOp SUM dim1", 16, lambda: dk.rdOp(OP["RD_SUM"], P, N, [n, n], far, z.data_ptr(), N, 1)),
            ("rdOp SUM dim0", 16, lambda: dk.rdOp(OP["RD_SUM"], P, N, [n, n], far, z.data_ptr(), N, 0)),
            ("rdOp SUM dims 0,1", 32, lambda: dk.rdOp(OP["RD_SUM"], P, N, [n, n], far, z.data_ptr(), N, 0, 1)),
            ("riOp MAX dim1", 12, lambda: dk.riOp(OP["RI_MAX"], P, N, [n, n], far, idx.data_ptr(), N, 1)),
            ("map transpose", 16, lambda: dk.map(8, [0, 0, n, 0, 0, n], P, N, [n, n], far, z.data_ptr(), N, [n, n], [1, n])),
            ("map copy", 16, lambda: dk.map(8, [0, 0, n, 0, 0, n], P, N, [n, n], far, z.data_ptr(), N, [n, n], far)),
            ("shift by (n/2, n/2) (sstc_shift_dev: ProtoArray.shift / fftShift)", 16, lambda: _structural(
                "sstc_shift_dev", 8, ctypes.c_void_p(P), ctypes.c_void_p(z.data_ptr()), _i32([n, n]), 2, 0, _i32([n // 2, n // 2]),
                ctypes.c_void_p(st))),
            ("transpose (sstc_transpose_dev)", 16, lambda: _structural(
                "sstc_transpose_dev", 8, ctypes.c_void_p(P), ctypes.c_void_p(z.data_ptr()), _i32([n, n]), 2, 0, _i32([1, 0]),
                ctypes.c_void_p(st))),
            ("ImageKernel createIntegralImage", 16, lambda: ik.createIntegralImage(
                P, N, [n, n], far, ii.data_ptr(), ii.numel(), [n + 1, n + 1], [n + 1, 1])),
        ]
        ops = []
        for name, bpe, fn in cases:
            ms = _ev(torch, fn, 10)
            ops.append({"op": name, "ms": ms, "algorithmic_bytes_per_element": bpe, "gbs": bpe * N / ms / 1e6,
                        "frac_of_measured_hbm": bpe * N / ms / 1e6 / peak})
        m = 4096
        a = torch.rand(m * m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        b = torch.rand(m * m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        c = torch.empty(m * m, dtype=torch.float64, device="cuda")
        ms_mul = _ev(torch, lambda: dk.mul(a.data_ptr(), m * m, b.data_ptr(), m * m, m, m, c.data_ptr(), m * m, False), 5)
        want = torch.matmul(a.view(m, m), b.view(m, m)).reshape(-1)  # cuBLAS: a same-box yardstick, not the oracle
        mul_err = _rel(torch, c, want)
        ms_cublas = _ev(torch, lambda: torch.matmul(a.view(m, m), b.view(m, m)), 5)
    worst = min(ops, key=lambda o: o["frac_of_measured_hbm"])
    res = {"config": "C4: RealArray 8192x8192 ArrayKernel map / dimension-reduce families + 4096^2 mul, 1 B200 "
                     "(BASELINE configs[3])",
           "metric": "mul GFLOP/s (2n^3); per-op GB/s in ops[]", "value": 2.0 * m ** 3 / ms_mul / 1e6, "unit": "GFLOP/s",
           "ms": ms_mul, "mul": {"ms": ms_mul, "gflops": 2.0 * m ** 3 / ms_mul / 1e6, "cublas_yardstick_gflops":
                                 2.0 * m ** 3 / ms_cublas / 1e6, "rel_l2_vs_cublas": mul_err,
                                 "pipe": "FP64 tensor pipe (mma.sync m8n8k4 -> SASS DMMA.8x8x4); tcgen05 has no f64 kind"},
           "ops": ops,
           "roofline": {"bound": "hbm", "kernel": "slowest streaming op: " + worst["op"], "achieved": worst["gbs"], "peak": peak,
                        "unit": "GB/s", "frac": worst["frac_of_measured_hbm"]},
           "parity": "bit-exact against the reference's own C++ at these sizes: tests/test_array_kernel_gpu.py "
                     "(test_c4_full_size_against_reference)", "clocks": w.clocks()}
    if with_cpu:
        ref = oracle.ref()
        cn = 2048  # bounded sample: 1/16 of the elements
        cx = np.random.default_rng(0).uniform(0.5, 1.5, cn * cn)
        cy, cz = cx.copy(), np.zeros(cn * cn)
        crow, cidx = np.zeros(cn), np.zeros(cn * cn, dtype=np.int32)
        cres = []
        for name, bpe, fn in [
            ("ruOp MUL", 16, lambda: ref.ruOp(1, 1.0000001, cz)),
            ("ruOp EXP", 16, lambda: ref.ruOp(4, 0, cz * 0 + 1)),
            ("eOp ADD", 24, lambda: ref.eOp(0, cx, cy, cz, False)),
            ("rrOp SUM dim1", 8, lambda: ref.rrOp(0, cx, [cn, cn], [cn, 1], crow, [cn, 1], [1, 1], 1)),
            ("rrOp SUM dim0", 8, lambda: ref.rrOp(0, cx, [cn, cn], [cn, 1], crow, [1, cn], [cn, 1], 0)),
            ("rdOp SUM dim1", 16, lambda: ref.rdOp(0, cx, [cn, cn], [cn, 1], cz, 1)),
            ("riOp MAX dim1", 12, lambda: ref.riOp(0, cx, [cn, cn], [cn, 1], cidx, 1)),
        ]:
            secs, _ = _timed_cpu(fn)
            cres.append({"op": name, "seconds": secs, "gbs": bpe * cn * cn / secs / 1e9})
        mm = 512
        ma, mb, mc = cx[:mm * mm].copy(), cx[:mm * mm].copy(), np.zeros(mm * mm)
        secs, _ = _timed_cpu(lambda: ref.mul(ma, mb, mm, mm, mc, False))
        res["cpu_baseline"] = {"value": 2.0 * mm ** 3 / secs / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "reference",
                               "sample": "mul at 512^2 (naive ijk: 4096^2 would take minutes); ops[] at 2048x2048 = 1/16 of the "
                                         "elements; oracle/_ref = the reference's own native/src/shared/*.cpp, g++ -O2, 1 thread",
                               "ops": cres}
    return res


def config_c5(torch, dist, slab_cls, dims, dev, rank, world, sampler):
    """1024^3 round trip over the ranks (N = 8): forward, inverse, round-trip error, Parseval."""
    n = dims[0] * dims[1] * dims[2]
    n_loc = n // world
    with Window(sampler) as w:
        g = torch.Generator(device=dev).manual_seed(20261019 + rank)
        x = torch.rand(n_loc, 2, dtype=torch.float64, device=dev, generator=g) - 0.5
        slab = slab_cls(dims, exchange="pipe", device=dev.index)

        def timed(fn, reps):
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return out, t.item()

        timed(lambda: slab.forward(x), 3)
        y, ms_f = timed(lambda: slab.forward(x), 5)
        spec = y.clone()
        ex = torch.tensor([(x * x).sum().item(), (spec * spec).sum().item()], dtype=torch.float64, device=dev)
        dist.all_reduce(ex)
        timed(lambda: slab.inverse(spec), 3)
        back, ms_i = timed(lambda: slab.inverse(spec), 5)
        err = torch.tensor([((back - x) ** 2).sum().item(), (x * x).sum().item()], dtype=torch.float64, device=dev)
        dist.all_reduce(err)
        torch.cuda.synchronize()
        dist.barrier()
        slab.close()
    rel = math.sqrt(err[0].item() / err[1].item())
    flops = 5.0 * n * math.log2(n)
    return {"config": "C5: ComplexArray %dx%dx%d complex128 3-D fft/ifft round trip, %d B200, exchange over NVLink "
                      "(BASELINE configs[4])" % (dims[0], dims[1], dims[2], world),
            "metric": "forward GFLOP/s (5N*log2N)", "value": flops / ms_f / 1e6, "unit": "GFLOP/s", "ms": ms_f,
            "inverse_ms": ms_i, "roundtrip_ms": ms_f + ms_i, "bytes_per_gpu": 16 * n_loc,
            "parity": {"roundtrip_rel_l2": rel, "parseval_ratio_minus_1": ex[1].item() / (n * ex[0].item()) - 1.0,
                       "tolerance": 1e-12 * math.log2(n), "ok": rel < 1e-12 * math.log2(n)},
            "roofline": None, "clocks": w.clocks() if rank == 0 else None}


def run_single_gpu(torch, s, peak, sampler, with_cpu, only=None):
    out = []
    for name, fn in (("c1", config_c1), ("c1b", config_c1b), ("c2", config_c2), ("c4", config_c4)):
        if only and name not in only:
            continue
        try:
            out.append(fn(torch, s, peak, sampler, with_cpu))
        except Exception as e:  # noqa: BLE001 -- an extra config must never take the headline line down with it
            out.append({"config": name, "error": "%s: %s" % (type(e).__name__, e)})
        torch.cuda.empty_cache()
    return out
