#!/usr/bin/env python3
"""bench.py -- the north-star measurement: complex128 3-D FFT at 512^3 on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward c2c transform of a 512x512x512 complex128 array (BASELINE.json configs[2]) through
the CUDA provider, input resident in HBM. N > 1: the array is slab-sharded over the ranks (strong scaling:
the total work is fixed) and the step includes the slab exchange over NVLink. Rank 0 prints ONE JSON line.

Keys beyond the base contract:
  roofline      the dominant kernel, fft_pow2_kernel<512, 8, STRIDED> (it runs the y and the x pass: 2 of the 3
                launches, ~70 % of the step): algorithmic bytes per launch (16 B read + 16 B written per point,
                DESIGN.md) over its average launch duration, timed with CUDA events between the launches of
                back-to-back steps, against MEASURED_PEAKS.json's HBM copy bandwidth; `passes` lists all three
                launches and `slowest_pass` the worst of them
  cpu_baseline  the CPU oracle (C restatement of the reference's single-threaded FftOps.java) timed on this
                box's host on a bounded sample (N = 1 only); `all_cores` = that many independent transforms at once
  e2e           the same transform through the host-tier C ABI call a JNI caller makes (sstc_fft) from pinned host
                arrays, H2D + kernels + D2H inside the timed region. Every output depends on every input, so ONE
                synchronous call cannot take less than H2D + D2H (2 x 39 ms at this pool's 55 GB/s): `single_caller`
                reports that latency. `value` is the throughput of TWO caller threads, each making the same
                synchronous call (the SPI is thread-safe; the host tier gives every call its own slot), so that one
                transform's D2H overlaps the other's H2D on the full-duplex link. `pageable` repeats both with plain
                (unpinned) arrays, which is what a JVM heap array is. The reference arm is the same kind of client with one
                caller thread per host core (the reference's FFT is single-threaded; callers are what can use the cores).
  parity        (N > 1) the step that is timed -- the recorded CUDA-graph replay -- checked outside the timed region
                against the 1-GPU plan on the gathered input, plus the forward -> inverse round trip
  extra_configs BASELINE.json's other configs (C1, the reference's own benchmark loop, C2, C4; C5 at N = 8), each with
                its own clocks, roofline, parity and cpu_baseline (bench_extra.py)

`--impl reference` times the reference's own CPU algorithm (the oracle; there is no JVM or FFTW in the image)
on a bounded sample per step, from one caller thread per host core (all the host threads the single-threaded reference
can use), and prints the same line with "impl": "reference".
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (512, 512, 512)
METRIC = "complex128 3-D FFT GFLOP/s (5N*log2N) at 512^3"
SEED = 20261019


def fft_flops(n_points):
    return 5.0 * n_points * math.log2(n_points)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,enforced.power.limit")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8] or \
                 [r for (_, r) in self.rows if len(r) >= 8]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in inside)
        reasons = set()
        for r in inside:
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                              ("sw_power_cap", 7)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][2]),
               "power_w_max": max(float(r[3]) for r in inside), "reasons": sorted(reasons), "samples": len(inside)}
        try:
            out["power_limit_w"] = float(inside[0][8])  # the limit sw_power_cap refers to
        except (IndexError, ValueError):
            pass
        if "sw_power_cap" in reasons:
            out["note"] = ("sustained FP64 FFT passes draw the board's power limit; the cap lowers the SM clock after ~0.1 s of "
                           "load (profiles/r2_power_cap_vs_steps.json): --steps 20 measures the burst figure, longer runs the "
                           "sustained one")
        return out


# --------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (single-threaded, like every provider of the reference)
# --------------------------------------------------------------------------------------------------------

def cpu_model():
    """Model name of the host CPU the baseline ran on (SURVEY.md 8d asks for it next to the core count)."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def base_config(n_gpus, extra=None):
    """The workload both arms are labelled with (the reference arm must carry OUR config: same metric, same workload)."""
    cfg = {"workload": "ComplexArray 512x512x512 complex128 3-D fft (forward)", "dims": list(DIMS),
           "parallelism": "1 GPU" if n_gpus == 1 else "slab x%d over d0, exchange over NVLink" % n_gpus,
           "l2": "inputs larger than L2 (2 GiB in + 2 GiB out vs 126 MB), no flush needed",
           "client": "device-resident steps for `value`; caller threads making synchronous provider calls for `e2e` (ours: two, "
                     "the PCIe link is full then; reference arm: one per host core)"}
    if extra:
        cfg.update(extra)
    return cfg


def cpu_fft_seconds(dims, reps, callers=1):
    """Best wall time of `callers` concurrent forward transforms of `dims` (each on its own array, one thread each: the
    reference's FFT is single-threaded; ctypes releases the GIL)."""
    import numpy as np
    import oracle
    rng = np.random.default_rng(SEED)
    n = int(np.prod(dims))
    xs = [rng.uniform(-0.5, 0.5, 2 * n) for _ in range(callers)]
    best = float("inf")
    for _ in range(reps):
        th = [threading.Thread(target=oracle.fft, args=(list(dims), x)) for x in xs]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(sample=(256, 256, 256)):
    """Bounded sample of the same workload: one forward c2c of a `sample` cube after a 64^3 warm-up; then as many
    independent transforms at once as the host has cores (bounded by memory: 1 GiB per 256^3 transform)."""
    cpu_fft_seconds((64, 64, 64), 1)
    secs = cpu_fft_seconds(sample, 1)
    n = sample[0] * sample[1] * sample[2]
    cores = os.cpu_count() or 1
    many = max(1, min(cores, 16))
    secs_many = cpu_fft_seconds(sample, 1, callers=many)
    return {"value": fft_flops(n) / secs / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "host_cores": cores, "cpu_model": cpu_model(), "seconds": secs,
            "all_cores": {"value": many * fft_flops(n) / secs_many / 1e9, "unit": "GFLOP/s", "cores": many,
                          "note": "%d independent transforms at once, one thread each (the reference has no threaded FFT)" % many},
            "sample": "forward c2c %dx%dx%d (1/%d of the 512^3 points), oracle/fftops_oracle.c = C restatement of "
                      "FftOps.java, gcc -O2, 1 thread (the reference path is single-threaded)" %
                      (sample + ((DIMS[0] * DIMS[1] * DIMS[2]) // n,))}


def reference_callers():
    """The reference's FFT is single-threaded, and its SPI may be called from any thread (Threading): the
    arm that uses "all the host threads it can" is one caller thread per host core, each making synchronous calls on its
    own arrays -- the same kind of client as our e2e leg, which stops at two callers only because the PCIe link is full
    then. Bounded by memory (about 1 GiB per 256^3 caller) and by SSTC_REFERENCE_CALLERS."""
    env = os.environ.get("SSTC_REFERENCE_CALLERS")
    if env:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    return max(1, min(cores, 32))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # 256^3 costs ~2.6 s per transform on this pool's hosts; 512^3 would be ~25 s and 8 GiB per caller, so the step is a
    # 1/8 sample of the workload -- GFLOP/s normalises the size (and the smaller cube fits the CPU's caches better, so the
    # sample flatters the reference if anything)
    sample = (256, 256, 256)
    import numpy as np
    import oracle
    n = sample[0] * sample[1] * sample[2]
    callers = reference_callers()
    rng = np.random.default_rng(SEED)
    xs = [rng.uniform(-0.5, 0.5, 2 * n) for _ in range(callers)]

    def caller(x, steps):
        for _ in range(steps):
            oracle.fft(list(sample), x)

    def run(steps, k):
        th = [threading.Thread(target=caller, args=(x, steps)) for x in xs[:k]]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    # K steps in total, shared by the callers (rounded up to a whole number per caller)
    per_caller = max(1, -(-args.steps // callers))
    if args.warmup:
        run(max(1, -(-args.warmup // callers)), callers)
    wall = run(per_caller, callers)
    transforms = per_caller * callers
    secs = wall / transforms
    value = fft_flops(n) / secs / 1e9
    single = run(1, 1)
    two = run(1, min(2, callers)) / min(2, callers)
    desc = ("forward c2c %dx%dx%d per step (1/8 of the 512^3 points), oracle/fftops_oracle.c (C restatement of the "
            "reference's FftOps.java; no JVM/FFTW in the image); %d caller threads (one per host core), one single-threaded "
            "transform each -- the reference has no threaded FFT" % (sample + (callers,)))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args.gpus), "transforms_timed": transforms,
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": callers, "kind": "port", "sample": desc,
                         "host_cores": os.cpu_count(), "cpu_model": cpu_model(),
                         "single_caller": {"value": fft_flops(n) / single / 1e9, "unit": "GFLOP/s", "cores": 1,
                                           "ms_per_step": single * 1e3},
                         "two_callers": {"value": fft_flops(n) / two / 1e9, "unit": "GFLOP/s", "cores": min(2, callers),
                                         "ms_per_step": two * 1e3,
                                         "note": "the client our e2e leg uses (its link is full with two)"}},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "callers": callers},
    }))


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------

def event_ms(fn, reps, torch):
    """Average milliseconds of fn() over `reps` back-to-back calls, CUDA events on the current stream."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def pinned_array(n_doubles):
    import numpy as np
    from shared_b200 import _lib
    p = ctypes.c_void_p()
    _lib.check(_lib.lib().sstc_host_alloc(ctypes.byref(p), n_doubles * 8))
    buf = (ctypes.c_double * n_doubles).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64), p


def run_ours(args):
    import torch
    import torch.distributed as dist
    import shared_b200
    from shared_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d needs WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.init(local)
    lib = _lib.lib()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = DIMS[0] * DIMS[1] * DIMS[2]
    flops = fft_flops(n)
    peak, peak_src = measured_peaks()
    K, W = args.steps, args.warmup
    stream = torch.cuda.current_stream().cuda_stream

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    gen = torch.Generator(device=dev).manual_seed(SEED + rank)
    n_local = n // world
    x = torch.rand(n_local, 2, dtype=torch.float64, device=dev, generator=gen) - 0.5

    passes = []
    if world == 1:
        plan = shared_b200.FftPlan(shared_b200.FORWARD, DIMS, device=local)
        if args.transposed_passes is not None:
            plan.set_hint("transposed_passes", args.transposed_passes)
        y = torch.empty_like(x)

        def step():
            plan.execute(x.data_ptr(), y.data_ptr(), stream)

        launches_per_step = plan.launches
        schedule = ("z in->out, y out->workspace (transposed store), x workspace->out" if plan.workspace_bytes
                    else "one pass per axis, in place after the first")
    else:
        from shared_b200.sharded import SlabFft3d
        slab = SlabFft3d(DIMS, exchange=args.exchange, device=local, groups=args.groups,
                         exchange_ctas=args.exchange_ctas, chunks=args.chunks, bulk_store=args.bulk_store,
                         graph=0 if args.no_graph else None)
        if args.exchange == "pipe":
            slab.capture(x)  # two eager steps; the C side replays recorded CUDA graphs from the third call on

        def step():
            slab.forward(x)

        ahead = bool(args.begin_ahead) and args.exchange == "pipe"
        launches_per_step = None  # counted below from the library's launch counter
        schedule = "%s exchange: y pass, z pass storing lines into the peers' slabs in groups, barrier, x pass per group" % args.exchange

    def timed_loop(step_fn):
        """W untimed + K timed steps, barrier + synchronise on both sides, CUDA events, max over ranks."""
        for _ in range(W):
            step_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = lib.sstc_launch_count()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step_fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.time()
        ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        return ms_total.item() / K, lib.sstc_launch_count() - l0, t0, t1

    ms_step, launches, t_wall0, t_wall1 = timed_loop(step)
    one_at_a_time = None
    if world > 1 and ahead:
        # The steps above ran one at a time: a transform starts when its predecessor has finished. A STREAM of independent
        # transforms -- which is what K back-to-back steps on fresh inputs are -- lets the local y pass of the next transform
        # run while the current one is exchanging (sstc_slab_forward_begin): every timed step below issues the y pass of a
        # later transform and finishes an earlier one, so K steps still complete K whole transforms. This is the headline;
        # the one-at-a-time figure stays in the line.
        one_at_a_time = {"ms_per_step": ms_step, "value": flops / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s"}
        slab.forward_begin(x)

        def step_ahead():
            slab.forward_begin(x)
            slab.forward(x)

        ms_step, launches, t_wall0, t_wall1 = timed_loop(step_ahead)
        slab.forward(x)  # finishes the transform begun by the last timed step
        torch.cuda.synchronize()
        schedule += "; y pass of transform k+1 issued beside the exchange of transform k (forward_begin)"

    # ---- roofline: each axis pass timed alone (device-resident, events on the launching stream) ----------
    roofline = None
    if world == 1:
        reps = max(5, min(K, 20))
        d0, d1, d2 = DIMS
        bytes_per_launch = 32.0 * n  # 16 B read + 16 B written per complex point
        if plan.workspace_bytes:  # transposed schedule: time the passes exactly as the plan issues them
            w = torch.empty_like(x)
            pl = d1 * d2
            specs = [("z (contiguous axis)", lambda: shared_b200.fft_axis(x.data_ptr(), y.data_ptr(), d0 * d1, d2, 1,
                                                                          False, 1.0, stream)),
                     ("y (stride d2, transposed store)", lambda: shared_b200.fft_axis_strided(
                         y.data_ptr(), w.data_ptr(), d0, d1, d2, pl, d2, d2, d0 * d2, False, 1.0, stream)),
                     ("x (stride d2 after transpose, natural store)", lambda: shared_b200.fft_axis_strided(
                         w.data_ptr(), y.data_ptr(), d1, d0, d2, d0 * d2, d2, d2, pl, False, 1.0, stream))]
        else:
            specs = [("z (contiguous axis)", lambda: shared_b200.fft_axis(x.data_ptr(), y.data_ptr(), d0 * d1, d2, 1,
                                                                          False, 1.0, stream)),
                     ("y (stride d2)", lambda: shared_b200.fft_axis(y.data_ptr(), y.data_ptr(), d0, d1, d2, False,
                                                                    1.0, stream)),
                     ("x (stride d1*d2)", lambda: shared_b200.fft_axis(y.data_ptr(), y.data_ptr(), 1, d0, d1 * d2,
                                                                       False, 1.0, stream))]
        # per-pass times INSIDE back-to-back steps (events between the launches), so that each pass starts from the
        # cache / power state its predecessor leaves -- the state it has in the timed region above
        for _, f in specs:
            f()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(specs) + 1)] for _ in range(reps)]
        torch.cuda.synchronize()
        for r in range(reps):
            evs[r][0].record()
            for i, (_, f) in enumerate(specs):
                f()
                evs[r][i + 1].record()
        torch.cuda.synchronize()
        for i, (name, _) in enumerate(specs):
            ms = sum(evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(reps)) / reps
            passes.append({"pass": name, "ms": ms, "gbs": bytes_per_launch / ms / 1e6})
        # the dominant kernel: fft_pow2_kernel<512, 8, STRIDED> runs the y and the x pass (2 of the 3 launches and
        # ~70 % of the step); its average launch duration is the mean of those two launches
        strided = passes[1:]
        dom_ms = sum(p["ms"] for p in strided) / len(strided)
        dom_gbs = bytes_per_launch / dom_ms / 1e6
        worst = max(passes, key=lambda p: p["ms"])
        traffic = None  # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
        try:
            with open(os.path.join(ROOT, "profiles", "r2_fft512_traffic.json")) as f:
                per = json.load(f)["dram_bytes_read_plus_write_per_launch"]
                traffic = (per["y"] + per["x"]) / 2
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "fft_pow2_kernel<512, 8, STRIDED> (y and x passes, 2 launches per step)",
                    "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "avg_launch_ms": dom_ms, "kernel_share_of_step": sum(p["ms"] for p in strided) / sum(p["ms"] for p in passes),
                    "slowest_pass": {"pass": worst["pass"], "gbs": worst["gbs"], "frac": worst["gbs"] / peak},
                    "step_model_gbs": 3 * bytes_per_launch / ms_step / 1e6,
                    "step_frac": 3 * bytes_per_launch / ms_step / 1e6 / peak, "passes": passes,
                    "timing": "CUDA events between the launches of back-to-back steps, %d steps" % reps}
    elif world > 1:
        # N > 1: the local x pass (fft_pow2_ext_kernel<512, 8, STRIDED> on the received slab) timed alone against the
        # HBM roofline, and the exchange as NVLink bytes per GPU over the step (the step is exchange-bound from 4 GPUs on)
        reps = max(5, min(K, 20))
        out = slab.forward(x)
        d0, d1, d2 = DIMS
        cols = (d1 // world) * d2

        def x_pass():
            slab.engine.axis(out, out, 1, d0, cols)

        x_pass()
        ms = event_ms(x_pass, reps, torch)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bytes_per_launch = 32.0 * n_local
        gbs = bytes_per_launch / t.item() / 1e6
        sent = 16.0 * n_local * (world - 1) / world  # bytes each GPU stores into its peers per step
        roofline = {"bound": "hbm", "kernel": "fft_pow2_ext_kernel<512, 8, STRIDED> (local x pass on the received slab, "
                    "timed alone, max over ranks)", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                    "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "avg_launch_ms": t.item(),
                    "exchange": {"nvlink_bytes_sent_per_gpu_per_step": sent,
                                 "gbs_per_gpu_over_whole_step": sent / ms_step / 1e6,
                                 "note": "the z pass stores whole lines into the peers' slabs (sstc_fft_lines_scatter_dev); "
                                         "measured all-to-all rate of this box: ~590 GB/s per GPU (DESIGN.md 3.4)"}}
    t_load1 = time.time()

    # ---- parity of the step that was timed (N > 1): outside the timed region --------------------------------------
    parity = None
    if world > 1:
        parity = sharded_parity(torch, dist, shared_b200, slab, x, dev, rank, world, local, stream, ahead)

    # ---- e2e: host arrays, H2D + kernels + D2H inside the timed region -------------------------------------------
    e2e_steps = max(2, min(K, 6))
    if args.no_e2e:
        e2e = None  # schedule sweeps only: a line without e2e is not a bench line
    elif world == 1:
        e2e = e2e_single_gpu(shared_b200, lib, local, n, flops, e2e_steps)
    else:
        e2e = e2e_sharded(torch, dist, lib, slab, x, dev, n, n_local, flops, stream, e2e_steps)

    # ---- the other BASELINE configs -------------------------------------------------------------------------------
    extra = None
    t_extra0 = time.time()
    if not args.no_extra:
        import bench_extra
        if world == 1:
            del x, y
            torch.cuda.empty_cache()
            extra = bench_extra.run_single_gpu(torch, shared_b200, peak, sampler, not args.no_cpu,
                                               only=args.extra.split(",") if args.extra else None)
        elif world == 8:
            from shared_b200.sharded import SlabFft3d
            slab.close()
            del x
            torch.cuda.empty_cache()
            try:
                extra = [bench_extra.config_c5(torch, dist, SlabFft3d, [1024, 1024, 1024], dev, rank, world, sampler)]
            except Exception as e:  # noqa: BLE001
                extra = [{"config": "C5", "error": "%s: %s" % (type(e).__name__, e)}]
    if sampler:
        sampler.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": flops / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(world), "schedule": schedule,
            "hbm_model_gbs": 3 * 32.0 * n / ms_step / 1e6,
            "gpu_launches": int(launches), "launches_per_step": launches_per_step,
            "e2e": e2e,
            "clocks": sampler.summary(t_wall0, t_load1) if sampler else None,
        }
        if roofline:
            line["roofline"] = roofline
        if parity:
            line["parity"] = parity
        if one_at_a_time:
            line["one_at_a_time"] = one_at_a_time
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        if extra is not None:
            line["extra_configs"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sharded_parity(torch, dist, shared_b200, slab, x, dev, rank, world, local, stream, ahead=False):
    """The timed step is the recorded CUDA-graph replay of the pipelined schedule: run it once more on the bench input,
    gather input and output on rank 0, and compare with the 1-GPU plan on the same data; then the inverse."""
    import math
    n = DIMS[0] * DIMS[1] * DIMS[2]
    replayed = slab.graph_count() if hasattr(slab, "graph_count") else None
    if ahead:
        slab.forward_begin(x)  # the variant the headline loop timed: y pass issued ahead, exchange + x passes from a graph
    out = slab.forward(x)
    spec = out.clone()
    back = slab.inverse(spec)
    rt = torch.tensor([((back - x) ** 2).sum().item(), (x * x).sum().item()], dtype=torch.float64, device=dev)
    dist.all_reduce(rt)
    roundtrip = math.sqrt(rt[0].item() / rt[1].item())
    xs = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
    ys = [torch.empty_like(spec) for _ in range(world)] if rank == 0 else None
    dist.gather(x, xs, dst=0)
    dist.gather(spec, ys, dst=0)
    res = None
    if rank == 0:
        from shared_b200.sharded import gather_transposed
        full = torch.cat(xs)
        del xs
        got = gather_transposed(ys, DIMS).reshape(n, 2)
        del ys
        ref = torch.empty_like(full)
        plan = shared_b200.FftPlan(shared_b200.FORWARD, DIMS, device=local)
        plan.execute(full.data_ptr(), ref.data_ptr(), stream)
        torch.cuda.synchronize()
        err = (torch.linalg.norm(got - ref) / torch.linalg.norm(ref)).item()
        res = {"what": "the CUDA-graph replay of the pipelined slab schedule (the timed step) against the 1-GPU plan on the "
                       "gathered bench input; forward -> inverse round trip over all ranks",
               "rel_l2_vs_1gpu": err, "bit_identical": bool(torch.equal(got, ref)),
               "roundtrip_rel_l2": roundtrip, "tolerance": 1e-12 * math.log2(n),
               "ok": bool(err < 1e-12 * math.log2(n) and roundtrip < 1e-12 * math.log2(n)),
               "recorded_graphs": replayed,
               "note": "not bit-identical by construction: the pipelined schedule transforms y before z, the 1-GPU plan z "
                       "before y (same butterflies, different rounding order)"}
        del full, got, ref
        plan.close()
        torch.cuda.empty_cache()
    dist.barrier()
    return res


def e2e_single_gpu(shared_b200, lib, local, n, flops, steps):
    """sstc_fft through CudaFftService.fft: single caller and two callers, pinned and pageable arrays."""
    import numpy as np
    svc = shared_b200.CudaFftService(local)

    def run(arrays, callers, reps):
        def caller(i):
            for _ in range(reps):
                svc.fft(DIMS, arrays[i][0], arrays[i][1])
        th = [threading.Thread(target=caller, args=(i,)) for i in range(callers)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return (time.perf_counter() - t0) / (reps * callers)

    def check(arrays):
        for h_in, h_out in arrays:
            assert abs(float(h_out[0]) - 0.25 * n) < 1e-6 * n and abs(float(h_out[1])) < 1e-6 * n, (h_out[0], h_out[1])
            assert float(np.abs(h_out[2:4096]).max()) < 1e-6 * n

    res = {}
    pins = []
    for _ in range(2):
        h_in, p_in = pinned_array(2 * n)
        h_out, p_out = pinned_array(2 * n)
        h_in[0::2] = 0.25
        h_in[1::2] = 0.0
        pins.append((h_in, h_out, p_in, p_out))
    arrays = [(a, b) for (a, b, _, _) in pins]
    run(arrays, 2, 1)  # warm-up: plans, slots, device buffers
    single = run(arrays, 1, steps)
    reps2 = max(12, 2 * steps)  # per caller: the first upload and the last download have no partner (38 ms of ramp)
    double = run(arrays, 2, reps2)
    check(arrays)
    for (_, _, p_in, p_out) in pins:
        lib.sstc_host_free(p_in)
        lib.sstc_host_free(p_out)
    del pins, arrays
    pages = []
    for _ in range(2):
        h_in = np.empty(2 * n)
        h_in[0::2] = 0.25
        h_in[1::2] = 0.0
        pages.append((h_in, np.zeros(2 * n)))
    run(pages, 2, 1)  # warm-up: staging rings, copy threads
    p_single = run(pages, 1, max(2, steps // 2))
    p_double = run(pages, 2, max(3, steps // 2 + 1))
    check(pages)
    h2d = d2h = 16 * n
    res = {"value": flops / double / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": double * 1e3, "steps": steps, "callers": 2, "transforms_timed": 2 * reps2,
           "host_memory": "pinned",
           "api": "CudaFftService.fft -> sstc_fft (host tier), two caller threads, each call synchronous",
           "single_caller": {"ms_per_step": single * 1e3, "value": flops / single / 1e9, "unit": "GFLOP/s",
                             "bound": "one transform cannot finish before H2D + D2H: every output depends on every input"},
           "pageable": {"two_callers_ms_per_step": p_double * 1e3, "two_callers_value": flops / p_double / 1e9,
                        "single_caller_ms_per_step": p_single * 1e3, "single_caller_value": flops / p_single / 1e9,
                        "unit": "GFLOP/s", "note": "plain numpy arrays = what GetPrimitiveArrayCritical hands the JNI glue; "
                        "staged through the pinned ring by the copy threads"},
           "host": host_topology()}
    return res


def host_topology():
    """NUMA nodes of the host and of the visible GPUs (sysfs), so that a flattening e2e curve can be read against the box."""
    import glob
    import os
    out = {"numa_nodes": len(glob.glob("/sys/devices/system/node/node[0-9]*")), "cpus": os.cpu_count()}
    try:
        out["cpus_allowed"] = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    nodes = {}
    for d in glob.glob("/sys/bus/pci/devices/*"):
        try:
            if open(d + "/vendor").read().strip() == "0x10de" and open(d + "/class").read().startswith("0x0302"):
                nodes[os.path.basename(d)] = int(open(d + "/numa_node").read())
        except (OSError, ValueError):
            pass
    out["gpu_numa_node"] = nodes
    return out


def e2e_sharded(torch, dist, lib, slab, x, dev, n, n_local, flops, stream, steps):
    """Per rank: H2D of its slab from pinned memory, the sharded forward, D2H of its result slab; max over ranks."""
    h_in, p_in = pinned_array(2 * n_local)
    h_out, p_out = pinned_array(2 * n_local)
    h_in[:] = 0.25
    s = ctypes.c_void_p(stream)

    def h2d():
        lib.sstc_h2d(ctypes.c_void_p(x.data_ptr()), p_in, 16 * n_local, s)

    def e2e_step():
        h2d()
        out = slab.forward(x)
        lib.sstc_d2h(p_out, ctypes.c_void_p(out.data_ptr()), 16 * n_local, s)
        torch.cuda.synchronize()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
            torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def d2h():
        lib.sstc_d2h(p_out, ctypes.c_void_p(x.data_ptr()), 16 * n_local, s)

    # A stream of independent transforms, the N > 1 twin of the two-caller leg at N = 1: the upload of transform i + 1 and
    # the download of transform i share the link in both directions. Two input slabs; the results alternate between the
    # slab context's two receive buffers, and a result must have left before the next call (include/sst_cuda.h, sstc_slab).
    x2 = torch.empty_like(x)
    xs = (x, x2)
    h_out2, p_out2 = pinned_array(2 * n_local)
    outs = (p_out, p_out2)
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    c_in, c_out = ctypes.c_void_p(s_in.cuda_stream), ctypes.c_void_p(s_out.cuda_stream)

    def stream_of(reps):
        up = [torch.cuda.Event(), torch.cuda.Event()]
        used = [None, None]  # forward that last read xs[b]
        down = None

        def upload(i):
            b = i % 2
            if used[b] is not None:
                s_in.wait_event(used[b])
            lib.sstc_h2d(ctypes.c_void_p(xs[b].data_ptr()), p_in, 16 * n_local, c_in)
            up[b].record(s_in)

        s_in.wait_stream(main)
        s_out.wait_stream(main)
        upload(0)
        for i in range(reps):
            b = i % 2
            main.wait_event(up[b])
            if down is not None:
                main.wait_event(down)  # the previous result has left its receive buffer
            out = slab.forward(xs[b])
            done = torch.cuda.Event()
            done.record(main)
            used[b] = done
            if i + 1 < reps:
                upload(i + 1)
            s_out.wait_event(done)
            lib.sstc_d2h(outs[b], ctypes.c_void_p(out.data_ptr()), 16 * n_local, c_out)
            down = torch.cuda.Event()
            down.record(s_out)
        main.wait_stream(s_in)
        main.wait_stream(s_out)

    def timed_stream(reps):
        stream_of(6)  # both input pointers reach their graphs (recorded on the third call with the same pointer)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        stream_of(reps)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    e2e_s = timed(e2e_step, steps)
    h2d_s = timed(h2d, steps)  # all ranks copying at once: what the host's PCIe fabric gives N links together
    d2h_s = timed(d2h, steps)  # and what the host's memory takes from them
    reps_stream = max(24, 4 * steps)  # the first upload and the last download have no partner
    stream_s = timed_stream(reps_stream)
    expect = 0.25 * n
    got = (float(h_out[0]), float(h_out2[0]))
    assert dist.get_rank() != 0 or all(abs(g - expect) < 1e-6 * n for g in got), (got, expect)
    world = dist.get_world_size()
    lib.sstc_host_free(p_in)
    lib.sstc_host_free(p_out)
    lib.sstc_host_free(p_out2)
    return {"value": flops / stream_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n,
            "ms_per_step": stream_s * 1e3, "steps": steps, "transforms_timed": reps_stream, "host_memory": "pinned",
            "api": "per rank: sstc_h2d + SlabFft3d.forward (sstc_slab_forward) + sstc_d2h; a stream of independent transforms, "
                   "the upload of the next one beside the download of the last one (two input slabs, two result buffers)",
            "single_caller": {"ms_per_step": e2e_s * 1e3, "value": flops / e2e_s / 1e9, "unit": "GFLOP/s",
                              "note": "one transform at a time: upload, transform, download, synchronize"},
            "h2d_alone": {"ms": h2d_s * 1e3, "gbs_per_gpu": 16 * n_local / h2d_s / 1e9,
                          "gbs_all_gpus": 16 * n / h2d_s / 1e9,
                          "note": "all %d ranks copying their slab at the same time" % world},
            "d2h_alone": {"ms": d2h_s * 1e3, "gbs_per_gpu": 16 * n_local / d2h_s / 1e9,
                          "gbs_all_gpus": 16 * n / d2h_s / 1e9},
            "host": host_topology()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--exchange", choices=["pipe", "p2p", "collective"], default="pipe")
    ap.add_argument("--exchange-ctas", type=int, default=None, dest="exchange_ctas",
                    help="cap on the exchange pass's grid (0 = one CTA per tile)")
    ap.add_argument("--groups", type=int, default=None, help="pipeline groups of the pipe exchange")
    ap.add_argument("--transposed-passes", type=int, default=None, dest="transposed_passes",
                    help="plan hint: -1 auto (default), 0 never, 1 always")
    ap.add_argument("--chunks", type=int, default=None, help="pieces of the local pass in front of the exchange")
    ap.add_argument("--bulk-store", type=int, default=None, dest="bulk_store",
                    help="1: the exchange pass sends every line as one cp.async.bulk copy")
    ap.add_argument("--begin-ahead", type=int, default=1, dest="begin_ahead",
                    help="N > 1: 1 (default) = also time the steps as a stream of independent transforms, every transform's "
                         "local y pass issued one step ahead; 0 = one at a time only")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", dest="no_extra", help="skip extra_configs")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e", help="skip the e2e leg (schedule sweeps only)")
    ap.add_argument("--extra", default=None, help="comma-separated subset of extra configs (c1,c1b,c2,c4)")
    ap.add_argument("--no-graph", action="store_true", dest="no_graph", help="issue the pipe schedule eagerly")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
