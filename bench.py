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
                box's host on a bounded sample (N = 1 only)
  e2e           the same transform through the host-tier C ABI call a JNI caller makes (sstc_fft): pinned host
                arrays in, H2D + kernels + D2H inside the timed region

`--impl reference` times the reference's own CPU algorithm (the oracle; there is no JVM or FFTW in the image)
on a bounded sample per step and prints the same line with "impl": "reference".
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
              "clocks_event_reasons.sw_power_cap")

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
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][2]),
                "power_w_max": max(float(r[3]) for r in inside), "reasons": sorted(reasons), "samples": len(inside)}


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


def cpu_fft_seconds(dims, reps):
    import numpy as np
    import oracle
    rng = np.random.default_rng(SEED)
    n = int(np.prod(dims))
    x = rng.uniform(-0.5, 0.5, 2 * n)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.fft(list(dims), x)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(sample=(256, 256, 256)):
    """Bounded sample of the same workload: one forward c2c of a `sample` cube after a 64^3 warm-up."""
    cpu_fft_seconds((64, 64, 64), 1)
    secs = cpu_fft_seconds(sample, 1)
    n = sample[0] * sample[1] * sample[2]
    return {"value": fft_flops(n) / secs / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "host_cores": os.cpu_count(), "cpu_model": cpu_model(), "seconds": secs,
            "sample": "forward c2c %dx%dx%d (1/%d of the 512^3 points), oracle/fftops_oracle.c = C restatement of "
                      "FftOps.java, gcc -O2, 1 thread (the reference path is single-threaded)" %
                      (sample + ((DIMS[0] * DIMS[1] * DIMS[2]) // n,))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # keep (K + W) steps within a few minutes: 256^3 costs ~7 s per step, 128^3 ~0.8 s
    sample = (256, 256, 256) if (args.steps + args.warmup) <= 12 else (128, 128, 128)
    import numpy as np
    import oracle
    n = sample[0] * sample[1] * sample[2]
    x = np.random.default_rng(SEED).uniform(-0.5, 0.5, 2 * n)
    for _ in range(args.warmup):
        oracle.fft(list(sample), x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.fft(list(sample), x)
    secs = (time.perf_counter() - t0) / max(1, args.steps)
    value = fft_flops(n) / secs / 1e9
    desc = ("forward c2c %dx%dx%d per step, oracle/fftops_oracle.c (C restatement of the reference's FftOps.java; "
            "no JVM/FFTW in the image), 1 thread: the reference path is single-threaded" % sample)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ComplexArray 512x512x512 complex128 3-D fft (forward)", "sample": desc},
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": desc,
                         "host_cores": os.cpu_count(), "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        parallelism = "1 GPU"
    else:
        from shared_b200.sharded import SlabFft3d
        slab = SlabFft3d(DIMS, exchange=args.exchange, device=local, groups=args.groups,
                         exchange_ctas=args.exchange_ctas)
        if args.exchange == "pipe" and not args.no_graph:
            slab.capture(x)

        def step():
            slab.forward(x)

        launches_per_step = None  # counted below from the library's launch counter
        parallelism = "slab x%d over d0, %s exchange over NVLink" % (world, args.exchange)

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.sstc_launch_count()
    replayed0 = slab.replayed_kernels if world > 1 else 0
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = lib.sstc_launch_count() - l0
    if world > 1:
        launches += slab.replayed_kernels - replayed0  # kernels inside the replayed CUDA graphs
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / K

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
            with open(os.path.join(ROOT, "profiles", "r1_fft512_traffic.json")) as f:
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

    # ---- e2e: host arrays in pinned memory, H2D + kernels + D2H inside the timed region ------------------
    e2e_steps = max(1, min(K, 5))
    if world == 1:
        h_in, p_in = pinned_array(2 * n)
        h_out, p_out = pinned_array(2 * n)
        h_in[:] = 0.25
        svc = shared_b200.CudaFftService(local)
        svc.fft(DIMS, h_in, h_out)  # warm-up (plan + staging buffers)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            svc.fft(DIMS, h_in, h_out)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        h2d = d2h = 16 * n
        checksum = float(h_out[0])
        assert abs(checksum - 0.25 * n) < 1e-6 * n, checksum
    else:
        h_in, p_in = pinned_array(2 * n_local)
        h_out, p_out = pinned_array(2 * n_local)
        h_in[:] = 0.25

        def e2e_step():
            lib.sstc_h2d(ctypes.c_void_p(x.data_ptr()), p_in, 16 * n_local, ctypes.c_void_p(stream))
            out = slab.forward(x)
            lib.sstc_d2h(p_out, ctypes.c_void_p(out.data_ptr()), 16 * n_local, ctypes.c_void_p(stream))
            torch.cuda.synchronize()

        e2e_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
        h2d = d2h = 16 * n  # summed over ranks
    if sampler:
        sampler.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": flops / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ComplexArray 512x512x512 complex128 3-D fft (forward), device-resident",
                       "dims": list(DIMS), "parallelism": parallelism,
                       "l2": "inputs larger than L2 (2 GiB in + 2 GiB out vs 126 MB), no flush needed",
                       "schedule": ("z in->out, y out->workspace (transposed store), x workspace->out"
                                    if world == 1 and plan.workspace_bytes else "one pass per axis, in place after the first")},
            "hbm_model_gbs": 3 * 32.0 * n / ms_step / 1e6,
            "gpu_launches": int(launches), "launches_per_step": launches_per_step,
            "e2e": {"value": flops / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "api": "CudaFftService.fft -> sstc_fft (host tier)" if world == 1 else
                           "sstc_h2d + SlabFft3d.forward + sstc_d2h per rank"},
            "clocks": sampler.summary(t_wall0, t_load1) if sampler else None,
        }
        if roofline:
            line["roofline"] = roofline
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", dest="no_graph", help="issue the pipe schedule eagerly")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
