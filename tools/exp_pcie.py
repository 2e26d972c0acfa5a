"""Host <-> device transfer rates of this box, the numbers the host tier's pipeline is designed around.

    python tools/exp_pcie.py [--mib 2048] > gpurun_out/exp_pcie.json

Measures, for one 2 GiB array (the 512^3 complex128 input): pinned H2D and D2H alone, both directions at once
(two streams), chunked transfers (is there a per-copy cost?), the driver's pageable path, and host memcpy with 1..16
threads (what a pinned staging ring would have to sustain for pageable JVM arrays). Prints one JSON object.
"""
import argparse
import ctypes
import json
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")


def timed(fn, reps=3):
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=2048)
    args = ap.parse_args()
    nbytes = args.mib << 20
    dev = torch.device("cuda", 0)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_a.fill_(1)
    h_b.fill_(2)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {"bytes": nbytes}

    def h2d():
        d_a.copy_(h_a, non_blocking=True)

    def d2h():
        h_b.copy_(d_b, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    h2d(); d2h(); both()
    out["pinned_h2d_gbs"] = nbytes / timed(h2d) / 1e9
    out["pinned_d2h_gbs"] = nbytes / timed(d2h) / 1e9
    t = timed(both)
    out["pinned_bidirectional_gbs_each"] = nbytes / t / 1e9
    out["pinned_bidirectional_ms"] = t * 1e3

    chunked = {}
    for chunk_mib in (4, 16, 64, 256):
        c = chunk_mib << 20

        def run():
            for off in range(0, nbytes, c):
                d_a[off:off + c].copy_(h_a[off:off + c], non_blocking=True)
        run()
        chunked[str(chunk_mib)] = nbytes / timed(run) / 1e9
    out["pinned_h2d_chunked_gbs_by_chunk_mib"] = chunked

    # the driver's pageable path (what a plain cudaMemcpy from a JVM heap array gets)
    p_a = np.ones(nbytes, dtype=np.uint8)
    p_b = np.empty(nbytes, dtype=np.uint8)
    p_b[:] = 0  # touch the pages
    t_a, t_b = torch.from_numpy(p_a), torch.from_numpy(p_b)

    def h2d_pageable():
        d_a.copy_(t_a)

    def d2h_pageable():
        t_b.copy_(d_b)

    h2d_pageable(); d2h_pageable()
    out["pageable_h2d_gbs"] = nbytes / timed(h2d_pageable) / 1e9
    out["pageable_d2h_gbs"] = nbytes / timed(d2h_pageable) / 1e9

    # host memcpy pageable -> pinned with k threads (ctypes.memmove releases the GIL)
    src, dst = p_a.ctypes.data, h_a.data_ptr()
    memcpy = {}
    for k in (1, 2, 4, 8, 12, 16):
        part = (nbytes // k) & ~4095

        def worker(i):
            n = part if i < k - 1 else nbytes - part * (k - 1)
            ctypes.memmove(dst + i * part, src + i * part, n)

        def run():
            th = [threading.Thread(target=worker, args=(i,)) for i in range(k)]
            for x in th:
                x.start()
            for x in th:
                x.join()
        run()
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            run()
            best = min(best, time.perf_counter() - t0)
        memcpy[str(k)] = nbytes / best / 1e9
    out["host_memcpy_pageable_to_pinned_gbs_by_threads"] = memcpy

    # cudaHostRegister of a pageable 2 GiB array (the alternative to staging): cost per call
    rt = ctypes.CDLL("libcudart.so.12")
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(ctypes.c_void_p(p_a.ctypes.data), ctypes.c_size_t(nbytes), 0)
    out["host_register_ms"] = (time.perf_counter() - t0) * 1e3
    out["host_register_rc"] = rc
    if rc == 0:
        reg = torch.from_numpy(p_a)

        def h2d_registered():
            rt.cudaMemcpyAsync(ctypes.c_void_p(d_a.data_ptr()), ctypes.c_void_p(p_a.ctypes.data), ctypes.c_size_t(nbytes), 1,
                               ctypes.c_void_p(0))
        h2d_registered()
        out["registered_h2d_gbs"] = nbytes / timed(h2d_registered) / 1e9
        t0 = time.perf_counter()
        rt.cudaHostUnregister(ctypes.c_void_p(p_a.ctypes.data))
        out["host_unregister_ms"] = (time.perf_counter() - t0) * 1e3
        del reg
    import os
    out["host_cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
