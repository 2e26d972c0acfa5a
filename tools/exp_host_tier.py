"""Where a 512^3 transform through the host tier (sstc_fft) spends its time, pinned and pageable arrays, as a function
of the copy-thread count and chunk size. One JSON object."""
import ctypes
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200 import _lib

s.init(0)
lib = _lib.lib()
dims = [512, 512, 512]
n = 512 ** 3
NAMES = ["total", "phase1_in", "phase2_out", "stage_in_copy", "stage_out_copy", "wait_ring_in", "wait_ring_out", "pipelined",
         "in_pinned", "out_pinned", "chunks_in", "groups_out"]


def call(x, y):
    _lib.check(lib.sstc_fft(s.FORWARD, 3, _lib.i32_array(dims), x.ctypes.data, x.size, y.ctypes.data, y.size))
    st = (ctypes.c_double * 12)()
    lib.sstc_host_last_stats(st, 12)
    return dict(zip(NAMES, [round(v, 2) for v in st]))


def pinned(nd):
    p = ctypes.c_void_p()
    _lib.check(lib.sstc_host_alloc(ctypes.byref(p), nd * 8))
    return np.frombuffer((ctypes.c_double * nd).from_address(p.value), dtype=np.float64)


out = {"runs": []}
px, py = pinned(2 * n), pinned(2 * n)
px[:] = 0.25
gx = np.full(2 * n, 0.25)
gy = np.zeros(2 * n)
for threads, chunk in [(8, 32), (4, 32), (12, 32), (16, 32), (8, 16), (8, 64), (8, 128)]:
    _lib.check(lib.sstc_host_option(b"copy_threads", threads))
    _lib.check(lib.sstc_host_option(b"chunk_mib", chunk))
    for label, x, y in (("pinned", px, py), ("pageable", gx, gy), ("pageable in, pinned out", gx, py), ("pinned in, pageable out", px, gy)):
        if label != "pageable" and (threads, chunk) != (8, 32):
            continue
        call(x, y)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            st = call(x, y)
            st["wall_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            if best is None or st["wall_ms"] < best["wall_ms"]:
                best = st
        out["runs"].append(dict(best, arrays=label, copy_threads=threads, chunk_mib=chunk))
print(json.dumps(out))
