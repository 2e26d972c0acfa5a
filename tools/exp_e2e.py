"""Provider-call throughput (sstc_fft on 512^3, pinned and pageable arrays) by number of caller threads, slots and with /
without the per-direction link tokens. One JSON object."""
import ctypes
import json
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200 import _lib

s.init(0)
lib = _lib.lib()
dims = [512, 512, 512]
n = 512 ** 3
svc = s.CudaFftService(0)


def pinned(nd):
    p = ctypes.c_void_p()
    _lib.check(lib.sstc_host_alloc(ctypes.byref(p), nd * 8))
    return np.frombuffer((ctypes.c_double * nd).from_address(p.value), dtype=np.float64)


def run(arrays, callers, reps):
    def caller(i):
        for _ in range(reps):
            svc.fft(dims, arrays[i][0], arrays[i][1])
    th = [threading.Thread(target=caller, args=(i,)) for i in range(callers)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return (time.perf_counter() - t0) / (reps * callers) * 1e3


maxc = 3
pins = [(pinned(2 * n), pinned(2 * n)) for _ in range(maxc)]
pages = [(np.full(2 * n, 0.25), np.zeros(2 * n)) for _ in range(maxc)]
for a, _ in pins:
    a[:] = 0.25
out = {"runs": []}
for tokens in (1, 0):
    _lib.check(lib.sstc_host_option(b"link_tokens", tokens))
    for slots, callers in ((2, 1), (2, 2), (3, 3)):
        _lib.check(lib.sstc_host_option(b"slots", slots))
        for label, arrays in (("pinned", pins), ("pageable", pages)):
            run(arrays, callers, 1)
            ms = run(arrays, callers, 4)
            out["runs"].append({"link_tokens": tokens, "slots": slots, "callers": callers, "arrays": label, "ms_per_transform": round(ms, 2)})
print(json.dumps(out))
