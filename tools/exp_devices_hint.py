"""The provider call (CudaFftService.fft on host arrays, 512^3) served by 1, 2, 4 ... devices of one box from ONE process
(FftService.setHint("devices", ...) -> sstc_fft3d_sharded_host): every device moves its slab over its own PCIe link.
Pinned and pageable arrays; result checked against the one-device result. One JSON object."""
import ctypes
import json
import sys
import time

import numpy as np
import torch

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


px, py = pinned(2 * n), pinned(2 * n)
rng = np.random.default_rng(5)
px[:] = rng.uniform(-0.5, 0.5, 2 * n)
gx, gy = px.copy(), np.zeros(2 * n)
svc.fft(dims, px, py)
ref = py.copy()
out = {"runs": []}
ndev = torch.cuda.device_count()
for p in [d for d in (1, 2, 4, 8) if d <= ndev]:
    svc.setHint("devices", ",".join(str(d) for d in range(p)) if p > 1 else "")
    for label, x, y in (("pinned", px, py), ("pageable", gx, gy)):
        svc.fft(dims, x, y)
        err = float(np.linalg.norm(y - ref) / np.linalg.norm(ref))
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            svc.fft(dims, x, y)
            best = min(best, (time.perf_counter() - t0) * 1e3)
        out["runs"].append({"devices": p, "arrays": label, "ms": round(best, 2), "rel_l2_vs_one_device": err})
svc.setHint("devices", "")
print(json.dumps(out))
