"""Does an HBM-bound pass run beside the NVLink-bound exchange pass, or do they serialise? (one process, two devices)

Device 0 runs the exchange pass of one 8-rank slab (64 x 512 x 512, every destination block in device 1's memory)
on one stream and the x pass of the received groups (512 x 8192-column groups, in place) on a high-priority stream,
alone and together. Variants: the tile kernel with per-thread stores / the software-pipelined kernel, grid caps, and
the pipelined kernel holding its SMs exclusively (shared-memory request = a whole SM).

    python tools/exp_overlap.py > gpurun_out/exp_overlap.json
"""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
from shared_b200 import _lib

for d in range(2):
    _lib.init(d)
lib = _lib.lib()
rt = ctypes.CDLL("libcudart.so.12")
for d in range(2):
    with torch.cuda.device(d):
        rc = rt.cudaDeviceEnablePeerAccess(1 - d, 0)
        assert rc in (0, 704), rc
        rt.cudaGetLastError()
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
planes, lines, L = 64, 512, 512
n = planes * lines * L
src = torch.rand(n, 2, dtype=torch.float64, device=dev) - 0.5
peer = torch.empty(n, 2, dtype=torch.float64, device="cuda:1")
recv = torch.rand(n, 2, dtype=torch.float64, device=dev) - 0.5   # (512, 64*512) received slab
nblk = 8
block = n // nblk
arr = (ctypes.c_void_p * nblk)(*[peer.data_ptr() + q * block * 16 for q in range(nblk)])
s_e = torch.cuda.Stream(dev)
s_x = torch.cuda.Stream(dev, priority=-1)
d0, kstride, cols = 512, 64 * 512, 16 * 512


def exchange(bulk, ctas):
    _lib.check(lib.sstc_fft_lines_scatter_ex_dev(ctypes.c_void_p(src.data_ptr()), arr, nblk, planes, lines, L, 0, lines // nblk, 0,
                                                 1.0, ctas, bulk, ctypes.c_void_p(s_e.cuda_stream)))


def xpass(g):
    p = recv.data_ptr() + g * cols * 16
    _lib.check(lib.sstc_fft_axis_strided_dev(ctypes.c_void_p(p), ctypes.c_void_p(p), 1, d0, cols, 0, kstride, 0, kstride, 0, 1.0,
                                             ctypes.c_void_p(s_x.cuda_stream)))


def run(do_e, do_x, bulk, ctas, reps=6):
    torch.cuda.synchronize()
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in "ex"}
    if do_e:
        ev["e"][0].record(s_e)
        for _ in range(reps):
            exchange(bulk, ctas)
        ev["e"][1].record(s_e)
    if do_x:
        ev["x"][0].record(s_x)
        for _ in range(reps):
            for g in range(4):
                xpass(g)
        ev["x"][1].record(s_x)
    torch.cuda.synchronize()
    return (ev["e"][0].elapsed_time(ev["e"][1]) / reps if do_e else None,
            ev["x"][0].elapsed_time(ev["x"][1]) / reps if do_x else None)


out = {"what": "ms per slab: exchange of 256 MiB to the peer; x pass of the 4 groups of the received slab", "runs": []}
xpass(0)
exchange(0, 148)
exchange(2, 148)
_, x_alone = run(False, True, 0, 0)
out["x_alone_ms"] = x_alone
for bulk, ctas, smem in [(0, 148, 0), (0, 296, 0), (1, 148, 0), (2, 148, 0), (2, 74, 0), (2, 100, 0), (2, 74, 227), (2, 100, 227),
                         (2, 48, 227), (2, 148, 227)]:
    _lib.check(lib.sstc_exp_set(b"exchange_smem_kb", smem))
    e_alone, _ = run(True, False, bulk, ctas)
    e_both, x_both = run(True, True, bulk, ctas)
    out["runs"].append({"kernel": {0: "tile kernel, per-thread stores", 1: "tile kernel, bulk stores", 2: "pipelined"}[bulk],
                        "ctas": ctas, "exclusive_smem_kb": smem, "exchange_alone_ms": e_alone, "exchange_beside_x_ms": e_both,
                        "x_beside_exchange_ms": x_both, "nvlink_gbs_alone": 16 * n / e_alone / 1e6,
                        "nvlink_gbs_beside_x": 16 * n / e_both / 1e6})
_lib.check(lib.sstc_exp_set(b"exchange_smem_kb", 0))
print(json.dumps(out))
