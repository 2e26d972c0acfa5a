"""The exchange pass of the slab transform in isolation (one process, two devices): sstc_fft_lines_scatter_ex_dev on
device 0 with a share of its destination blocks in device 1's memory. How close does the FFT-and-store kernel get to
the plain store rate of tools/exp_nvlink.py, as a function of the grid cap and the store variant?

    python tools/exp_exchange.py > gpurun_out/exp_exchange.json
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
planes, lines, L = 64, 512, 512            # one rank's slab at 8 ranks: 64 x 512 x 512 complex128 = 256 MiB
n = planes * lines * L
src = torch.rand(n, 2, dtype=torch.float64, device="cuda:0") - 0.5
local = torch.empty(n, 2, dtype=torch.float64, device="cuda:0")
peer = torch.empty(n, 2, dtype=torch.float64, device="cuda:1")
rt = ctypes.CDLL("libcudart.so.12")
for d in range(2):
    with torch.cuda.device(d):
        rc = rt.cudaDeviceEnablePeerAccess(1 - d, 0)
        assert rc in (0, 704), rc  # 704 = already enabled
        rt.cudaGetLastError()
        torch.cuda.synchronize()
st = torch.cuda.current_stream(torch.device("cuda:0")).cuda_stream
nblk = 8
block = n // nblk  # complex elements per destination block
out = {"slab": [planes, lines, L], "bytes_written": 16 * n, "runs": []}
for remote_blocks in (0, 4, 7, 8):
    ptrs = [(peer if q < remote_blocks else local).data_ptr() + q * block * 16 for q in range(nblk)]
    arr = (ctypes.c_void_p * nblk)(*ptrs)
    for bulk in (0, 1, 2):
        for ctas in (148, 296, 0):
            def run():
                _lib.check(lib.sstc_fft_lines_scatter_ex_dev(ctypes.c_void_p(src.data_ptr()), arr, nblk, planes, lines, L, 0,
                                                             lines // nblk, 0, 1.0, ctas, bulk, ctypes.c_void_p(st)))
            with torch.cuda.device(0):
                for _ in range(3):
                    run()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(20):
                    run()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 20
            remote_bytes = 16 * block * remote_blocks
            out["runs"].append({"remote_blocks_of_8": remote_blocks, "bulk": bulk, "max_ctas": ctas, "ms": ms,
                                "nvlink_gbs": remote_bytes / ms / 1e6, "hbm_read_gbs": 16 * n / ms / 1e6})
# correctness of the bulk variant: same bits as the store variant
res = []
for bulk in (0, 1, 2):
    peer.zero_(); local.zero_()
    ptrs = [(peer if q < 4 else local).data_ptr() + q * block * 16 for q in range(nblk)]
    arr = (ctypes.c_void_p * nblk)(*ptrs)
    with torch.cuda.device(0):
        _lib.check(lib.sstc_fft_lines_scatter_ex_dev(ctypes.c_void_p(src.data_ptr()), arr, nblk, planes, lines, L, 0, lines // nblk,
                                                     0, 1.0, 148, bulk, ctypes.c_void_p(st)))
        torch.cuda.synchronize()
    res.append((peer.cpu().clone(), local.cpu().clone()))
out["bulk_bit_identical"] = bool(all(torch.equal(res[0][0], r[0]) and torch.equal(res[0][1], r[1]) for r in res[1:]))
print(json.dumps(out))
