"""One launch sequence of every ArrayKernel / ImageKernel / convolution kernel at the BASELINE configs[3] shapes,
meant to run under `ncu -k regex:sstc --metrics ...` (one warm-up call, one measured call per operation)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200 import _lib
from shared_b200.array_kernel import COMPLEX, REAL, DeviceArrayKernel
from shared_b200.image_kernel import DeviceImageKernel

st = torch.cuda.current_stream().cuda_stream
dk, ik = DeviceArrayKernel(0, st), DeviceImageKernel(0, st)
n = 8192
N = n * n
x = torch.rand(N, dtype=torch.float64, device="cuda") + 0.5
y = torch.rand(N, dtype=torch.float64, device="cuda") + 0.5
z = x.clone()
row = torch.empty(n, dtype=torch.float64, device="cuda")
idx = torch.empty(N, dtype=torch.int32, device="cuda")
one = torch.empty(2, dtype=torch.float64, device="cuda")
ii = torch.zeros((n + 1) * (n + 1), dtype=torch.float64, device="cuda")
far = [n, 1]
P = x.data_ptr()
m = 4096
a = torch.rand(m * m, dtype=torch.float64, device="cuda")
b = torch.rand(m * m, dtype=torch.float64, device="cuda")
c = torch.empty(m * m, dtype=torch.float64, device="cuda")
h = 4096 * 2049
spec = torch.rand(h, 2, dtype=torch.float64, device="cuda")
kerf = torch.rand(h, 2, dtype=torch.float64, device="cuda")
crop = torch.empty(4066 * 4066, dtype=torch.float64, device="cuda")
ri = s.FftPlan(s.C2R, [4096, 4096])
od = np.array([4066, 4066], dtype=np.int32)
ops = [
    lambda: dk.ruOp(1, 1.0000001, z.data_ptr(), N), lambda: dk.ruOp(7, 0, z.data_ptr(), N),
    lambda: dk.ruOp(4, 0, z.data_ptr(), N), lambda: dk.eOp(0, REAL, P, y.data_ptr(), z.data_ptr(), N),
    lambda: dk.raOp(0, P, N, one.data_ptr()),
    lambda: dk.rrOp(0, P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1),
    lambda: dk.rrOp(0, P, N, [n, n], far, row.data_ptr(), n, [1, n], [n, 1], 0),
    lambda: dk.rrOp(2, P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1),
    lambda: dk.rrOp(4, P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1),
    lambda: dk.rdOp(0, P, N, [n, n], far, z.data_ptr(), N, 1), lambda: dk.rdOp(0, P, N, [n, n], far, z.data_ptr(), N, 0),
    lambda: dk.riOp(0, P, N, [n, n], far, idx.data_ptr(), N, 1),
    lambda: dk.map(8, [0, 0, n, 0, 0, n], P, N, [n, n], far, z.data_ptr(), N, [n, n], [1, n]),
    lambda: dk.map(8, [0, 0, n, 0, 0, n], P, N, [n, n], far, z.data_ptr(), N, [n, n], far),
    lambda: ik.createIntegralImage(P, N, [n, n], far, ii.data_ptr(), ii.numel(), [n + 1, n + 1], [n + 1, 1]),
    lambda: dk.mul(a.data_ptr(), m * m, b.data_ptr(), m * m, m, m, c.data_ptr(), m * m, False),
    lambda: dk.mul(a.data_ptr(), m * m // 2, b.data_ptr(), m * m // 2, 2048, 2048, c.data_ptr(), m * m // 2, True),
    lambda: _lib.check(_lib.lib().sstc_fft_convolve_dev(ri._h, ctypes.c_void_p(spec.data_ptr()),
                                                       ctypes.c_void_p(kerf.data_ptr()), ctypes.c_void_p(crop.data_ptr()),
                                                       od.ctypes.data_as(_lib.c_i32_p), ctypes.c_void_p(st))),
]
for rep in range(2):
    for f in ops:
        f()
    torch.cuda.synchronize()
print("done")
