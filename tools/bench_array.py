"""BASELINE.json configs[3] on one B200: ArrayKernel ops at 8192 x 8192 (and mul at 4096^2), device tier,
CUDA-event timing; algorithmic bytes per SURVEY.md section 8d. Optionally times the reference's native kernel
(oracle/_ref) on the host for the same ops at a bounded size (--cpu). Prints one JSON object."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from shared_b200.array_kernel import REAL, DeviceArrayKernel

OP = dict(RU_MUL=1, RU_EXP=4, RU_SQRT=7, RE_ADD=0, RE_MUL=2, RR_SUM=0, RR_MAX=2, RR_VAR=4, RD_SUM=0, RI_MAX=0, RA_SUM=0)
n = 8192
N = n * n
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
dk = DeviceArrayKernel(0, torch.cuda.current_stream().cuda_stream)
x = torch.rand(N, dtype=torch.float64, device="cuda") + 0.5
y = torch.rand(N, dtype=torch.float64, device="cuda") + 0.5
z = torch.empty_like(x)
row = torch.empty(n, dtype=torch.float64, device="cuda")
idx = torch.empty(N, dtype=torch.int32, device="cuda")
one = torch.empty(2, dtype=torch.float64, device="cuda")
far = [n, 1]


def ev(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


P = x.data_ptr()
cases = [
    ("ruOp MUL", 16 * N, lambda: dk.ruOp(OP["RU_MUL"], 1.0000001, z.data_ptr(), N)),
    ("ruOp SQRT", 16 * N, lambda: dk.ruOp(OP["RU_SQRT"], 0, z.data_ptr(), N)),
    ("ruOp EXP", 16 * N, lambda: dk.ruOp(OP["RU_EXP"], 0, z.data_ptr(), N)),
    ("eOp ADD", 24 * N, lambda: dk.eOp(OP["RE_ADD"], REAL, P, y.data_ptr(), z.data_ptr(), N)),
    ("eOp MUL", 24 * N, lambda: dk.eOp(OP["RE_MUL"], REAL, P, y.data_ptr(), z.data_ptr(), N)),
    ("raOp SUM", 8 * N, lambda: dk.raOp(OP["RA_SUM"], P, N, one.data_ptr())),
]
for name, op in (("SUM", "RR_SUM"), ("MAX", "RR_MAX"), ("VAR", "RR_VAR")):
    mult = 2 if name == "VAR" else 1
    cases.append(("rrOp %s dim1 (contiguous)" % name, 8 * N * mult,
                  lambda op=op: dk.rrOp(OP[op], P, N, [n, n], far, row.data_ptr(), n, [n, 1], [1, 1], 1)))
    cases.append(("rrOp %s dim0 (strided)" % name, 8 * N * mult,
                  lambda op=op: dk.rrOp(OP[op], P, N, [n, n], far, row.data_ptr(), n, [1, n], [n, 1], 0)))
cases += [
    ("rdOp SUM dim1", 16 * N, lambda: dk.rdOp(OP["RD_SUM"], P, N, [n, n], far, z.data_ptr(), N, 1)),
    ("rdOp SUM dim0", 16 * N, lambda: dk.rdOp(OP["RD_SUM"], P, N, [n, n], far, z.data_ptr(), N, 0)),
    ("rdOp SUM dims 0,1", 32 * N, lambda: dk.rdOp(OP["RD_SUM"], P, N, [n, n], far, z.data_ptr(), N, 0, 1)),
    ("riOp MAX dim1 (12 B/elt: SURVEY 8d)", 12 * N, lambda: dk.riOp(OP["RI_MAX"], P, N, [n, n], far, idx.data_ptr(), N, 1)),
    ("map transpose", 16 * N, lambda: dk.map(8, [0, 0, n, 0, 0, n], P, N, [n, n], far, z.data_ptr(), N, [n, n], [1, n])),
    ("map copy", 16 * N, lambda: dk.map(8, [0, 0, n, 0, 0, n], P, N, [n, n], far, z.data_ptr(), N, [n, n], far)),
]
from shared_b200.image_kernel import DeviceImageKernel
ik = DeviceImageKernel(0, torch.cuda.current_stream().cuda_stream)
ii = torch.zeros((n + 1) * (n + 1), dtype=torch.float64, device="cuda")
cases.append(("ImageKernel createIntegralImage (16 B/elt minimum)", 8 * N + 8 * (n + 1) * (n + 1),
              lambda: ik.createIntegralImage(P, N, [n, n], far, ii.data_ptr(), ii.numel(), [n + 1, n + 1], [n + 1, 1])))
z.copy_(x)
res = {"shape": [n, n], "peak_gbs": peak, "ops": []}
for name, nbytes, fn in cases:
    ms = ev(fn)
    res["ops"].append({"op": name, "ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1),
                       "frac_of_measured_hbm": round(nbytes / ms / 1e6 / peak, 3)})
m = 4096
a = torch.rand(m * m, dtype=torch.float64, device="cuda") * 2 - 1
b = torch.rand(m * m, dtype=torch.float64, device="cuda") * 2 - 1
c = torch.empty(m * m, dtype=torch.float64, device="cuda")
ms = ev(lambda: dk.mul(a.data_ptr(), m * m, b.data_ptr(), m * m, m, m, c.data_ptr(), m * m, False), 5)
res["mul_4096_real"] = {"ms": round(ms, 3), "gflops": round(2 * m ** 3 / ms / 1e6, 1)}
ms = ev(lambda: torch.matmul(a.view(m, m), b.view(m, m)), 5)
res["mul_4096_real_cublas_yardstick"] = {"ms": round(ms, 3), "gflops": round(2 * m ** 3 / ms / 1e6, 1)}
m2 = 2048
ca = torch.rand(2 * m2 * m2, dtype=torch.float64, device="cuda")
cb = torch.rand(2 * m2 * m2, dtype=torch.float64, device="cuda")
cc = torch.empty(2 * m2 * m2, dtype=torch.float64, device="cuda")
ms = ev(lambda: dk.mul(ca.data_ptr(), 2 * m2 * m2, cb.data_ptr(), 2 * m2 * m2, m2, m2, cc.data_ptr(), 2 * m2 * m2, True), 5)
res["mul_2048_complex"] = {"ms": round(ms, 3), "gflops": round(8 * m2 ** 3 / ms / 1e6, 1)}

if "--cpu" in sys.argv:
    import oracle
    ref = oracle.ref()
    cn = 2048  # bounded sample: 2048 x 2048 (1/16 of the elements), the reference's own single-threaded C++
    cx = np.random.default_rng(0).uniform(0.5, 1.5, cn * cn)
    cy = cx.copy()
    cz = np.zeros(cn * cn)
    crow = np.zeros(cn)
    cidx = np.zeros(cn * cn, dtype=np.int32)

    def cpu(fn):
        t0 = time.perf_counter()
        fn()
        return time.perf_counter() - t0

    cres = []
    for name, nbytes, fn in [
        ("ruOp MUL", 16, lambda: ref.ruOp(1, 1.0000001, cz)), ("ruOp EXP", 16, lambda: ref.ruOp(4, 0, cz * 0 + 1)),
        ("eOp ADD", 24, lambda: ref.eOp(0, cx, cy, cz, False)),
        ("rrOp SUM dim1", 8, lambda: ref.rrOp(0, cx, [cn, cn], [cn, 1], crow, [cn, 1], [1, 1], 1)),
        ("rrOp SUM dim0", 8, lambda: ref.rrOp(0, cx, [cn, cn], [cn, 1], crow, [1, cn], [cn, 1], 0)),
        ("rdOp SUM dim1", 16, lambda: ref.rdOp(0, cx, [cn, cn], [cn, 1], cz, 1)),
        ("riOp MAX dim1", 20, lambda: ref.riOp(0, cx, [cn, cn], [cn, 1], cidx, 1)),
    ]:
        s = cpu(fn)
        cres.append({"op": name, "seconds": round(s, 4), "gbs": round(nbytes * cn * cn / s / 1e9, 2)})
    mm = 512
    ma, mb, mc = cx[:mm * mm].copy(), cx[:mm * mm].copy(), np.zeros(mm * mm)
    s = cpu(lambda: ref.mul(ma, mb, mm, mm, mc, False))
    cres.append({"op": "mul 512^2 real", "seconds": round(s, 4), "gflops": round(2 * mm ** 3 / s / 1e9, 3)})
    res["cpu_reference"] = {"kind": "reference", "cores": 1, "sample": "2048x2048 (mul: 512^2), oracle/_ref = the "
                            "reference's native/src/shared/*.cpp, g++ -O2, single thread", "ops": cres}
print(json.dumps(res))
