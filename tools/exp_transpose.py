"""Experiment: is a large stride cheaper on the store side than on the load side? (512^3, per-pass times)"""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200 import _lib

n = 512 ** 3
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
w = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
s.init(0)
L = _lib.lib()


def strided(src, dst, outer, ln, inner, ip, ik, op, ok):
    _lib.check(L.sstc_fft_axis_strided_dev(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), outer, ln,
                                           inner, ip, ik, op, ok, 0, 1.0, ctypes.c_void_p(st)))


def ev(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)


P = 512 * 512
res = {}
res["y_inplace"] = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st))
res["y_outofplace"] = ev(lambda: s.fft_axis(x.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st))
res["y_transposing_store"] = ev(lambda: strided(x, y, 512, 512, 512, P, 512, 512, P))
res["x_inplace"] = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 1, 512, P, False, 1.0, st))
res["x_outofplace"] = ev(lambda: s.fft_axis(x.data_ptr(), y.data_ptr(), 1, 512, P, False, 1.0, st))
res["x_as_transposing_load"] = ev(lambda: strided(x, y, 512, 512, 512, 512, P, P, 512))
# correctness of the transposing pass: fft along y of (x,y,z), stored as (y,x,z)
strided(x, y, 512, 512, 512, P, 512, 512, P)
s.fft_axis(x.data_ptr(), w.data_ptr(), 512, 512, 512, False, 1.0, st)
ok = torch.equal(y.view(512, 512, 512, 2), w.view(512, 512, 512, 2).transpose(0, 1))
res["transposing_correct"] = bool(ok)
print(json.dumps(res))
