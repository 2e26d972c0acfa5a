"""TMA tile movement against per-thread loads/stores on the strided passes: the y and x passes of the 512^3
transform and the 4096-point column pass of the 4096 x 2049 half-spectrum, each timed alone with CUDA events, results
compared bit for bit with the register path. One JSON object."""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200 import _lib

s.init(0)
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream


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
    return a.elapsed_time(b) / reps


def tma(v):
    _lib.check(lib.sstc_exp_set(b"tma_strided", v))


out = {"cases": []}
n = 512 ** 3
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
ref = torch.empty_like(x)
for name, (outer, L, inner) in (("512^3 y pass (stride 8 KB)", (512, 512, 512)), ("512^3 x pass (stride 4 MiB)", (1, 512, 512 * 512))):
    tma(0)
    s.fft_axis(x.data_ptr(), ref.data_ptr(), outer, L, inner, False, 1.0, st)
    rec = {"case": name, "bytes": 32 * n}
    for mode, label in ((0, "registers (LDG/STG)"), (1, "tma, 1 buffer, 2 CTAs/SM"), (2, "tma, 2 buffers"), (3, "tma, 3 buffers, persistent")):
        tma(mode)
        y.zero_()
        print("running", name, label, file=sys.stderr, flush=True)
        s.fft_axis(x.data_ptr(), y.data_ptr(), outer, L, inner, False, 1.0, st)
        torch.cuda.synchronize()
        same = bool(torch.equal(y, ref))
        ms_oop = ev(lambda: s.fft_axis(x.data_ptr(), y.data_ptr(), outer, L, inner, False, 1.0, st))
        ms_inp = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), outer, L, inner, False, 1.0, st))
        rec[label] = {"out_of_place_ms": ms_oop, "in_place_ms": ms_inp, "in_place_gbs": 32 * n / ms_inp / 1e6, "bit_identical": same}
    tma(0)
    for ahead in (74, 148, 296, 592, 1184, 2368):
        _lib.check(lib.sstc_exp_set(b"l2_prefetch", ahead))
        y.zero_()
        s.fft_axis(x.data_ptr(), y.data_ptr(), outer, L, inner, False, 1.0, st)
        torch.cuda.synchronize()
        same = bool(torch.equal(y, ref))
        ms_inp = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), outer, L, inner, False, 1.0, st))
        rec["registers + L2 prefetch %d tiles ahead" % ahead] = {"in_place_ms": ms_inp, "in_place_gbs": 32 * n / ms_inp / 1e6,
                                                               "bit_identical": same}
    _lib.check(lib.sstc_exp_set(b"l2_prefetch", 0))
    out["cases"].append(rec)
del x, y, ref
torch.cuda.empty_cache()
rows, cols = 4096, 2049
m = rows * cols
z = torch.rand(m, 2, dtype=torch.float64, device="cuda") - 0.5
w, ref = torch.empty_like(z), torch.empty_like(z)
tma(0)
s.fft_axis(z.data_ptr(), ref.data_ptr(), 1, rows, cols, False, 1.0, st)
rec = {"case": "4096-point column pass of the 4096 x 2049 half-spectrum", "bytes": 32 * m}
for mode, label in ((0, "registers (LDG/STG)"), (1, "tma, 1 buffer")):
    tma(mode)
    w.zero_()
    s.fft_axis(z.data_ptr(), w.data_ptr(), 1, rows, cols, False, 1.0, st)
    same = bool(torch.equal(w, ref))
    ms = ev(lambda: s.fft_axis(w.data_ptr(), w.data_ptr(), 1, rows, cols, False, 1.0, st))
    rec[label] = {"in_place_ms": ms, "in_place_gbs": 32 * m / ms / 1e6, "bit_identical": same}
out["cases"].append(rec)
tma(0)
print(json.dumps(out))
