"""Experiment: does the y pass run faster when the z pass has just left its output in L2?"""
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

n = 512 ** 3
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
s.init(0)
P = 512 * 512  # complex elements per plane


def ev(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)


def chunked(cp, fixed_out):
    def run():
        for c in range(512 // cp):
            src = x.data_ptr() + c * cp * P * 16
            dst = y.data_ptr() + (0 if fixed_out else c * cp * P * 16)
            s.fft_axis(src, dst, cp * 512, 512, 1, False, 1.0, st)
            s.fft_axis(dst, dst, cp, 512, 512, False, 1.0, st)
    return run


res = {}
res["z_full+y_full"] = ev(lambda: (s.fft_axis(x.data_ptr(), y.data_ptr(), 512 * 512, 512, 1, False, 1.0, st),
                                   s.fft_axis(y.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st)))
for cp in (4, 8, 16):
    res["chunk%d_rotating_out" % cp] = ev(chunked(cp, False))
    res["chunk%d_fixed_out(L2)" % cp] = ev(chunked(cp, True))
# the y pass alone, repeatedly on the same 8 planes (fully L2 resident) vs on rotating planes
res["y_8planes_same"] = ev(lambda: [s.fft_axis(y.data_ptr(), y.data_ptr(), 8, 512, 512, False, 1.0, st) for _ in range(64)])
res["y_8planes_rotating"] = ev(lambda: [s.fft_axis(y.data_ptr() + c * 8 * P * 16, y.data_ptr() + c * 8 * P * 16, 8, 512, 512,
                                                   False, 1.0, st) for c in range(64)])
print(json.dumps(res))
