"""Experiment: time of a 512-point strided pass over 2^27 points as a function of the stride between the points of a
line (8 KB ... 4 MiB): where does the outermost pass of 512^3 lose its 15-20 %?"""
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

st = torch.cuda.current_stream().cuda_stream
s.init(0)
n = 512 ** 3
y = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5


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


res = {}
for rep in range(2):
    for inner in (512, 2048, 8192, 32768, 65536, 131072, 262144):
        outer = n // (512 * inner)
        res["stride_%dKB_run%d" % (inner * 16 // 1024, rep)] = ev(
            lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), outer, 512, inner, False, 1.0, st))
print(json.dumps(res))
