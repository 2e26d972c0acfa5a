"""Experiment: strided axis passes of length 1024 / 2048 with the narrow (4 / 2 columns) and wide (8 / 4) tiles."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

st = torch.cuda.current_stream().cuda_stream
s.init(0)


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


n = 4096 * 4096
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
res = {"SSTC_FFT_WIDE_C": os.environ.get("SSTC_FFT_WIDE_C", "1"), "bytes_MB": 32 * n / 1e6, "roof_us_at_6545": 32 * n / 6545e3}
res["L1024_inner16384"] = ev(lambda: s.fft_axis(x.data_ptr(), x.data_ptr(), 1, 1024, 16384, False, 1.0, st))
res["L1024_inner1024_outer16"] = ev(lambda: s.fft_axis(x.data_ptr(), x.data_ptr(), 16, 1024, 1024, False, 1.0, st))
res["L2048_inner8192"] = ev(lambda: s.fft_axis(x.data_ptr(), x.data_ptr(), 1, 2048, 8192, False, 1.0, st))
res["L2048_inner2048_outer4"] = ev(lambda: s.fft_axis(x.data_ptr(), x.data_ptr(), 4, 2048, 2048, False, 1.0, st))
print(json.dumps(res))
