"""Tuning sweep (run on the GPU box): per-pass and per-step CUDA-event times at 512^3 for the current
SSTC_FFT512_* tile settings and the plan hints."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

dims = [512, 512, 512]
n = 512 ** 3
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
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


res = {"env": {k: v for k, v in os.environ.items() if k.startswith("SSTC_")}}
res["z"] = ev(lambda: s.fft_axis(x.data_ptr(), y.data_ptr(), 512 * 512, 512, 1, False, 1.0, st))
res["y"] = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st))
res["x"] = ev(lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 1, 512, 512 * 512, False, 1.0, st))
res["copy"] = ev(lambda: y.copy_(x))
plan = s.FftPlan(s.FORWARD, dims)
for hint, values in (("alternate_order", (0, 1)), ("transposed_passes", (1, 0))):
    for v in values:
        plan.set_hint(hint, v)
        res["step_%s=%d" % (hint, v)] = ev(lambda: plan.execute(x.data_ptr(), y.data_ptr(), st), 50)
xc = torch.view_as_complex(x).reshape(dims)
res["cufft_torch"] = ev(lambda: torch.fft.fftn(xc), 10)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
