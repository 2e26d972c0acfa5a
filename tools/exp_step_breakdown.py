"""Experiment: per-pass times INSIDE the back-to-back 512^3 step (events between the launches), for the
in-place schedule and the transposed-pass schedule."""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

n = 512 ** 3
P = 512 * 512
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
w = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
s.init(0)


def z():
    s.fft_axis(x.data_ptr(), y.data_ptr(), P, 512, 1, False, 1.0, st)


scheds = {
    "inplace": [z, lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st),
                lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 1, 512, P, False, 1.0, st)],
    "transposed": [z, lambda: s.fft_axis_strided(y.data_ptr(), w.data_ptr(), 512, 512, 512, P, 512, 512, P, False, 1.0, st),
                   lambda: s.fft_axis_strided(w.data_ptr(), y.data_ptr(), 512, 512, 512, P, 512, 512, P, False, 1.0, st)],
    "y_outofplace_x_inplace": [z, lambda: s.fft_axis(y.data_ptr(), w.data_ptr(), 512, 512, 512, False, 1.0, st),
                               lambda: s.fft_axis(w.data_ptr(), w.data_ptr(), 1, 512, P, False, 1.0, st)],
    "x_first": [lambda: s.fft_axis(x.data_ptr(), y.data_ptr(), 1, 512, P, False, 1.0, st),
                lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), 512, 512, 512, False, 1.0, st),
                lambda: s.fft_axis(y.data_ptr(), y.data_ptr(), P, 512, 1, False, 1.0, st)],
}
res = {}
for name, passes in scheds.items():
    reps = 30
    for f in passes:
        f()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
    torch.cuda.synchronize()
    for r in range(reps):
        evs[r][0].record()
        for i, f in enumerate(passes):
            f()
            evs[r][i + 1].record()
    torch.cuda.synchronize()
    t = [sum(evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(5, reps)) / (reps - 5) for i in range(3)]
    res[name] = {"passes_ms": [round(v, 4) for v in t], "step_ms": round(sum(t), 4)}
print(json.dumps(res))
