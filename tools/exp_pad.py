"""Experiment: is the outermost-axis pass slow because its stride is a power of two (L2-slice / channel
hot-spotting)? Time the x pass of 512^3 with a padded plane pitch on the read side, the write side, or both."""
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

P = 512 * 512
st = torch.cuda.current_stream().cuda_stream
s.init(0)
res = {}


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


for pad in (0, 512, 8 * 512, 520, 64):
    Pp = P + pad
    a = torch.rand(512 * Pp, 2, dtype=torch.float64, device="cuda") - 0.5
    b = torch.empty_like(a)
    nat = torch.empty(512 * P, 2, dtype=torch.float64, device="cuda")
    r = {}
    r["read_padded_write_padded_inplace"] = ev(lambda: s.fft_axis_strided(a.data_ptr(), a.data_ptr(), 1, 512, P, 0, Pp, 0, Pp,
                                                                          False, 1.0, st))
    r["read_padded_write_natural"] = ev(lambda: s.fft_axis_strided(a.data_ptr(), nat.data_ptr(), 1, 512, P, 0, Pp, 0, P,
                                                                   False, 1.0, st))
    r["read_natural_write_padded"] = ev(lambda: s.fft_axis_strided(nat.data_ptr(), b.data_ptr(), 1, 512, P, 0, P, 0, Pp,
                                                                   False, 1.0, st))
    res["pad%d" % pad] = r
    del a, b, nat
print(json.dumps(res))
