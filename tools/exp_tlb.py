"""Experiment: do the 512^3 passes depend on the memory footprint they touch (TLB reach)? Per-pass CUDA-event times
for (A) separate in / out buffers, (B) one buffer, every pass in place, (C) in / out carved from one allocation."""
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

n = 512 ** 3
P = 512 * 512
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


def passes(i, o):
    return {"z": ev(lambda: s.fft_axis(i, o, P, 512, 1, False, 1.0, st)),
            "y": ev(lambda: s.fft_axis(o, o, 512, 512, 512, False, 1.0, st)),
            "x": ev(lambda: s.fft_axis(o, o, 1, 512, P, False, 1.0, st)),
            "step": ev(lambda: (s.fft_axis(i, o, P, 512, 1, False, 1.0, st), s.fft_axis(o, o, 512, 512, 512, False, 1.0, st),
                                s.fft_axis(o, o, 1, 512, P, False, 1.0, st)))}


res = {}
big = torch.empty(2 * n, 2, dtype=torch.float64, device="cuda")
big.uniform_(-0.5, 0.5)
res["C_one_allocation_in_out"] = passes(big.data_ptr(), big.data_ptr() + 16 * n)
res["B_single_buffer_all_in_place"] = passes(big.data_ptr(), big.data_ptr())
del big
torch.cuda.empty_cache()
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
res["A_separate_in_out"] = passes(x.data_ptr(), y.data_ptr())
pad = torch.empty(4 * n, 2, dtype=torch.float64, device="cuda")  # 8 GiB allocated, never touched by the passes
res["A_with_8GiB_bystander"] = passes(x.data_ptr(), y.data_ptr())
print(json.dumps(res))
