"""Profiling driver: a few forward 512^3 transforms on device-resident data (run under ncu by hand)."""
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

dims = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [512, 512, 512]
reps = int(sys.argv[4]) if len(sys.argv) >= 5 else 2
n = dims[0] * dims[1] * dims[2]
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
plan = s.FftPlan(s.FORWARD, dims)
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    plan.execute(x.data_ptr(), y.data_ptr(), st)
torch.cuda.synchronize()
print("ok", float(y[0, 0]))
