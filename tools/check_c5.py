"""BASELINE.json configs[4]: ComplexArray 1024^3 complex128 fft/ifft round trip, slab-sharded over 8 B200
(16 GiB of data; not representable as one Java array). Run under torch.distributed.run with 8 ranks.
Checks the round-trip error against the input, Parseval on the spectrum, and times both directions."""
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from shared_b200.sharded import SlabFft3d

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dims = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [1024, 1024, 1024]
n = dims[0] * dims[1] * dims[2]
n_loc = n // world
g = torch.Generator(device=dev).manual_seed(20261019 + rank)
x = torch.rand(n_loc, 2, dtype=torch.float64, device=dev, generator=g) - 0.5
slab = SlabFft3d(dims, exchange="pipe", device=local)


def timed(fn, reps):
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, t.item()


y, _ = timed(lambda: slab.forward(x), 2)
y, ms_f = timed(lambda: slab.forward(x), 5)
spec = y.clone()
ex = torch.tensor([(x * x).sum().item(), (spec * spec).sum().item()], dtype=torch.float64, device=dev)
dist.all_reduce(ex)
back, _ = timed(lambda: slab.inverse(spec.clone()), 1)
back, ms_i = timed(lambda: slab.inverse(spec.clone()), 3)
err = torch.tensor([((back - x) ** 2).sum().item(), (x * x).sum().item()], dtype=torch.float64, device=dev)
dist.all_reduce(err)
if rank == 0:
    rel = math.sqrt(err[0].item() / err[1].item())
    flops = 5.0 * n * math.log2(n)
    print(json.dumps({"config": "ComplexArray %dx%dx%d complex128 3-D fft/ifft round trip, %d B200" % (*dims, world),
                      "roundtrip_rel_l2": rel, "tolerance": 1e-12 * math.log2(n), "ok": rel < 1e-12 * math.log2(n),
                      "parseval_ratio_minus_1": ex[1].item() / (n * ex[0].item()) - 1.0,
                      "forward_ms": ms_f, "forward_tflops": flops / ms_f / 1e9,
                      "inverse_ms_incl_clone": ms_i, "bytes_per_gpu": 16 * n_loc}))
dist.barrier()
slab.close()
dist.destroy_process_group()
