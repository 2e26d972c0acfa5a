"""Sweep of the pipelined slab exchange on N GPUs (run under torch.distributed.run): plane chunks of the local y
pass, y groups of the exchange pass, cap on the exchange grid. Every configuration is checked bit-for-bit against
the first one, then timed over CUDA-graph replays (CUDA events, max over ranks). Rank 0 prints one JSON object."""
import json
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
dims = [512, 512, 512]
n_loc = dims[0] * dims[1] * dims[2] // world
x = torch.rand(n_loc, 2, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank)) - 0.5
sms = torch.cuda.get_device_properties(dev).multi_processor_count
KEYS = ("order", "chunks", "groups", "exchange_ctas", "head_ctas", "bulk_store", "split_barrier", "begin_ahead")
configs = [dict(zip(KEYS, v)) for v in
           [(0, 4, 4, sms, 0, 0, 0, 0), (0, 4, 4, sms, 0, 0, 0, 1), (0, 4, 8, sms, 0, 0, 0, 1), (0, 4, 2, sms, 0, 0, 0, 1),
            (0, 4, 4, 120, 0, 0, 0, 1), (0, 4, 4, 190, 0, 0, 0, 1), (0, 4, 4, sms, 0, 2, 0, 1), (0, 4, 4, sms, 0, 0, 1, 1),
            (0, 4, 4, 2 * sms, 0, 2, 0, 1), (0, 4, 8, sms, 0, 0, 1, 1)]]
if len(sys.argv) > 1:
    configs = [dict(zip(KEYS, map(int, a.split(",")))) for a in sys.argv[1:]]
ref, out = None, []
for cfg in configs:
    ahead = cfg.pop("begin_ahead", 0)
    slab = SlabFft3d(dims, exchange="pipe", device=local, **cfg)
    cfg["begin_ahead"] = ahead
    y = slab.forward(x).clone()
    if ref is None:
        ref = y
    same = bool(torch.equal(y, ref))
    slab.capture(x)
    if ahead:
        slab.forward_begin(x)

    def step():
        if ahead:
            slab.forward_begin(x)
        return slab.forward(x)

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        step()
    b.record()
    if ahead:
        slab.forward(x)  # finishes the last begun transform
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 200], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.tensor([int(same and torch.equal(slab.forward(x), ref))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out.append(dict(cfg, ms=round(t.item(), 4), bit_identical=bool(ok.item())))
    torch.cuda.synchronize()
    dist.barrier()
    slab.close()
    del slab
if rank == 0:
    print(json.dumps({"gpus": world, "dims": dims, "configs": out}))
dist.barrier()
dist.destroy_process_group()
