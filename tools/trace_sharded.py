"""Timeline of one pipelined slab step on every rank (run under torch.distributed.run): the step is issued eagerly
with a timing event behind every launch (sstc_slab option "trace"), after warm-up, and compared with the replayed
step's duration. Rank 0 prints one JSON object: per rank [(label, ms)], plus the graph-replay time for reference."""
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
opts = dict(a.split("=") for a in sys.argv[1:])
slab = SlabFft3d(dims, exchange="pipe", device=local, **{k: int(v) for k, v in opts.items()})


def timed(reps):
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        slab.forward(x)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


slab.capture(x)
timed(10)
graph_ms = timed(100)
slab.set_option("trace", 1)
for _ in range(4):
    slab.forward(x)
eager_ms = timed(20)
torch.cuda.synchronize()
dist.barrier()
slab.forward(x)
torch.cuda.synchronize()
mine = slab.trace()
allt = [None] * world
dist.all_gather_object(allt, mine)
if rank == 0:
    print(json.dumps({"gpus": world, "options": opts, "graph_replay_ms": graph_ms, "eager_traced_ms": eager_ms, "ranks": allt}))
dist.barrier()
slab.close()
dist.destroy_process_group()
