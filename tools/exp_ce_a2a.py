"""Experiment: all-to-all rate of the copy engines over NVLink. Every rank copies one block to each peer's receive
buffer with cudaMemcpyAsync (one stream per peer), all ranks at once; compare with the SM-store exchange pass."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from shared_b200 import _lib
from shared_b200.sharded import PeerBuffer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
_lib.init(local)
lib = _lib.lib()
n_local = 512 ** 3 // world
block = n_local // world          # complex values per pairwise block
src = torch.rand(n_local, 2, dtype=torch.float64, device=dev)
peer = PeerBuffer(n_local, dev, None)
streams = [torch.cuda.Stream(dev) for _ in range(world)]
res = {}
for name, nstreams, pieces in (("one_stream_per_peer", world, 1), ("single_stream", 1, 1), ("per_peer_4_pieces", world, 4)):
    def run():
        cur = torch.cuda.current_stream(dev)
        ev = cur.record_event()
        for k in range(1, world):
            q = (rank + k) % world
            s = streams[k % nstreams]
            s.wait_event(ev)
            for pc in range(pieces):
                nb = block * 16 // pieces
                _lib.check(lib.sstc_d2d(ctypes.c_void_p(peer.ptrs[q] + rank * block * 16 + pc * nb),
                                        ctypes.c_void_p(src.data_ptr() + q * block * 16 + pc * nb), nb,
                                        ctypes.c_void_p(s.cuda_stream)))
        for s in streams[:nstreams]:
            cur.wait_stream(s)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 20
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out_bytes = (world - 1) * block * 16
    res[name] = {"ms": round(t.item(), 4), "egress_gbs_per_gpu": round(out_bytes / t.item() / 1e6, 1)}
    dist.barrier()
if rank == 0:
    print(json.dumps({"gpus": world, "egress_bytes_per_gpu": (world - 1) * block * 16, "copy_engine_all_to_all": res}))
dist.barrier()
peer.close()
dist.destroy_process_group()
