"""Multi-GPU parity check (run under torch.distributed.run on N GPUs): the slab-sharded transform against the
single-GPU plan on the gathered data, both exchange modes, forward and inverse. Prints one line per case."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200.sharded import SlabFft3d, gather_transposed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true", help="skip the 512^3 case (used by tests/test_sharded_gpu.py)")
args = ap.parse_args()
ok_all = True
for dims in ([64, 64, 64], [128, 64, 32]) + (() if args.small else ([512, 512, 512],)):
    n = dims[0] * dims[1] * dims[2]
    g = torch.Generator(device=dev).manual_seed(1234)
    full = torch.rand(n, 2, dtype=torch.float64, device=dev, generator=g) - 0.5     # same on every rank
    ref = torch.empty_like(full)
    plan = s.FftPlan(s.FORWARD, dims, device=local)
    plan.execute(full.data_ptr(), ref.data_ptr(), torch.cuda.current_stream().cuda_stream)
    n_loc = n // world
    mine = full[rank * n_loc:(rank + 1) * n_loc].clone()
    for mode in ("collective", "p2p", "pipe"):
        slab = SlabFft3d(dims, exchange=mode, device=local)
        first = None
        for it in range(5):  # both receive buffers; "pipe": steps 0-1 eager, 2-4 replayed from the recorded CUDA graphs
            out = slab.forward(mine)
            shards = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(shards, out.contiguous())
            got = gather_transposed(shards, dims).reshape(n, 2)
            if first is None:
                first = got.clone()
            elif not torch.equal(got, first):
                ok_all = False
                if rank == 0:
                    print("dims", dims, "world", world, mode, "step", it, "differs from step 0 FAIL", flush=True)
        if mode == "pipe":
            # a stream of independent transforms with every y pass issued one step ahead (forward_begin): same bits
            slab.forward_begin(mine)
            for it in range(6):
                slab.forward_begin(mine)
                out = slab.forward(mine)
                shards = [torch.empty_like(out) for _ in range(world)]
                dist.all_gather(shards, out.contiguous())
                if not torch.equal(gather_transposed(shards, dims).reshape(n, 2), first):
                    ok_all = False
                    if rank == 0:
                        print("dims", dims, "world", world, "forward_begin step", it, "differs FAIL", flush=True)
            out = slab.forward(mine)  # finishes the last begun transform
            shards = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(shards, out.contiguous())
            got = gather_transposed(shards, dims).reshape(n, 2)
        err = (torch.linalg.norm(got - ref) / torch.linalg.norm(ref)).item()
        exact = torch.equal(got, ref)
        back = slab.inverse(out.clone())
        rt = (torch.linalg.norm(back - mine) / torch.linalg.norm(mine)).item()
        ok = err < 1e-13 and rt < 1e-13
        ok_all &= ok
        if rank == 0:
            print("dims", dims, "world", world, mode, "fwd rel-L2 vs 1-GPU %.2e" % err, "bit-identical", exact,
                  "roundtrip %.2e" % rt, "OK" if ok else "FAIL", flush=True)
        if dims[0] >= 512:  # inverse timing (eager launches), max over ranks
            spec = out.clone()
            for _ in range(3):
                slab.inverse(spec)
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                slab.inverse(spec)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / 20], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("dims", dims, "world", world, mode, "inverse ms %.4f" % t.item(), flush=True)
        torch.cuda.synchronize()
        dist.barrier()
        slab.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
