"""What a kernel can push into a peer's HBM over NVLink on this box (one process, two or more devices).

    python tools/exp_nvlink.py [--mib 1024] [--quick]  > gpurun_out/exp_nvlink.json

Rates of sstc_exp_peer_copy from device 0 into device 1 (and both directions at once, and local for reference):
per-thread 16-byte stores at several grid sizes against bulk asynchronous copies (cp.async.bulk) of 8-64 KB pieces. The
numbers bound the exchange pass of the slab transform. Under ncu (`--quick`, NVLink metrics) the same
launches give user bytes against protocol bytes on the link."""
import argparse
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
from shared_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--quick", action="store_true", help="one launch per variant (for ncu)")
args = ap.parse_args()
nbytes = args.mib << 20
ndev = torch.cuda.device_count()
assert ndev >= 2, "needs two devices"
for d in range(2):
    _lib.init(d)
lib = _lib.lib()
bufs = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % d) for d in range(2)]
srcs = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device="cuda:%d" % d) for d in range(2)]
for d in range(2):
    with torch.cuda.device(d):
        # peer access both ways (torch enables it lazily on the first cross-device copy)
        bufs[d][:16].copy_(srcs[1 - d][:16])
        torch.cuda.synchronize()


def copy(dst, src, mode, piece, ctas, dev):
    with torch.cuda.device(dev):
        _lib.check(lib.sstc_exp_peer_copy(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), nbytes, mode, piece,
                                          ctas, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def timed(fn, reps):
    fn()
    for d in range(2):
        torch.cuda.synchronize(d)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2)]
    for d in range(2):
        with torch.cuda.device(d):
            ev[d][0].record()
    for _ in range(reps):
        fn()
    for d in range(2):
        with torch.cuda.device(d):
            ev[d][1].record()
    for d in range(2):
        torch.cuda.synchronize(d)
    return max(ev[d][0].elapsed_time(ev[d][1]) for d in range(2)) / reps


reps = 1 if args.quick else 10
out = {"bytes": nbytes, "variants": []}
variants = [("ldst", 0, 0, c) for c in (148, 296, 592, 1184)] + [("bulk", 1, p, c) for p in (8192, 32768, 65536) for c in (148, 296, 444)]
if args.quick:
    variants = [("ldst", 0, 0, 296), ("bulk", 1, 8192, 296), ("bulk", 1, 32768, 296)]
for name, mode, piece, ctas in variants:
    one = timed(lambda: copy(bufs[1], srcs[0], mode, piece, ctas, 0), reps)
    rec = {"variant": name, "piece": piece, "ctas": ctas, "to_peer_gbs": nbytes / one / 1e6}
    if not args.quick:
        both = timed(lambda: (copy(bufs[1], srcs[0], mode, piece, ctas, 0), copy(bufs[0], srcs[1], mode, piece, ctas, 1)), reps)
        local = timed(lambda: copy(bufs[0], srcs[0], mode, piece, ctas, 0), reps)
        rec.update({"both_directions_gbs_each": nbytes / both / 1e6, "local_gbs_read_plus_write": 2 * nbytes / local / 1e6})
        ok = bool(torch.equal(bufs[1].cpu(), srcs[0].cpu())) if ctas == 296 else None
        rec["correct"] = ok
    out["variants"].append(rec)
if not args.quick:
    t = timed(lambda: bufs[1].copy_(srcs[0], non_blocking=True), reps)
    out["cudaMemcpyPeer_gbs"] = nbytes / t / 1e6
print(json.dumps(out))
