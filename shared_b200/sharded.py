"""Slab-decomposed 3-D complex128 FFT over the GPUs of one box (one process per GPU).

The reference has no distributed path; this is the one exchange step the
north star adds. Layout contract, for dims (d0, d1, d2) over P ranks (P divides d0 and d1):

    natural slabs      rank g holds x[g*d0/P:(g+1)*d0/P, :, :]            (a contiguous range of the flat array)
    transposed slabs   rank g holds X[:, g*d1/P:(g+1)*d1/P, :]            (like FFTW-MPI TRANSPOSED_OUT)

forward(): natural -> transposed.  z pass and y pass are local to a slab; the y pass does not store its
result locally but scatters block q of the y axis straight to rank q (sstc_fft_axis_blocked_dev), so pack,
transfer and unpack are part of the FFT kernel's store; the x pass then runs on the received
(d0, d1/P, d2) array. inverse(): transposed -> natural, mirrored (x pass scatters x blocks to their owners).

Three exchange modes:
    "pipe"        (default on GPUs) the y pass runs first, locally; the z pass is the exchange pass: it stores whole
                  z lines straight into the owner's receive buffer over NVLink (CUDA IPC mapping), in G groups of
                  y indices. After each group a flag barrier kernel runs on a side stream and the x pass of that
                  group's columns starts on a third stream, so the x pass (HBM-bound) overlaps the remaining
                  exchange (NVLink-bound). The schedule itself (streams, events, CUDA graphs) lives behind the C ABI
                  (sstc_slab_*, shared_b200/csrc/fft_slab.cu); this class only maps the peers' buffers and calls it.
    "p2p"         z pass, then the y pass scatters 128-byte row segments into the peers' receive buffers; one
                  cross-rank barrier, then the x pass. Also implements inverse().
    "collective"  destination blocks are a local send buffer; torch.distributed all_to_all_single moves them
                  (NCCL on GPUs, gloo in the CPU tests). Baseline and fallback for the host-logic tests.

The local passes run through an engine: `CudaEngine` (the C ABI) here; the gloo tests under tests/ plug a
numpy engine into the same host logic.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class CudaEngine:
    """Axis passes on CUDA tensors / raw device pointers through libsst_cuda.so."""

    def __init__(self, device):
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        _lib.init(self.device.index or 0)

    def empty(self, n_complex):
        return torch.empty((int(n_complex), 2), dtype=torch.float64, device=self.device)

    @staticmethod
    def _ptr(x):
        return int(x.data_ptr()) if isinstance(x, torch.Tensor) else int(x)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def axis(self, src, dst, outer, length, inner, inverse=False, scale=1.0):
        _lib.check(_lib.lib().sstc_fft_axis_dev(
            ctypes.c_void_p(self._ptr(src)), ctypes.c_void_p(self._ptr(dst)), int(outer), int(length), int(inner),
            int(bool(inverse)), float(scale), ctypes.c_void_p(self._stream())))

    def axis_blocked(self, src_blocks, src_pitch, dst_blocks, dst_pitch, outer, length, inner, inverse=False,
                     scale=1.0):
        sp = (ctypes.c_void_p * len(src_blocks))(*[self._ptr(b) for b in src_blocks])
        dp = (ctypes.c_void_p * len(dst_blocks))(*[self._ptr(b) for b in dst_blocks])
        _lib.check(_lib.lib().sstc_fft_axis_blocked_dev(
            sp, len(src_blocks), int(src_pitch), dp, len(dst_blocks), int(dst_pitch), int(outer), int(length),
            int(inner), int(bool(inverse)), float(scale), ctypes.c_void_p(self._stream())))


    def axis_strided(self, src, dst, outer, length, inner, in_plane, in_kstride, out_plane, out_kstride,
                     inverse=False, scale=1.0):
        _lib.check(_lib.lib().sstc_fft_axis_strided_dev(
            ctypes.c_void_p(self._ptr(src)), ctypes.c_void_p(self._ptr(dst)), int(outer), int(length), int(inner),
            int(in_plane), int(in_kstride), int(out_plane), int(out_kstride), int(bool(inverse)), float(scale),
            ctypes.c_void_p(self._stream())))

    def lines_scatter(self, src, dst_blocks, planes, plane_lines, length, first, group, inverse=False, scale=1.0,
                      max_ctas=0):
        dp = (ctypes.c_void_p * len(dst_blocks))(*[self._ptr(b) for b in dst_blocks])
        _lib.check(_lib.lib().sstc_fft_lines_scatter_dev(
            ctypes.c_void_p(self._ptr(src)), dp, len(dst_blocks), int(planes), int(plane_lines), int(length),
            int(first), int(group), int(bool(inverse)), float(scale), int(max_ctas),
            ctypes.c_void_p(self._stream())))

    def planes_scatter(self, src, dst_blocks, block_planes, plane_lines, dst_plane_lines, length, first, group,
                       inverse=False, scale=1.0, max_ctas=0):
        dp = (ctypes.c_void_p * len(dst_blocks))(*[self._ptr(b) for b in dst_blocks])
        _lib.check(_lib.lib().sstc_fft_planes_scatter_dev(
            ctypes.c_void_p(self._ptr(src)), dp, len(dst_blocks), int(block_planes), int(plane_lines),
            int(dst_plane_lines), int(length), int(first), int(group), int(bool(inverse)), float(scale), int(max_ctas),
            ctypes.c_void_p(self._stream())))

    def peer_barrier(self, flag_ptrs, rank, epoch):
        fp = (ctypes.c_void_p * len(flag_ptrs))(*[int(p) for p in flag_ptrs])
        _lib.check(_lib.lib().sstc_peer_barrier_dev(fp, len(flag_ptrs), int(rank), ctypes.c_uint64(int(epoch)),
                                                    ctypes.c_void_p(self._stream())))


class _DeviceArray:
    """A raw device allocation seen by torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n_complex):
        self.__cuda_array_interface__ = {"shape": (int(n_complex), 2), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3}


class PeerBuffer:
    """A cudaMalloc'd buffer of n complex128 values that every rank of `group` maps (CUDA IPC)."""

    def __init__(self, n_complex, device, group):
        lib = _lib.lib()
        self.n = int(n_complex)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        p = ctypes.c_void_p()
        _lib.check(lib.sstc_malloc(ctypes.byref(p), self.n * 16))
        self.local_ptr = p.value
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.sstc_ipc_export(ctypes.c_void_p(self.local_ptr), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
            else:
                q = ctypes.c_void_p()
                _lib.check(lib.sstc_ipc_open(h, ctypes.byref(q)))
                self.ptrs.append(q.value)
        self.tensor = torch.as_tensor(_DeviceArray(self.local_ptr, self.n), device=device)

    def close(self):
        lib = _lib.lib()
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                lib.sstc_ipc_close(ctypes.c_void_p(p))
        if self.local_ptr:
            lib.sstc_free(ctypes.c_void_p(self.local_ptr))
        self.ptrs, self.local_ptr = [], None


def block_ptrs(base, n_blocks, block_elems):
    """Pointers (or tensor views) of n consecutive blocks of `block_elems` complex values inside `base`."""
    if isinstance(base, torch.Tensor):
        return [base[q * block_elems:(q + 1) * block_elems] for q in range(n_blocks)]
    return [int(base) + q * block_elems * 16 for q in range(n_blocks)]


class SlabFft3d:
    """3-D c2c transform of (d0, d1, d2), slab-sharded over the ranks of `group`."""

    def __init__(self, dims, group=None, exchange="pipe", engine=None, device=None, groups=None,
                 exchange_ctas=None, chunks=None, bulk_store=None, graph=None, order=None, head_ctas=None,
                 split_barrier=None):
        if len(dims) != 3:
            raise ValueError("SlabFft3d needs three dimensions")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.d0, self.d1, self.d2 = (int(d) for d in dims)
        p = self.world
        if self.d0 % p or self.d1 % p or p > 8:
            raise ValueError("the rank count must divide d0 and d1 and be at most 8")
        self.x_loc, self.y_loc = self.d0 // p, self.d1 // p
        self.n_local = self.x_loc * self.d1 * self.d2          # == d0 * y_loc * d2
        self.block = self.x_loc * self.y_loc * self.d2         # complex values per pairwise block
        self.n_total = self.d0 * self.d1 * self.d2
        self.exchange = exchange
        self.engine = engine if engine is not None else CudaEngine(device if device is not None else 0)
        self.work = self.engine.empty(self.n_local)
        self._step = 0
        self.replayed_kernels = 0  # kernels launched through CUDA-graph replays (not seen by sstc_launch_count)
        self._slab = None
        if exchange == "pipe" and not (8 <= self.d2 <= 4096 and self.d2 & (self.d2 - 1) == 0):
            exchange = self.exchange = "p2p"  # the line scatter needs a power-of-two d2 the tile kernels hold
        if exchange in ("p2p", "pipe"):
            dev = self.engine.device
            self.peer = [PeerBuffer(self.n_local, dev, group) for _ in range(2)]
            self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
            self.flags = PeerBuffer(16, dev, group)  # 8 uint64 epoch slots per rank (+ the rank's own call counter)
            self.flags.tensor.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier(group)
            if exchange == "pipe":
                lib = _lib.lib()
                arr = lambda ptrs: (ctypes.c_void_p * len(ptrs))(*[int(q) for q in ptrs])
                h = ctypes.c_void_p()
                _lib.check(lib.sstc_slab_create(ctypes.byref(h), _lib.i32_array([self.d0, self.d1, self.d2]), p, self.rank,
                                                arr(self.peer[0].ptrs), arr(self.peer[1].ptrs), arr(self.flags.ptrs)))
                self._slab = h
                for name, value in (("groups", groups), ("exchange_ctas", exchange_ctas), ("chunks", chunks),
                                    ("bulk_store", bulk_store), ("graph", graph), ("order", order), ("head_ctas", head_ctas),
                                    ("split_barrier", split_barrier)):
                    if value is not None:
                        _lib.check(lib.sstc_slab_set_option(h, name.encode(), int(value)))
                torch.cuda.synchronize(dev)
                dist.barrier(group)
        elif exchange == "collective":
            self.send = self.engine.empty(self.n_local)
            self.recv = self.engine.empty(self.n_local)
        else:
            raise ValueError("exchange must be 'pipe', 'p2p' or 'collective'")

    # -- cross-rank ordering ----------------------------------------------------------------------------
    def _barrier(self):
        # stream-ordered: the collective starts after this rank's scatter kernel and the next pass starts
        # after every rank has joined, i.e. after every peer's stores into this rank's buffer are complete.
        dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)

    # -- transforms ---------------------------------------------------------------------------------------
    def _run_slab(self, fn, x):
        out = ctypes.c_void_p()
        stream = torch.cuda.current_stream(self.engine.device).cuda_stream
        _lib.check(fn(self._slab, ctypes.c_void_p(int(x.data_ptr())), ctypes.byref(out), ctypes.c_void_p(stream)))
        for b in self.peer:
            if b.local_ptr == out.value:
                return b.tensor
        raise RuntimeError("sstc_slab returned a pointer outside its receive buffers")

    def forward_begin(self, x):
        """Issues the local y pass of a LATER forward(x) now (sstc_slab_forward_begin): in a loop over independent
        transforms, `forward_begin(next); out = forward(current)` hides every step's head behind its predecessor's
        exchange. `x` must stay unchanged until the matching forward."""
        if self._slab is None:
            raise ValueError("forward_begin() is for the 'pipe' exchange")
        stream = torch.cuda.current_stream(self.engine.device).cuda_stream
        _lib.check(_lib.lib().sstc_slab_forward_begin(self._slab, ctypes.c_void_p(int(x.data_ptr())), ctypes.c_void_p(stream)))

    def capture(self, x):
        """Two eager steps with input buffer `x` (one per receive buffer): from the third call on, the C side replays
        the schedule from a CUDA graph (sstc_slab_forward records one per input pointer and buffer parity)."""
        if self.exchange != "pipe":
            raise ValueError("capture() is for the 'pipe' exchange")
        for _ in range(2):
            self.forward(x)
        torch.cuda.synchronize(self.engine.device)
        dist.barrier(self.group)

    def forward(self, x):
        """x: this rank's natural slab, (n_local, 2) float64. Returns this rank's transposed slab: a view of an
        internal receive buffer, valid until the next-but-one call provided its consumers are enqueued on the current
        stream before the next call (peers write into it again two calls later; include/sst_cuda.h, sstc_slab)."""
        if self.exchange == "pipe":
            return self._run_slab(_lib.lib().sstc_slab_forward, x)
        e, p = self.engine, self.world
        e.axis(x, self.work, self.x_loc * self.d1, self.d2, 1)                                   # z
        if self.exchange == "p2p":
            buf = self.peer[self._step % 2]
            self._step += 1
            dst = [buf.ptrs[q] + self.rank * self.block * 16 for q in range(p)]
            e.axis_blocked([self.work], 0, dst, 0, self.x_loc, self.d1, self.d2)                 # y + transpose
            self._barrier()
            out = buf.tensor
        else:
            e.axis_blocked([self.work], 0, block_ptrs(self.send, p, self.block), 0, self.x_loc, self.d1, self.d2)
            dist.all_to_all_single(self.recv, self.send, group=self.group)
            out = self.recv
        e.axis(out, out, 1, self.d0, self.y_loc * self.d2)                                       # x
        return out

    def inverse(self, y):
        """y: this rank's transposed slab. Returns this rank's natural slab of ifft (scaled by 1/N)."""
        if self.exchange == "pipe":
            return self._run_slab(_lib.lib().sstc_slab_inverse, y)
        e, p = self.engine, self.world
        row = self.d1 * self.d2  # one x row of a natural slab
        if self.exchange in ("p2p", "pipe"):
            buf = self.peer[self._step % 2]
            self._step += 1
            # x block q -> rank q, landing at columns [rank*y_loc*d2, ...) of each of its x rows
            dst = [buf.ptrs[q] + self.rank * self.y_loc * self.d2 * 16 for q in range(p)]
            e.axis_blocked([y], 0, dst, row, 1, self.d0, self.y_loc * self.d2, inverse=True)    # x + transpose
            self._barrier()
            slab = buf.tensor
            e.axis(slab, slab, self.x_loc, self.d1, self.d2, inverse=True)                       # y
        else:
            e.axis_blocked([y], 0, block_ptrs(self.send, p, self.block), 0, 1, self.d0, self.y_loc * self.d2,
                           inverse=True)
            dist.all_to_all_single(self.recv, self.send, group=self.group)
            # the y axis arrives as one (x_loc, y_loc, d2) block per source rank
            slab = self.work
            e.axis_blocked(block_ptrs(self.recv, p, self.block), 0, [slab], 0, self.x_loc, self.d1, self.d2,
                           inverse=True)
        e.axis(slab, slab, self.x_loc * self.d1, self.d2, 1, inverse=True, scale=1.0 / self.n_total)  # z
        return slab

    def set_option(self, name, value):
        _lib.check(_lib.lib().sstc_slab_set_option(self._slab, name.encode(), int(value)))

    def trace(self):
        """[(label, ms since the step's start)] of the last step issued with option "trace" = 1."""
        buf = ctypes.create_string_buffer(1 << 16)
        n = _lib.lib().sstc_slab_trace(self._slab, buf, len(buf))
        return [(a, float(b)) for a, b in (ln.split("\t") for ln in buf.raw[:n].decode().splitlines())]

    def graph_count(self):
        """Recorded CUDA graphs on the C side (0 while the schedule is still issued eagerly)."""
        return int(_lib.lib().sstc_slab_graph_count(self._slab)) if self._slab is not None else 0

    def launches_per_forward(self):
        return 3

    def close(self):
        if self._slab is not None:
            torch.cuda.synchronize(self.engine.device)
            _lib.lib().sstc_slab_destroy(self._slab)
            self._slab = None
        if self.exchange in ("p2p", "pipe"):
            for b in self.peer + [self.flags]:
                b.close()


def gather_transposed(shards, dims):
    """Reassemble the full (d0, d1, d2, 2) array from the ranks' transposed slabs (test / verification helper)."""
    d0, d1, d2 = dims
    p = len(shards)
    return torch.cat([s.reshape(d0, d1 // p, d2, 2) for s in shards], dim=1)


class ShardedFft3d:
    """The slab transform driven from ONE process over several devices (sstc_fft3d_sharded_*): what a JVM-side
    CudaFftService binds when setHint("devices", ...) names more than one GPU. Device tier: lists of per-device slabs;
    host tier: whole host arrays in natural order, every device moving its slab over its own PCIe link."""

    def __init__(self, dims, devices):
        self.dims = tuple(int(d) for d in dims)
        self.devices = [int(d) for d in devices]
        _lib.init(self.devices[0])  # sstc_init also names the device NEW threads are bound to: keep it the first one
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().sstc_fft3d_sharded_create(ctypes.byref(h), _lib.i32_array(self.dims), len(self.devices),
                                                        _lib.i32_array(self.devices)))
        self._h = h
        d0, d1, d2 = self.dims
        self.n_local = d0 * d1 * d2 // len(self.devices)

    def set_option(self, name, value):
        _lib.check(_lib.lib().sstc_fft3d_sharded_set_option(self._h, name.encode(), int(value)))

    def _run(self, fn, slabs):
        p = len(self.devices)
        ins = (ctypes.c_void_p * p)(*[int(t.data_ptr()) for t in slabs])
        outs = (ctypes.c_void_p * p)()
        streams = (ctypes.c_void_p * p)(*[torch.cuda.current_stream(torch.device("cuda", d)).cuda_stream
                                          for d in self.devices])
        _lib.check(fn(self._h, ins, outs, streams))
        return [torch.as_tensor(_DeviceArray(outs[r], self.n_local), device=torch.device("cuda", d))
                for r, d in enumerate(self.devices)]

    def forward(self, slabs):
        """slabs[r]: rank r's natural slab on devices[r], (n_local, 2) float64. Returns the transposed slabs (views of
        the plan's receive buffers, valid until the next-but-one call)."""
        return self._run(_lib.lib().sstc_fft3d_sharded_forward, slabs)

    def inverse(self, slabs):
        return self._run(_lib.lib().sstc_fft3d_sharded_inverse, slabs)

    def sync(self):
        _lib.check(_lib.lib().sstc_fft3d_sharded_sync(self._h))

    def host(self, in_, out, inverse=False):
        """Whole-array transform on host arrays (numpy float64, interleaved complex, natural order in and out)."""
        _lib.check(_lib.lib().sstc_fft3d_sharded_host(self._h, int(bool(inverse)), ctypes.c_void_p(in_.ctypes.data),
                                                      ctypes.c_void_p(out.ctypes.data)))

    def close(self):
        if self._h is not None and self._h.value:
            _lib.lib().sstc_fft3d_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
