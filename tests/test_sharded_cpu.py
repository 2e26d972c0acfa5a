"""Host logic of the slab-sharded 3-D transform (shared_b200/sharded.py) on CPU: two gloo ranks run the real
SlabFft3d schedule ("collective" exchange) with a numpy stand-in for the per-axis kernels, and the gathered
result must equal numpy's fftn / ifftn. This covers the layout contract (natural slabs in, transposed slabs
out, block addressing on both sides of the all-to-all); the CUDA kernels themselves are covered by
tests/test_fft_gpu.py and, on 2+ GPUs, tools/check_sharded.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyEngine:
    """Test-only engine: the semantics of sstc_fft_axis_dev / sstc_fft_axis_blocked_dev on CPU tensors."""

    device = torch.device("cpu")

    def empty(self, n_complex):
        return torch.zeros((int(n_complex), 2), dtype=torch.float64)

    @staticmethod
    def _c(t):
        return t.numpy().view(np.complex128)[..., 0]

    def axis(self, src, dst, outer, length, inner, inverse=False, scale=1.0):
        x = self._c(src).reshape(outer, length, inner)
        y = (np.fft.ifft(x, axis=1) * length if inverse else np.fft.fft(x, axis=1)) * scale
        self._c(dst)[:] = y.reshape(-1)

    def axis_blocked(self, src_blocks, src_pitch, dst_blocks, dst_pitch, outer, length, inner, inverse=False,
                     scale=1.0):
        assert src_pitch in (0, inner) and dst_pitch in (0, inner), "the collective path uses dense blocks"
        ks = length // len(src_blocks)
        x = np.concatenate([self._c(b).reshape(outer, ks, inner) for b in src_blocks], axis=1)
        y = (np.fft.ifft(x, axis=1) * length if inverse else np.fft.fft(x, axis=1)) * scale
        kd = length // len(dst_blocks)
        for q, b in enumerate(dst_blocks):
            self._c(b)[:] = y[:, q * kd:(q + 1) * kd, :].reshape(-1)


def _worker(rank, world, port, dims, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shared_b200.sharded import SlabFft3d, gather_transposed
        n = dims[0] * dims[1] * dims[2]
        full = np.random.default_rng(5).uniform(-0.5, 0.5, (n, 2))
        n_loc = n // world
        mine = torch.from_numpy(full[rank * n_loc:(rank + 1) * n_loc].copy())
        slab = SlabFft3d(dims, exchange="collective", engine=NumpyEngine())
        y = slab.forward(mine).clone()
        shards = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(shards, y)
        got = gather_transposed(shards, dims).numpy().view(np.complex128)[..., 0]
        want = np.fft.fftn(full.view(np.complex128)[:, 0].reshape(dims))
        err_f = np.linalg.norm(got - want) / np.linalg.norm(want)
        back = slab.inverse(y).numpy()
        err_b = np.linalg.norm(back - mine.numpy()) / np.linalg.norm(mine.numpy())
        out[rank] = (float(err_f), float(err_b))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dims", [(4, 6, 5), (8, 4, 16)])
def test_slab_fft_two_ranks_gloo(dims):
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), dims, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err_f, err_b = out[rank]
        assert err_f < 1e-13 and err_b < 1e-13, (rank, err_f, err_b)


def test_block_pointer_arithmetic():
    sys.path.insert(0, ROOT)
    from shared_b200.sharded import block_ptrs
    assert block_ptrs(1000, 3, 10) == [1000, 1160, 1320]  # 16 bytes per complex value
    t = torch.zeros(30, 2, dtype=torch.float64)
    views = block_ptrs(t, 3, 10)
    assert [v.data_ptr() - t.data_ptr() for v in views] == [0, 160, 320]
