"""One ArrayKernel.eigs call at the reference test's size (128) for `ncu --set full -k regex:eigs_schur_kernel`.
Usage: ncu --set full --clock-control none --import-source on -k regex:eigs_schur_kernel -c 1 -o gpurun_out/prof_eigs \\
       python tools/prof_eigs.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
from shared_b200.array_kernel import CudaArrayKernel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
k = CudaArrayKernel(0)
a = np.random.default_rng(1).uniform(0, 1, (n, n)).ravel()
vec, val = np.zeros(n * n), np.zeros(2 * n)
k.eigs(a, vec, val, n)
print("eigs", n, "ok", float(np.abs(val).max()))
