"""Times the SURVEY.md section-8 f-4 members (eigs, insertSparse, sliceSparse) through the host tier of the C ABI on
one B200, next to the reference's own native code (oracle/_ref, one host core) on the same inputs. These are
correctness-first kernels off the hot path; the numbers are for the record, not a target.
Prints one JSON object. Usage: python tools/bench_linalg_sparse.py [--cpu]   (--cpu: also time the reference and compare bits)"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from shared_b200.array_kernel import CudaArrayKernel  # noqa: E402
from sparse_cases import dim_offsets, empty_state, row_major  # noqa: E402

k = CudaArrayKernel(0)
ref = None
if "--cpu" in sys.argv:  # the CPU baseline leg (like bench.py's cpu_baseline): the reference's native code beside the GPU
    import oracle
    ref = oracle.ref()
rng = np.random.default_rng(1)
out = {"eigs": [], "sparse": []}


def best(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


for n in (64, 128, 158, 256):
    a = rng.uniform(0, 1, (n, n)).ravel()
    vec, val = np.zeros(n * n), np.zeros(2 * n)
    rec = {"size": n, "gpu_ms": round(best(lambda: k.eigs(a, vec, val, n), 3), 3),
           "layout": "shared memory" if n * (n | 1) * 8 <= 200 * 1024 else "HBM"}
    if ref is not None:
        rv, rl = np.zeros(n * n), np.zeros(2 * n)
        rec["cpu_ref_ms"] = round(best(lambda: ref.eigs(a, rv, rl, n), 2), 3)
        rec["bit_identical"] = bool(np.array_equal(rv.view(np.int64), vec.view(np.int64)))
    out["eigs"].append(rec)

for dims, n_old, n_new in (([128, 128, 128], 500_000, 500_000),):
    total = int(np.prod(dims))
    S, Do = row_major(dims), dim_offsets(dims)
    old_phys = np.sort(rng.choice(total, n_old, replace=False)).astype(np.int32)
    new_phys = rng.choice(total, n_new, replace=False).astype(np.int32)
    old_vals, new_vals = rng.uniform(-1, 1, n_old), rng.uniform(-1, 1, n_new)
    e = empty_state(dims, np.float64)
    old = k.insertSparse(e[0], dims, S, Do, e[1], old_vals, old_phys.copy())
    state = [None]

    def gpu_insert():
        state[0] = k.insertSparse(old[0], dims, S, Do, old[1], new_vals, new_phys.copy())
    rec = {"dims": dims, "n_old": n_old, "n_new": n_new, "insert_gpu_ms": round(best(gpu_insert, 2), 2)}
    st = state[0]
    rec["count"] = int(st[1].size)
    slices = np.array([v for d, size in enumerate(dims) for j in range(size) for v in (j, size - 1 - j, d)], dtype=np.int32)

    def gpu_slice():
        state[0] = k.sliceSparse(slices, st[0], dims, S, Do, st[1], st[2], st[3], e[0], dims, S, Do, e[1], e[2], e[3])
    rec["slice_reverse_all_gpu_ms"] = round(best(gpu_slice, 2), 2)
    if ref is not None:
        r = [None]

        def cpu_insert():
            r[0] = ref.insertSparse(old[0], dims, S, Do, old[1], new_vals, new_phys.copy())
        rec["insert_cpu_ref_ms"] = round(best(cpu_insert, 1), 2)
        rec["insert_bit_exact"] = all(np.array_equal(a, b) for a, b in zip(r[0], st))
        want = state[0]

        def cpu_slice():
            r[0] = ref.sliceSparse(slices, st[0], dims, S, Do, st[1], st[2], st[3], e[0], dims, S, Do, e[1], e[2], e[3])
        rec["slice_reverse_all_cpu_ref_ms"] = round(best(cpu_slice, 1), 2)
        rec["slice_bit_exact"] = all(np.array_equal(a, b) for a, b in zip(r[0], want))
    out["sparse"].append(rec)

print(json.dumps(out))
