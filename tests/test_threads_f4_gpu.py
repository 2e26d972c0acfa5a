"""eigs / insertSparse / sliceSparse called from several threads at once
(per-thread scratch and result buffers; SURVEY.md 8b "Threading"). Every result must equal the one computed alone."""
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eigs_cases import cases as eigs_cases  # noqa: E402
from sparse_cases import insert_cases, run_insert, run_slice, slice_cases  # noqa: E402

pytestmark = pytest.mark.gpu


def test_concurrent_eigs_and_sparse_calls():
    from shared_b200.array_kernel import CudaArrayKernel
    k = CudaArrayKernel()
    mats = dict(eigs_cases())

    def job_eigs(name):
        a = mats[name]
        n = a.shape[0]

        def run():
            vec, val = np.zeros(n * n), np.zeros(2 * n)
            k.eigs(a.ravel().copy(), vec, val, n)
            return np.concatenate([vec, val])
        return run

    def job_insert(c):
        return lambda: np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in run_insert(k, c)])

    def job_slice(c):
        return lambda: np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in run_slice(k, c)])

    jobs = [job_eigs("uniform33"), job_eigs("uniform64"), job_insert(insert_cases()[6]), job_insert(insert_cases()[20]),
            job_slice(slice_cases()[4]), job_slice(slice_cases()[22])]
    alone = [j() for j in jobs]
    errors = []

    def worker(i):
        try:
            for _ in range(8):
                got = jobs[i]()
                if not np.array_equal(got, alone[i], equal_nan=True):
                    errors.append((i, "mismatch"))
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
