"""Any Java thread may call the providers at any time (ModalArrayKernel / ModalFftService are lock-free volatile
delegates, SURVEY.md 8b "Threading"). ctypes releases the GIL during a call, so Python threads exercise the same
concurrency: several threads hammer the host tier with different operations and every result must match the value
computed when the operation ran alone."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_host_tier_calls():
    import shared_b200 as s
    from shared_b200.array_kernel import CudaArrayKernel
    from shared_b200.image_kernel import CudaImageKernel
    k, svc, ik = CudaArrayKernel(), s.CudaFftService(), CudaImageKernel()
    rng = np.random.default_rng(40)

    def job_fft(dims):
        n = int(np.prod(dims))
        x = rng.uniform(-0.5, 0.5, 2 * n)

        def run():
            out = np.empty_like(x)
            svc.fft(dims, x, out)
            return out
        return run

    def job_rfft(dims):
        n = int(np.prod(dims))
        x = rng.uniform(-0.5, 0.5, n)

        def run():
            out = np.empty(2 * (n // dims[-1]) * (dims[-1] // 2 + 1))
            svc.rfft(dims, x, out)
            return out
        return run

    def job_rrop():
        a = rng.uniform(-1, 1, 300 * 257)

        def run():
            out = np.zeros(300)
            k.rrOp(0, a, [300, 257], [257, 1], out, [300, 1], [1, 1], 1)
            return out
        return run

    def job_eop():
        a, b = rng.uniform(-1, 1, 100003), rng.uniform(-1, 1, 100003)

        def run():
            out = np.empty_like(a)
            k.eOp(2, a, b, out, False)
            return out
        return run

    def job_mul():
        a, b = rng.uniform(-1, 1, 130 * 67), rng.uniform(-1, 1, 67 * 90)

        def run():
            out = np.empty(130 * 90)
            k.mul(a, b, 130, 90, out, False)
            return out
        return run

    def job_sort():
        a = rng.permutation(5000).astype(np.float64)

        def run():
            v, perm = a.copy(), np.zeros(5000, dtype=np.int32)
            k.riOp(5, v, [5000], [1], perm, 0)
            return np.concatenate([v, perm.astype(np.float64)])
        return run

    def job_image():
        a = rng.integers(-3, 4, 120 * 77).astype(np.float64)

        def run():
            out = np.zeros(121 * 78)
            ik.createIntegralImage(a, [120, 77], [77, 1], out, [121, 78], [78, 1])
            return out
        return run

    jobs = [job_fft([64, 48]), job_fft([4099]), job_rfft([32, 100]), job_rfft([8192]), job_rrop(), job_eop(), job_mul(),
            job_sort(), job_image()]
    expected = [j() for j in jobs]
    failures = []

    def worker(i):
        try:
            for rep in range(15):
                for off in range(len(jobs)):
                    j = (i + off) % len(jobs)
                    got = jobs[j]()
                    if not np.array_equal(got, expected[j]):
                        failures.append((i, rep, j))
        except Exception as e:  # noqa: BLE001
            failures.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures, failures[:5]
