"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; shared_b200/ never does (the product path is CUDA-only and fails loudly without its extension).

`fft` / `ifft` / `rfft` / `rifft` run oracle/fftops_oracle.c, the C restatement of the reference's
FftOps.java / JavaFftService.java (see that file's header for the line map). `ref()` returns the reference's
own native ArrayKernel (oracle/_ref/libsst_ref.so, see oracle/ref/ref_driver.cpp) or None if it was not built.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)


def build(verbose=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building the checker is not using it."""
    out = subprocess.run(["make", "-s", "-f", os.path.join(_HERE, "Makefile"), "all"], cwd=_HERE,
                         capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout + out.stderr)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsst_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.sst_oracle_fft.argtypes = [ctypes.c_int, ctypes.c_int, _c_int_p, _c_dbl_p, _c_dbl_p]
        _LIB.sst_oracle_rfft.argtypes = [ctypes.c_int, _c_int_p, _c_dbl_p, _c_dbl_p]
        _LIB.sst_oracle_rifft.argtypes = [ctypes.c_int, _c_int_p, _c_dbl_p, _c_dbl_p]
        _LIB.sst_oracle_factors.argtypes = [ctypes.c_int, _c_int_p]
    return _LIB


def _dims(dims):
    arr = (ctypes.c_int * len(dims))(*[int(d) for d in dims])
    return arr


def _ptr(a):
    return a.ctypes.data_as(_c_dbl_p)


def factors(n):
    """FftOps.createFactors (FftOps.java:111-150): [p0, m0, p1, m1, ...]."""
    buf = (ctypes.c_int * 128)()
    cnt = lib().sst_oracle_factors(int(n), buf)
    return list(buf[:cnt])


def _c2c(direction, dims, values):
    """`values`: flat float64 array of 2*prod(dims) interleaved re/im. Returns a new flat float64 array."""
    values = np.ascontiguousarray(values, dtype=np.float64).ravel()
    n = int(np.prod(dims))
    assert values.size == 2 * n, "Input and/or output arrays do not have expected sizes"
    out = np.empty_like(values)
    rc = lib().sst_oracle_fft(direction, len(dims), _dims(dims), _ptr(values), _ptr(out))
    if rc != 0:
        raise RuntimeError("oracle fft failed")
    return out


def fft(dims, values):
    """JavaFftService.fft (JavaFftService.java:87-89)."""
    return _c2c(+1, dims, values)


def ifft(dims, values):
    """JavaFftService.ifft (JavaFftService.java:92-94)."""
    return _c2c(-1, dims, values)


def rfft(dims, values):
    """JavaFftService.rfft (JavaFftService.java:53-64). dims = real logical dims."""
    values = np.ascontiguousarray(values, dtype=np.float64).ravel()
    n = int(np.prod(dims))
    assert values.size == n
    out = np.empty(2 * (n // dims[-1]) * (dims[-1] // 2 + 1), dtype=np.float64)
    rc = lib().sst_oracle_rfft(len(dims), _dims(dims), _ptr(values), _ptr(out))
    if rc != 0:
        raise RuntimeError("oracle rfft failed")
    return out


def rifft(dims, values):
    """JavaFftService.rifft (JavaFftService.java:67-84). dims = real logical dims of the result."""
    values = np.ascontiguousarray(values, dtype=np.float64).ravel()
    n = int(np.prod(dims))
    assert values.size == 2 * (n // dims[-1]) * (dims[-1] // 2 + 1)
    out = np.empty(n, dtype=np.float64)
    rc = lib().sst_oracle_rifft(len(dims), _dims(dims), _ptr(values), _ptr(out))
    if rc != 0:
        raise RuntimeError("oracle rifft failed")
    return out


def ref():
    """The reference's own native ArrayKernel behind flat C wrappers, or None when oracle/_ref was not built."""
    from . import refkernel
    return refkernel.load()
