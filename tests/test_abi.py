"""CPU-side checks of the drop-in boundary: libsst_cuda.so loads, exports every symbol include/sst_cuda.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sst_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sstc_\w+)\s*\(", text)))


def test_header_declares_and_library_exports():
    from shared_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), "libsst_cuda.so does not export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert handle.sstc_abi_version() == 1


def test_lengths_follow_plan_rules():
    # native/src/sharedx/Plan.cpp:230-320
    import shared_b200 as s
    assert s.transform_lengths(s.R2C, [4, 6]) == (24, 32)
    assert s.transform_lengths(s.C2R, [4, 7]) == (32, 28)
    assert s.transform_lengths(s.FORWARD, [3, 3, 3]) == (54, 54)
    assert s.transform_lengths(s.BACKWARD, [1024, 1024, 1024]) == (2 ** 31, 2 ** 31)  # beyond Java's int
    with pytest.raises(s.SstCudaError, match="Invalid dimensions"):
        s.transform_lengths(s.FORWARD, [4, 0])
    with pytest.raises(s.SstCudaError, match="Transform type not recognized"):
        s.transform_lengths(7, [4])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import shared_b200 as s
    with pytest.raises(s.SstCudaError, match="no CPU fallback"):
        s.CudaFftService()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "shared_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".java")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle|oracle/|libsst_oracle|libsst_ref", src, re.M), f
