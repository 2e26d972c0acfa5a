#!/usr/bin/env python3
"""Generates tests/golden/eigs_golden.npz: the outputs of the reference's OWN LinearAlgebraOps::eigs
(/root/reference/native/src/shared/LinearAlgebraOpsEigs.cpp, compiled unmodified into oracle/_ref/libsst_ref.so by
oracle/Makefile) on the seeded inputs of tests/eigs_cases.py. Run in the build container (needs /root/reference):
    python tests/golden/gen_eigs_golden.py
The fixture lets the bit-exactness tests run where neither /root/reference nor oracle/_ref exists."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from eigs_cases import cases  # noqa: E402

ref = oracle.ref()
assert ref is not None, "oracle/_ref/libsst_ref.so missing: run `make -C oracle` where /root/reference exists"
out = {}
for name, a in cases():
    n = a.shape[0]
    vec, val = np.zeros(n * n), np.zeros(2 * n)
    ref.eigs(a.ravel().copy(), vec, val, n)
    out[name + "/vec"], out[name + "/val"] = vec, val
np.savez_compressed(os.path.join(HERE, "eigs_golden.npz"), **out)
print("wrote", len(out) // 2, "cases")
