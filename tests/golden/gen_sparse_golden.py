#!/usr/bin/env python3
"""Generates tests/golden/sparse_golden.npz: the SparseArrayStates the reference's OWN SparseOps::insert / ::slice
(/root/reference/native/src/shared/SparseOps{,Insert,Slice}.cpp, compiled unmodified into oracle/_ref/libsst_ref.so)
return on the seeded inputs of tests/sparse_cases.py. Run in the build container (needs /root/reference):
    python tests/golden/gen_sparse_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from sparse_cases import insert_cases, run_insert, run_slice, slice_cases  # noqa: E402

ref = oracle.ref()
assert ref is not None, "oracle/_ref/libsst_ref.so missing: run `make -C oracle` where /root/reference exists"
out = {}
for n, c in enumerate(insert_cases()):
    for name, a in zip(("values", "indices", "offsets", "indirections", "newI"), run_insert(ref, c)):
        out["insert%d/%s" % (n, name)] = a
for n, c in enumerate(slice_cases()):
    for name, a in zip(("values", "indices", "offsets", "indirections"), run_slice(ref, c)):
        out["slice%d/%s" % (n, name)] = a
np.savez_compressed(os.path.join(HERE, "sparse_golden.npz"), **out)
print("wrote", len(out), "arrays")
