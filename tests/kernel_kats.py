"""Replays the reference's call-level known-answer tests (tests/golden/ArrayKernelCalls.json, harvested from
src_test/org/shared/test/array/ArrayKernelTest.java:61-378) against any object with the ArrayKernel method
surface: the reference's own native kernel (oracle.ref()) on CPU, the CUDA provider on the GPU."""
import json
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-8  # src_test/org/shared/test/Tests.java:117

# opcode values: src/org/shared/array/kernel/ArrayKernel.java:45-278
OPCODES = {}
_names = """RR_SUM=0 RR_PROD=1 RR_MAX=2 RR_MIN=3 RR_VAR=4 RI_MAX=0 RI_MIN=1 RI_ZERO=2 RI_GZERO=3 RI_LZERO=4 RI_SORT=5
RD_SUM=0 RD_PROD=1 RA_SUM=0 RA_PROD=1 RA_VAR=2 RA_MAX=3 RA_MIN=4 RA_ENT=5 CA_SUM=0 CA_PROD=1
RU_ADD=0 RU_MUL=1 RU_ABS=2 RU_POW=3 RU_EXP=4 RU_RND=5 RU_LOG=6 RU_SQRT=7 RU_SQR=8 RU_INV=9 RU_COS=10 RU_SIN=11
RU_ATAN=12 RU_FILL=13 RU_SHUFFLE=14 CU_ADD=0 CU_MUL=1 CU_EXP=2 CU_RND=3 CU_CONJ=4 CU_COS=5 CU_SIN=6 CU_FILL=7
CU_SHUFFLE=8 IU_ADD=0 IU_MUL=1 IU_FILL=2 IU_SHUFFLE=3 RE_ADD=0 RE_SUB=1 RE_MUL=2 RE_DIV=3 RE_MAX=4 RE_MIN=5
IE_ADD=0 IE_SUB=1 IE_MUL=2 IE_MAX=3 IE_MIN=4 CE_ADD=0 CE_SUB=1 CE_MUL=2 CE_DIV=3 C_TO_R_ABS=0 C_TO_R_RE=1
C_TO_R_IM=2 R_TO_C_RE=0 R_TO_C_IM=1 I_TO_R=0"""
for tok in _names.split():
    k, v = tok.split("=")
    OPCODES[k] = int(v)


def load_cases():
    with open(os.path.join(ROOT, "tests", "golden", "ArrayKernelCalls.json")) as f:
        return [c for c in json.load(f)["cases"] if c["expected"] is not None]


def _arr(rec):
    return np.array(rec["values"], dtype=np.float64 if rec["type"] == "double" else np.int32)


def _new(rec):
    """`new double[3]` / `new int[3]` destination arguments."""
    m = re.match(r"new (double|int)\[(\d+)\]", rec["expr"])
    return np.zeros(int(m.group(2)), dtype=np.float64 if m.group(1) == "double" else np.int32)


def replay(kernel, case):
    """Runs one recorded call; returns (got, expected) as numpy arrays."""
    op, code, sc = case["op"], OPCODES[case["opcode"]], case["scalars"]
    exp = _arr(case["expected"])
    arrays = [(_arr(a) if "values" in a else _new(a)) for a in case["arrays"]]
    if op == "raOp":
        return np.array([kernel.raOp(code, arrays[0])]), exp
    if op == "caOp":
        return np.asarray(kernel.caOp(code, arrays[0])), exp
    if op == "ruOp":
        kernel.ruOp(code, sc[0], arrays[0])
        return arrays[0], exp
    if op == "cuOp":
        kernel.cuOp(code, sc[0], sc[1], arrays[0])
        return arrays[0], exp
    if op == "iuOp":
        kernel.iuOp(code, int(sc[0]), arrays[0])
        return arrays[0], exp
    if op == "eOp":
        kernel.eOp(code, arrays[0], arrays[1], arrays[2], sc[0])
        return arrays[2], exp
    if op == "convert":
        kernel.convert(code, arrays[0], sc[0], arrays[1], sc[1])
        return arrays[1], exp
    raise AssertionError("unhandled op " + op)


def check_all(kernel):
    cases = load_cases()
    assert len(cases) >= 50
    for case in cases:
        got, exp = replay(kernel, case)
        assert got.shape == exp.shape, (case["opcode"], case["line"])
        if exp.dtype == np.int32:
            assert np.array_equal(got, exp), (case["opcode"], case["line"], got, exp)
        else:
            assert np.max(np.abs(got - exp)) < TOL, (case["opcode"], case["line"], got, exp)
    return len(cases)
