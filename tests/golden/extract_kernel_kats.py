#!/usr/bin/env python3
"""Harvest the call-level known-answer tests of the reference's ArrayKernelTest into a JSON fixture.

    python tests/golden/extract_kernel_kats.py     (build container only: needs /root/reference)

src_test/org/shared/test/array/ArrayKernelTest.java:61-378 is a flat list of
`kernel.<op>(ArrayKernel.<OPCODE>, scalars..., arrays...)` calls, each followed by an assertion against a
literal expected array (or, for raOp/caOp, wrapping the call). This script records every such call:
{"op", "opcode", "scalars", "arrays": [...], "expected": [...] | null, "line"}. Java constant expressions
(Math.exp(1), Math.E * Math.E, Double.NaN ...) are evaluated here, once, with Python's math module (IEEE
double, same libm contract as java.lang.Math to within 1 ulp; the reference's own tolerance is 1e-8).
Calls whose outcome the reference only checks statistically (RND, SHUFFLE) get expected = null.
"""
import json
import math
import os
import re

SRC = "/root/reference/src_test/org/shared/test/array/ArrayKernelTest.java"


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def jeval(expr):
    e = expr.strip()
    e = re.sub(r"\bMath\.E\b", "math.e", e)
    e = re.sub(r"\bMath\.PI\b", "math.pi", e)
    e = re.sub(r"\bMath\.(exp|sqrt|log|cos|sin|atan|pow|abs)\b", r"math.\1", e)
    e = e.replace("math.abs", "abs")
    e = re.sub(r"\bDouble\.NaN\b", "float('nan')", e)
    e = re.sub(r"\bDouble\.POSITIVE_INFINITY\b", "float('inf')", e)
    e = re.sub(r"\bDouble\.NEGATIVE_INFINITY\b", "float('-inf')", e)
    e = re.sub(r"\bDouble\.MAX_VALUE\b", "1.7976931348623157e308", e)
    e = re.sub(r"(?<![\w.])(\d+\.?\d*(?:[eE][-+]?\d+)?)[dDfF]\b", r"\1", e)
    return eval(e, {"math": math, "__builtins__": {"float": float, "abs": abs}})


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def matching(text, i, open_ch="(", close_ch=")"):
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_ch:
            depth += 1
        elif text[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced")


LIT = re.compile(r"new (double|int)\[\]\s*\{")


def parse_literal(text, m):
    typ = m.group(1)
    close = matching(text, m.end() - 1, "{", "}")
    vals = [jeval(t) for t in split_top(text[m.end():close])]
    if typ == "int":
        vals = [int(v) for v in vals]
    return {"type": typ, "values": vals}, close


def main():
    raw = open(SRC).read()
    text = strip_comments(raw)
    start = text.index("public void testOperations()")
    body = text[start:]
    calls = [m for m in re.finditer(r"kernel\.(\w+)\(", body)]
    cases = []
    for idx, m in enumerate(calls):
        op = m.group(1)
        if op in ("randomize", "derandomize"):
            continue
        close = matching(body, m.end() - 1)
        args = split_top(body[m.end():close])
        seg_end = calls[idx + 1].start() if idx + 1 < len(calls) else len(body)
        case = {"op": op, "opcode": None, "scalars": [], "arrays": [], "expected": None,
                "line": text.count("\n", 0, start + m.start()) + 1}
        for a in args:
            a2 = re.sub(r"^\s*\w+\s*=\s*", "", a)  # drop `v = `
            lm = LIT.match(a2)
            if lm:
                try:
                    lit, _ = parse_literal(a2, lm)
                except NameError:
                    lit = {"expr": a2}
                case["arrays"].append(lit)
            elif a2.startswith("ArrayKernel."):
                case["opcode"] = a2.split(".")[1]
            elif a2 in ("true", "false"):
                case["scalars"].append(a2 == "true")
            elif re.fullmatch(r"[A-Za-z_]\w*", a2):
                case["arrays"].append({"ref": a2})  # a variable defined earlier (not used by the KAT calls)
            else:
                try:
                    case["scalars"].append(jeval(a2))
                except Exception:
                    case["arrays"].append({"expr": a2})  # e.g. `new double[nSamples]` of the statistical RND checks
        # expected: the first literal after the call, inside the same assertion segment
        tail = body[close:seg_end]
        if op in ("raOp",):
            em = re.search(r"-\s*([^)]+?)\)\s*<\s*1e-", tail)
            if em:
                case["expected"] = {"type": "double", "values": [jeval(em.group(1))]}
        else:
            lm = LIT.search(tail)
            if lm and "Assert" in body[m.start() - 200:close + lm.start() + 1 + len(tail[:lm.start()])]:
                try:
                    lit, _ = parse_literal(tail, lm)
                    case["expected"] = lit
                except NameError:
                    pass
        if case["opcode"] and ("RND" in case["opcode"] or "SHUFFLE" in case["opcode"]):
            case["expected"] = None
        cases.append(case)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "ArrayKernelCalls.json"), "w") as f:
        json.dump({"source": "src_test/org/shared/test/array/ArrayKernelTest.java", "cases": cases}, f,
                  separators=(",", ":"))
    for c in cases:
        print(c["line"], c["op"], c["opcode"], c["scalars"], [len(a.get("values", [])) for a in c["arrays"]],
              None if c["expected"] is None else len(c["expected"]["values"]))


if __name__ == "__main__":
    main()
