#!/usr/bin/env python3
"""Harvest the literal known-answer arrays of the reference's own JUnit tests into JSON fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/extract_kats.py

For every `@Test` method of the listed files it records, in source order, each array literal
`new double[] {...}` / `new int[] {...}` together with the wrapper constructor it is an argument of
(`new RealArray(<literal>, [IndexingOrder.X,] d0, d1, ...)` etc.) -> {"cls", "type", "values", "order", "dims"}.
Nothing is interpreted: the tests under tests/ say which record is an input and which the expected output,
citing the reference line of the assertion. Output: tests/golden/<TestClass>.json.
"""
import json
import os
import re
import sys

REF = "/root/reference/src_test/org/shared/test"
FILES = [
    "fft/ArrayFftTest.java",
    "fft/ConvolutionCacheTest.java",
    "array/ArrayKernelTest.java",
    "array/RealArrayTest.java",
    "array/MatrixTest.java",
    "array/IntegerArrayTest.java",
]

METHOD_RE = re.compile(r"public void (test\w+)\(\)")
LITERAL_RE = re.compile(r"new (double|int)\[\]\s*\{")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def parse_number(tok, typ):
    tok = tok.strip()
    if typ == "int":
        return int(tok)
    tok = tok.rstrip("dDfF")
    if tok in ("Double.POSITIVE_INFINITY",):
        return float("inf")
    return float(tok)


def split_args(s):
    """Split a constructor tail `, IndexingOrder.FAR, 3, 4, 5` on top-level commas."""
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


def extract(path):
    raw = open(path).read()
    text = strip_comments(raw)
    methods = [(m.start(), m.group(1)) for m in METHOD_RE.finditer(text)]
    res = {}
    for i, (pos, name) in enumerate(methods):
        end = methods[i + 1][0] if i + 1 < len(methods) else len(text)
        body = text[pos:end]
        recs = []
        for m in LITERAL_RE.finditer(body):
            typ = m.group(1)
            close = body.index("}", m.end())
            toks = [t for t in body[m.end():close].replace("\n", " ").split(",") if t.strip()]
            try:
                values = [parse_number(t, typ) for t in toks]
            except ValueError:
                continue  # expression-valued literal (e.g. Math.sqrt(...)); not a KAT vector
            rec = {"type": typ, "values": values, "cls": None, "order": None, "dims": None,
                   "line": text.count("\n", 0, pos + m.start()) + 1}
            head = body[:m.start()].rstrip()
            wm = re.search(r"new (\w+)\($", head)
            if wm:
                rec["cls"] = wm.group(1)
                # find the end of this constructor call
                depth, j = 1, close + 1
                while depth and j < len(body):
                    if body[j] == "(":
                        depth += 1
                    elif body[j] == ")":
                        depth -= 1
                    j += 1
                tail = split_args(body[close + 1:j - 1])
                dims = []
                for a in tail:
                    if a.startswith("IndexingOrder."):
                        rec["order"] = a.split(".")[1]
                    elif re.fullmatch(r"-?\d+", a):
                        dims.append(int(a))
                rec["dims"] = dims or None
            recs.append(rec)
        res[name] = recs
    return res


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for rel in FILES:
        data = extract(os.path.join(REF, rel))
        name = os.path.splitext(os.path.basename(rel))[0]
        out = os.path.join(here, name + ".json")
        with open(out, "w") as f:
            json.dump({"source": "src_test/org/shared/test/" + rel, "methods": data}, f, separators=(",", ":"))
        print(name, {k: len(v) for k, v in data.items()}, file=sys.stderr)


if __name__ == "__main__":
    main()
