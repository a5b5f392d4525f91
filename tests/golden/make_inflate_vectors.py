#!/usr/bin/env python3
"""Extract the deterministic decoder vectors of the reference test-suite.

Reads  /root/reference/test/io/nayuki/deflate/InflaterInputStreamTest.java
(the 39 hand-written bit-string cases, lines 24-510) and writes
tests/golden/inflate_vectors.json.  Only the *data* of the vectors (bit strings,
expected output hex, expected Reason) is taken; the harness rules
(pad-independence :522-531, end-exactly :557-558) are re-implemented in
tests/test_oracle_inflate.py.

This script needs /root/reference and therefore only runs in the build
container; the JSON it produces is committed and travels to the GPU box.
"""
import json, re, sys, pathlib

SRC = pathlib.Path("/root/reference/test/io/nayuki/deflate/InflaterInputStreamTest.java")
OUT = pathlib.Path(__file__).with_name("inflate_vectors.json")


def main():
    text = SRC.read_text()
    lines = text.split("\n")
    # locate @Test methods
    methods = []
    for i, ln in enumerate(lines):
        m = re.search(r"public void (test\w+)\(\)", ln)
        if m:
            methods.append((m.group(1), i))
    methods.append(("__end__", len(lines)))
    vectors = []
    for (name, start), (_, end) in zip(methods, methods[1:]):
        body = "\n".join(lines[start:end])
        # strip // comments
        body_nc = re.sub(r"//[^\n]*", "", body)
        # local String variables (dynamic-Huffman cases build the input from named pieces)
        svars = {m.group(1): m.group(2) for m in re.finditer(r"String (\w+) = \"([^\"]*)\";", body_nc)}
        for m in re.finditer(r"\b(testFail|test)\(((?:\s*(?:\"[^\"]*\"|\w+)\s*\+?)+?)\s*,\s*([^;]*?)\);", body_nc, re.S):
            kind = m.group(1)
            pieces = re.findall(r"\"([^\"]*)\"|(\w+)", m.group(2))
            if any(v and v not in svars for _, v in pieces):
                continue  # randomized cases (StringBuilder)
            bits = "".join(lit if not v else svars[v] for lit, v in pieces).replace(" ", "")
            if not re.fullmatch(r"[01]*", bits):
                continue  # randomized cases build strings dynamically
            arg = m.group(3).strip()
            line_no = start + 1
            if kind == "test":
                hexout = "".join(re.findall(r"\"([^\"]*)\"", arg)).replace(" ", "").lower()
                vectors.append(dict(name=name, line=line_no, bits=bits, expect="ok", output_hex=hexout))
            else:
                reason = arg.replace("Reason.", "")
                vectors.append(dict(name=name, line=line_no, bits=bits, expect="fail", reason=reason))
    OUT.write_text(json.dumps(vectors, indent=1) + "\n")
    print(f"{len(vectors)} vectors -> {OUT}")
    npos = sum(v["expect"] == "ok" for v in vectors)
    print(f"{npos} positive, {len(vectors) - npos} negative")


if __name__ == "__main__":
    sys.exit(main())
