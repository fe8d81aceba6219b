#!/usr/bin/env python3
"""Extracts the sequence / alignment literals of the reference's own unit tests
and doctests into tests/golden/reference_vectors.json.

Runs ONLY in the build container (reads /root/reference, which does not exist on
the GPU box); the JSON it writes is committed.  It copies test DATA (DNA strings,
expected alignment strings) -- not code.  The small scalar expectations (k,
thresholds, positions, variants) are transcribed by hand in
tests/test_oracle_golden.py with file:line citations.

Blocks are keyed "<file>::<test fn name>" for #[test] functions and
"<file>::doc@<line>" for doctest code fences (line = 1-based line of ```rust).
Each block maps variable name -> ASCII string for every
  let [mut] NAME[: Vec<u8>|Vec<char>] = vec![b'A',...] | vec!['M',...] | b"...".to_vec() | b"...";
"""
import json
import os
import re
import sys

REF = "/root/reference/src"
FILES = ["lib.rs", "index.rs", "gap_filling.rs", "variant_calling.rs", "translate.rs", "format.rs", "derandomize.rs"]

LET = re.compile(r"let\s+(?:mut\s+)?(\w+)\s*(?::\s*[\w<>:\s]+?)?\s*=\s*(.+?);\s*$", re.S)
VEC_BYTES = re.compile(r"^vec!\[\s*((?:b'.'\s*,?\s*)+)\]$", re.S)
VEC_CHARS = re.compile(r"^vec!\[\s*((?:'.'\s*,?\s*)+)\]$", re.S)
BSTR = re.compile(r'^b"([^"]*)"(?:\.to_vec\(\))?$')


def parse_value(v):
    v = v.strip()
    m = VEC_BYTES.match(v)
    if m:
        return "".join(re.findall(r"b'(.)'", m.group(1)))
    m = VEC_CHARS.match(v)
    if m:
        return "".join(re.findall(r"'(.)'", m.group(1)))
    m = BSTR.match(v)
    if m:
        return m.group(1)
    return None


def statements(lines):
    """Yields (possibly multi-line) `let ...;` statements."""
    cur = None
    for ln in lines:
        s = ln.strip()
        if s.startswith("#"):
            s = s[1:].strip()  # hidden doctest lines
        if cur is None:
            if s.startswith("let "):
                cur = s
            else:
                continue
        else:
            cur += " " + s
        if cur.endswith(";"):
            yield cur
            cur = None


def blocks(path):
    src = open(path).read().split("\n")
    out = {}
    i = 0
    while i < len(src):
        ln = src[i]
        if ln.strip().startswith("///") and "```rust" in ln:
            start = i + 1
            j = i + 1
            body = []
            while j < len(src) and "```" not in src[j]:
                body.append(re.sub(r"^\s*///\s?", "", src[j]))
                j += 1
            out["doc@%d" % start] = body
            i = j + 1
            continue
        if ln.strip() == "#[test]":
            m = re.search(r"fn\s+(\w+)", src[i + 1])
            name = m.group(1)
            j = i + 2
            body = []
            while j < len(src) and src[j].strip() != "#[test]":
                body.append(src[j])
                j += 1
            out[name] = body
            i = j
            continue
        i += 1
    return out


def main():
    result = {}
    for f in FILES:
        for name, body in blocks(os.path.join(REF, f)).items():
            vars_ = {}
            for st in statements(body):
                m = LET.match(st)
                if not m:
                    continue
                val = parse_value(m.group(2))
                if val is not None and len(val) > 0:
                    vars_[m.group(1)] = val
            if vars_:
                result["%s::%s" % (f, name)] = vars_
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as fh:
        json.dump(result, fh, indent=1, sort_keys=True)
    print("wrote", out, "with", len(result), "blocks")


if __name__ == "__main__":
    sys.exit(main())
