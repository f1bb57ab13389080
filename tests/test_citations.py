"""Every `file:line` citation of the reference in the headers, the docs, the oracle and the host mirror must point
into an existing file of the reference tree (run in the build container only; skipped where /root/reference is
absent, e.g. on the GPU box)."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"(?<![\w/.])((?:[\w.-]+/)*[\w.-]+\.(?:py|pyf|f|ipynb|txt)):(\d+)(?:-(\d+))?")

SOURCES = (["DESIGN.md", "INTEGRATION.md", "README.md"] + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
           + sorted(glob.glob(os.path.join(ROOT, "oracle", "*.py")))
           + sorted(glob.glob(os.path.join(ROOT, "richmol_b200", "*.py")))
           + sorted(glob.glob(os.path.join(ROOT, "richmol_b200", "csrc", "*.cu*")))
           + sorted(glob.glob(os.path.join(ROOT, "richmol_b200", "csrc", "*.h")))
           + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
OWN = {os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "**", "*.py"), recursive=True)}


def _resolve(name):
    for cand in (name, os.path.join("richmol", name), os.path.join("richmol", "rot", name),
                 os.path.join("examples", name), os.path.join("tests", name),
                 os.path.join("docs", "source", "notebooks", name), os.path.join("expokit", name)):
        p = os.path.join(REF, cand)
        if os.path.isfile(p):
            return p
    return None


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_citations_resolve():
    bad, checked = [], 0
    for src in SOURCES:
        path = src if os.path.isabs(src) else os.path.join(ROOT, src)
        text = open(path, encoding="utf-8").read()
        for m in CITE.finditer(text):
            name, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            ref = _resolve(name)
            if ref is None:
                if os.path.basename(name) in OWN or name.startswith(("tests/", "tools/", "oracle/", "richmol_b200/")):
                    continue                      # a citation of this repository's own files
                bad.append(f"{os.path.relpath(path, ROOT)}: {m.group(0)} (no such file in the reference)")
                continue
            with open(ref, encoding="utf-8", errors="replace") as f:
                nlines = sum(1 for _ in f)
            checked += 1
            if not (1 <= lo <= hi <= nlines):
                bad.append(f"{os.path.relpath(path, ROOT)}: {m.group(0)} ({os.path.relpath(ref, REF)} has {nlines} lines)")
    assert checked > 50, checked
    assert not bad, "\n".join(bad)
