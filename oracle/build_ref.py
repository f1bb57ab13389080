"""TEST INFRASTRUCTURE ONLY -- recipe that builds `oracle/_ref/` from the UNMODIFIED reference.

The reference (CFEL-CMI/richmol) is pure Python on the TDSE path, so "building" it means
byte-compiling the handful of modules the path imports, from the sources where they lie under
/root/reference, into `oracle/_ref/richmol/*.pyc` (sourceless layout: Python imports `X.pyc` next
to a missing `X.py`) and packing those into `oracle/_ref/richmol_ref.zip` (zipimport reads sourceless
byte code from an archive; the snapshot that travels to the GPU box drops loose `*.pyc` files).  Only
compiled outputs are written -- no reference source is copied into this repository -- and
`oracle/_ref/` is git-ignored (not gpurun-ignored), so the byte code travels to the GPU box where
/root/reference does not exist.  There `oracle/refshim.py` imports it with
the stub modules of SURVEY.md Appendix B, and `bench.py --impl reference` / `cpu_baseline` time the
reference's own `CarTens.field` + `TDSE.update` (cpu_baseline.kind = "reference").

    python oracle/build_ref.py            # no-op when /root/reference is absent

Modules (all of what `richmol.tdse` and `richmol.field` import from the package):
richmol/__init__.py, field.py, tdse.py, convert_units.py, pyexpokit.py, json_ext.py, trove.py.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RICHMOL_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("__init__", "field", "tdse", "convert_units", "pyexpokit", "json_ext", "trove")


def build(quiet=True):
    src_dir = os.path.join(REF_ROOT, "richmol")
    if not os.path.isdir(src_dir):
        return False
    dst_dir = os.path.join(OUT, "richmol")
    os.makedirs(dst_dir, exist_ok=True)
    for m in MODULES:
        src = os.path.join(src_dir, m + ".py")
        dst = os.path.join(dst_dir, m + ".pyc")
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        # unchecked-hash pycs: valid without the source file and independent of its mtime
        py_compile.compile(src, cfile=dst, dfile=f"richmol/{m}.py", doraise=True, quiet=2 if quiet else 0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    import zipfile
    with zipfile.ZipFile(os.path.join(OUT, "richmol_ref.zip"), "w", zipfile.ZIP_STORED) as z:
        for m in MODULES:
            z.write(os.path.join(dst_dir, m + ".pyc"), f"richmol/{m}.pyc")
    with open(os.path.join(OUT, "VERSION"), "w") as f:
        f.write(f"byte code of {', '.join(m + '.py' for m in MODULES)} from {src_dir}, "
                f"python {sys.version.split()[0]}\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref built" if build(quiet=False) else f"reference tree not found at {REF_ROOT}: nothing built")
