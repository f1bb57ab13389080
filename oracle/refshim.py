"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (richmol) from /root/reference.

In the build container it imports the sources under /root/reference; on the GPU box (no
/root/reference) it imports the byte code `oracle/build_ref.py` compiled from them into
`oracle/_ref/` (git-ignored, travels with the snapshot).  It is used by
`tests/golden/make_golden.py` to generate golden vectors, by the CPU tests that pin the
numpy restatement in `oracle/` against the real reference, and by the CPU legs of `bench.py`
(`--impl reference`, `cpu_baseline`).  Nothing in `richmol_b200/` may import this module.

Recipe follows SURVEY.md Appendix B: the reference's import-time dependencies that are absent
here (h5py, jax, the f2py `expokit` module) are replaced by empty stubs *outside* the
reference tree, and `numpy.float_` (removed in numpy 2) is aliased.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("RICHMOL_REFERENCE", "/root/reference")
BUILT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "richmol_ref.zip")


def available():
    """The reference SOURCE tree is present (build container): fixtures under tests/benchmarks exist."""
    return os.path.isdir(os.path.join(REF_ROOT, "richmol"))


def built():
    """The byte-compiled reference modules of `oracle/build_ref.py` are present."""
    return os.path.isfile(BUILT_ROOT)


def root():
    return REF_ROOT if available() else (BUILT_ROOT if built() else None)


def load():
    """Returns the imported reference modules (field, tdse, trove, convert_units)."""
    ref_root = root()
    if ref_root is None:
        raise ImportError(f"reference not found at {REF_ROOT} nor byte-compiled under {BUILT_ROOT}")
    import numpy as np
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.is_hdf5 = lambda f: False
        h5.File = None
        sys.modules["h5py"] = h5
    if "jax" not in sys.modules:
        jax = types.ModuleType("jax")
        lib = types.ModuleType("jax.lib")
        xb = types.ModuleType("jax.lib.xla_bridge")
        xb.get_backend = lambda: types.SimpleNamespace(platform="cpu")
        jax.lib = lib
        lib.xla_bridge = xb
        sys.modules["jax"] = jax
        sys.modules["jax.lib"] = lib
        sys.modules["jax.lib.xla_bridge"] = xb
    if "expokit" not in sys.modules:
        sys.modules["expokit"] = types.ModuleType("expokit")
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import richmol.field as field
        import richmol.tdse as tdse
        import richmol.trove as trove
        import richmol.convert_units as convert_units
    return types.SimpleNamespace(field=field, tdse=tdse, trove=trove, convert_units=convert_units)


_DATA_ATTRS = ("Jlist1", "Jlist2", "symlist1", "symlist2", "dim1", "dim2", "dim_k1", "dim_k2", "dim_m1",
               "dim_m2", "quanta_k1", "quanta_k2", "quanta_m1", "quanta_m2", "rank", "cart", "os", "kmat", "mmat")


def to_reference(r, tens):
    """An instance of the UNMODIFIED `richmol.field.CarTens` carrying the data model (nested dicts of
    scipy CSR blocks, richmol/field.py:58-170) of `tens` -- how synthetic benchmark tensors reach the
    reference's own `field` / `vec` / `TDSE.update`."""
    import copy
    out = r.field.CarTens()
    for a in _DATA_ATTRS:
        if hasattr(tens, a):
            setattr(out, a, copy.deepcopy(getattr(tens, a)))
    return out
