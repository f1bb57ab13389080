"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (richmol) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It is used by
`tests/golden/make_golden.py` to generate golden vectors and by the CPU tests that pin the
numpy restatement in `oracle/` against the real reference.  Nothing in `richmol_b200/`
may import this module.

Recipe follows SURVEY.md Appendix B: the reference's import-time dependencies that are absent
here (h5py, jax, the f2py `expokit` module) are replaced by empty stubs *outside* the
reference tree, and `numpy.float_` (removed in numpy 2) is aliased.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("RICHMOL_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "richmol"))


def load():
    """Returns the imported reference modules (field, tdse, trove, convert_units)."""
    if not available():
        raise ImportError(f"reference tree not found at {REF_ROOT}")
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
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import richmol.field as field
        import richmol.tdse as tdse
        import richmol.trove as trove
        import richmol.convert_units as convert_units
    return types.SimpleNamespace(field=field, tdse=tdse, trove=trove, convert_units=convert_units)
