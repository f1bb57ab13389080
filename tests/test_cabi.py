"""The C-ABI library loads on a box without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from richmol_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "richmol_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rmb_[a-z0-9_]+)\s*\(", txt)))


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_exports_every_header_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/richmol_b200.h but not exported"
    # and the ctypes binding covers exactly the header
    assert sorted(_lib.SYMBOLS) == names


def test_abi_version_and_error_string():
    lib = _lib.lib()
    assert lib.rmb_abi_version() == 1
    assert isinstance(lib.rmb_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_gpu():
    lib = _lib.lib()
    assert lib.rmb_operator_create(None, None) == _lib.RMB_ERR_INVALID
    assert lib.rmb_matvec(None, None, None, 0, 0, None) == _lib.RMB_ERR_INVALID
    assert lib.rmb_populations(None, 0, 0, 0, None, None) == _lib.RMB_ERR_INVALID


def test_product_path_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from richmol_b200 import synth
    t = synth.ocs(2)["pol"]
    t.field([0, 0, 1e8])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        t._device()
