"""CPU: the C-ABI library loads and exports every symbol include/dkd_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dkd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dkd_[a-z0-9_]+)\s*\(", text)))


def test_build_and_load(dkd):
    import __graft_entry__ as g
    g.build()
    from dkd_b200 import _lib
    lib = _lib.load()
    assert lib.dkd_version() >= 100


def test_every_header_symbol_is_exported_and_bound(dkd):
    from dkd_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(syms)


def test_error_strings(dkd):
    from dkd_b200 import _lib
    lib = _lib.load()
    assert b"argument" in lib.dkd_error_string(-1)
    assert b"shape" in lib.dkd_error_string(-2)


def test_no_cpu_fallback(dkd):
    """The tensor wrappers refuse CPU tensors instead of falling back to PyTorch."""
    import torch
    from dkd_b200 import ops, _lib
    with pytest.raises(_lib.DkdError):
        ops.normalize_rows(torch.zeros(4, 64))
    with pytest.raises(_lib.DkdError):
        ops.topk(torch.zeros(4, 64), 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dl-dkd_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("CPU oracle", ""), f"{f} references oracle/"
