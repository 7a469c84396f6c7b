"""The C-ABI library loads on a CPU-only box and exports every symbol include/limap_b200.h declares
(no compute calls without a GPU), and the product path fails loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "limap_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    import __graft_entry__ as g
    g.build()
    from limap_b200 import _cabi
    L = C.CDLL(_cabi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in include/limap_b200.h but not exported"
    # the ctypes table binds exactly the header's functions
    assert set(_cabi.EXPORTED_SYMBOLS) == set(syms)
    assert b"sm_100a" in _cabi.lib().lm_version()


def test_struct_layouts_match_header():
    from limap_b200 import _cabi
    from limap_b200.config import LinkerConfig, TriConfig
    assert C.sizeof(LinkerConfig) == 8 * 8 + 6 * 4
    assert C.sizeof(TriConfig) == 6 * 8 + 12 * 4 + 2 * C.sizeof(LinkerConfig)
    assert _cabi.NODE_RECORD_DTYPE.itemsize == 96
    assert C.sizeof(_cabi.TriStats) == 8 * 8 + 2 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from limap_b200._cabi import LimapB200Error
    from limap_b200.engine import TriEngine
    with pytest.raises(LimapB200Error, match="no CPU fallback"):
        TriEngine({})


def test_product_code_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under limap_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "limap_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|#include\s*[\"<][^\n]*oracle|liblimap_oracle|orc_[a-z_]+\s*\()")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), f"{os.path.join(dp, f)} uses the oracle"
