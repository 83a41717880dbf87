"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/csvplus_b200.h declares; without a GPU every compute entry fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_symbols():
    import __graft_entry__ as g
    g.build()
    from csvplus_b200 import _abi
    L = _abi.load()
    hdr = open(os.path.join(ROOT, "include", "csvplus_b200.h")).read()
    declared = set(re.findall(r"\b(cpb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.cpb_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import csvplus_b200 as cp
    with pytest.raises(cp.CsvPlusError, match="no CPU fallback"):
        cp.Context(0)
    with pytest.raises(cp.CsvPlusError):
        cp.Take(cp.FromBytes(b"a,b\n1,2\n")).ToRows()


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "csvplus_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in src.lower() or f == "api.py" and "oracle" not in src, f


def test_struct_layouts_match_header():
    from csvplus_b200 import _abi
    assert C.sizeof(_abi.Error) == 512
    assert C.sizeof(_abi.ReaderOpts) == 16
    assert C.sizeof(_abi.HeaderCol) == 24
    assert C.sizeof(_abi.KStat) == 72
    assert C.sizeof(_abi.Pred) == 32
