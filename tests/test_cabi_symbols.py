"""CPU test (no GPU, no compute calls): liblprb200.so loads and exports every symbol that
include/lprb200.h declares, and the ctypes table mirrors the header one to one."""
import ctypes
import os
import re

import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "lprb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return re.findall(r"\b(?:int|int64_t|const char\*)\s+(lpr_\w+)\s*\(", hdr)


def test_every_declared_symbol_is_exported():
    syms = header_symbols()
    assert len(syms) >= 60
    lib = ctypes.CDLL(N.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert sorted(N.SIGNATURES) == sorted(header_symbols())


def test_no_stub_left():
    assert not os.path.exists(os.path.join(ROOT, "lpr_381_group_v22_b200", "csrc", "stubs.cu")), \
        "a stub source was left in csrc/: every entry point of the header must be implemented"


def test_version_and_error_plumbing_without_gpu():
    lib = N.lib()
    assert lib.lpr_version() == 100
    assert isinstance(lib.lpr_last_error(), bytes)
    assert L.launch_count() >= 0
    if L.device_count() == 0:
        # no CPU fallback: compute entry points fail loudly with LPR_E_CUDA
        import numpy as np
        import pytest
        with pytest.raises(L.LprError) as ei:
            L.DeviceTableau.from_host(np.zeros((2, 3)))
        assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
        with pytest.raises(L.LprError):
            L.KnapsackBranchBoundSimplex(10.0, [1.0], [1.0]).Solve()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lpr_381_group_v22_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                # comments may CITE the oracle; nothing may import, include, link or dlopen it
                assert not re.search(r"^\s*(import|from)\s+oracle_lib", src, flags=re.M), f
                assert not re.search(r"#include\s*[<\"][^>\"]*lpr_oracle", src), f
                assert "liblpr_oracle" not in src and "oracle/" not in src.replace("oracle/lpr_oracle.cpp", ""), f
                # nor the C# interpreter, the P/Invoke layer or the oracle-backed double of the shim tests
                assert not re.search(r"^\s*(import|from)\s+(csharp|lprb200_double|csharp_shim_cases|net_reference)", src, flags=re.M), f
                assert "lprb200_double" not in src and "pinvoke" not in src, f
