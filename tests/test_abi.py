"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/rlb200.h declares, the ctypes table mirrors the header, and no compute path exists without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "rlb200.h")).read()
    return sorted(set(re.findall(r"RLB200_API\s+[\w\s\*]+?\b(rlb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from randlapack_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rlb200.h but not exported by librlb200.so"
    assert lib.rlb200_abi_version() == 1


def test_ctypes_table_matches_header():
    from randlapack_b200 import _capi
    assert sorted(_capi.SIGNATURES) == _header_symbols()
    _capi.load()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import randlapack_b200 as rl
    with pytest.raises(rl.Error):
        rl.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "randlapack_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no oracle", "").replace("imports oracle/", "") or f == "__init__.py", f
