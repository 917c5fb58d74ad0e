"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol that
include/bk200.h declares (no compute calls without a GPU)."""
import os
import re
import ctypes

import pytest

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bk200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    bk = g.load_package()
    if not os.path.exists(bk.lib.LIB_PATH):
        bk.build()
    lib = ctypes.CDLL(bk.lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bk200.h but not exported"
    assert set(bk.lib.SYMBOLS) == set(names), set(bk.lib.SYMBOLS) ^ set(names)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bk = g.load_package()
    with pytest.raises(bk.BK200Error):
        bk.Context(bk.BK_SH2D, (16, 16), (1.0, 1.0), krylov_m=4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bifurcationkit.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
