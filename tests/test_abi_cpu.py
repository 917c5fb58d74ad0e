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


def test_hessenberg_eig_golden_5x5():
    """Host part of the eigensolver (complex shifted QR + inverse iteration) against the reference's golden
    spectrum, test/linear_solvers/test_linear.jl:595-614 (matrix reduced to Hessenberg form first)."""
    import numpy as np
    import scipy.linalg as sla
    from tests.test_oracle_linear import J5, GOLD5
    bk = g.load_package()
    Hh, Q = sla.hessenberg(J5, calc_q=True)
    vals, vecs = bk.hessenberg_eig(Hh)
    key = lambda z: (-round(z.real, 9), z.imag)
    assert np.allclose(sorted(vals, key=key), sorted(GOLD5, key=key), atol=1e-12)
    for k in range(5):
        v = Q @ vecs[:, k]
        assert np.linalg.norm(J5 @ v - vals[k] * v) < 1e-8
    # random Hessenberg matrices, incl. symmetric tridiagonal (the SH case)
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 40):
        A = np.triu(rng.standard_normal((n, n)), -1)
        vals, vecs = bk.hessenberg_eig(A)
        ref = np.linalg.eigvals(A)
        assert np.allclose(sorted(vals, key=key), sorted(ref, key=key), atol=1e-9)
        for k in range(n):
            assert np.linalg.norm(A @ vecs[:, k] - vals[k] * vecs[:, k]) < 1e-7 * max(1, np.abs(vals).max())
    T = np.diag(rng.standard_normal(30)) + np.diag(rng.standard_normal(29), 1)
    T = T + np.triu(T, 1).T
    vals, _ = bk.hessenberg_eig(T, vectors=False)
    assert np.allclose(np.sort(vals.real), np.linalg.eigvalsh(T), atol=1e-10) and np.abs(vals.imag).max() < 1e-10


def test_header_is_plain_c_and_a_c_caller_links_and_fails_loudly(tmp_path):
    """include/bk200.h is valid C99 (the Julia ccall / cgo / JNI side sees a C ABI), and a plain C program -- tools/palc_cli, the
    native driver over bk_palc_run -- builds against libbk200.so with no C++ or torch in sight; without a GPU it must stop with the
    library's own error message, not fall back to anything."""
    import shutil
    import subprocess
    import torch
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    bk = g.load_package()
    if not os.path.exists(bk.lib.LIB_PATH):
        bk.build()
    exe = str(tmp_path / "bk_palc_cli")
    libdir = os.path.dirname(bk.lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tools", "palc_cli", "bk_palc_cli.c"), "-o", exe, "-L", libdir, "-lbk200", "-lm",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is a GPU job")
    run = subprocess.run([exe, "--grid", "64", "--steps", "2"], capture_output=True, text=True, timeout=120)
    assert run.returncode == 2 and "bk_ctx_create failed" in run.stderr and run.stdout == ""
