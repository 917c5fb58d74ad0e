"""CPU tests of the native PALC loop (bifurcationkit.jl_b200/csrc/bk_palc_loop.hpp = the body of bk_palc_run): the template
is instantiated with a host backend whose problem / solvers are Python callbacks (tests/native_loop/palc_loop_host.cpp, test
harness only) and compared row by row with the Python host loop (palc.py) and with the reference's known answers
(test/continuation/test-cont-non-vector.jl:22-45: param[end] == -1; src/continuation/Contbase.jl:77-102)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as g
from oracle import krylov, bls as obls, problems
from tests.test_host_logic_cpu import NumpyProblem, BlsAdapter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDIR = os.path.join(ROOT, "tests", "native_loop")
dp = C.POINTER(C.c_double)


class HostOpts(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ds", "dsmin", "dsmax", "a", "p_min", "p_max", "theta", "eta", "newton_tol", "fd_eps")] + \
               [(k, C.c_int32) for k in ("max_steps", "newton_maxit", "tangent", "normc")]


RES_F = C.CFUNCTYPE(None, dp, C.c_double, dp)
JAC_F = C.CFUNCTYPE(None, dp, C.c_double)
LIN_F = C.CFUNCTYPE(C.c_int32, dp, dp, C.POINTER(C.c_int32))
BLS_F = C.CFUNCTYPE(C.c_int32, dp, dp, C.c_double, dp, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, C.POINTER(C.c_int32))
STEP_F = C.CFUNCTYPE(C.c_int32, C.c_int32, dp, dp, C.c_double)


class HostCallbacks(C.Structure):
    _fields_ = [("residual", RES_F), ("jacobian", JAC_F), ("linsolve", LIN_F), ("bls", BLS_F), ("on_step", STEP_F)]


class HostResult(C.Structure):
    _fields_ = [("nrows", C.c_int32), ("steps", C.c_int32), ("nfail", C.c_int32), ("stopped", C.c_int32),
                ("work_newton", C.c_int64), ("work_linear", C.c_int64), ("p_final", C.c_double), ("ds_final", C.c_double)]


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(HDIR, "libpalc_loop_host.so")
    src = os.path.join(HDIR, "palc_loop_host.cpp")
    hdr = os.path.join(ROOT, "bifurcationkit.jl_b200", "csrc", "bk_palc_loop.hpp")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.palc_loop_host_run.restype = C.c_int32
    lib.palc_loop_host_run.argtypes = [C.POINTER(HostOpts), C.c_int64, C.POINTER(HostCallbacks), dp, C.c_double, dp, C.c_double,
                                       dp, C.c_int32, dp, C.POINTER(HostResult)]
    lib.palc_loop_host_step_size.restype = C.c_double
    lib.palc_loop_host_step_size.argtypes = [C.POINTER(HostOpts), C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    return lib


def native_run(lib, F, J, linsolver, bls, u0, p0, cp, theta=0.5, tangent="secant", normc=0, u1=None, p1=0.0, max_rows=None,
               on_step=None, eta=150.0):
    """drives the C++ loop with NumPy callbacks; J(x, p) -> dense matrix; linsolver / bls = the oracle's host solvers"""
    n = len(u0)
    view = lambda p: np.ctypeslib.as_array(p, shape=(n,))
    state = {}

    def residual(x, p, out):
        view(out)[...] = F(view(x).copy(), p)

    def jacobian(x, p):
        state["J"] = J(view(x).copy(), p)

    def linsolve(rhs, out, it):
        u, ok, k = linsolver(state["J"], view(rhs).copy())
        view(out)[...] = u
        it[0] = int(np.sum(k))
        return int(bool(ok))

    def blsf(dR, dzu, dzp, R, nn, xiu, xip, dotscale, dX, dl, it):
        u, up, ok, k = bls(state["J"], view(dR).copy(), view(dzu).copy(), dzp, view(R).copy(), nn, xiu, xip, shift=None, dotscale=dotscale)
        view(dX)[...] = u
        dl[0] = up
        it[0] = int(np.sum(k))
        return int(bool(ok))

    def step(k, row, z_u, z_p):
        return 1 if on_step is None else int(bool(on_step(k, np.ctypeslib.as_array(row, shape=(6,)).copy(), view(z_u).copy(), z_p)))

    cbs = HostCallbacks(RES_F(residual), JAC_F(jacobian), LIN_F(linsolve), BLS_F(blsf), STEP_F(step))
    no = cp.newton_options
    ho = HostOpts(cp.ds, cp.dsmin, cp.dsmax, cp.a, cp.p_min, cp.p_max, theta, eta, no.tol, 0.0, cp.max_steps, no.max_iterations,
                  0 if tangent == "secant" else 1, normc)
    max_rows = max_rows or cp.max_steps + 8
    rows = np.zeros((max_rows, 6))
    uf = np.zeros(n)
    res = HostResult()
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    u1p = None if u1 is None else np.ascontiguousarray(u1, dtype=np.float64).ctypes.data_as(dp)
    st = lib.palc_loop_host_run(C.byref(ho), n, C.byref(cbs), u0.ctypes.data_as(dp), p0, u1p, p1, rows.ctypes.data_as(dp), max_rows,
                                uf.ctypes.data_as(dp), C.byref(res))
    return st, rows[: res.nrows], uf, res


def _fold():
    return (lambda x, r: r + x - x**3), (lambda x, r: np.diag(1 - 3 * x**2))


@pytest.mark.parametrize("tangent", ["secant", "bordered"])
def test_native_loop_matches_host_loop_on_the_fold_known_answer(harness, tangent):
    """test-cont-non-vector.jl:22-45: r + x - x^3 from x = 0.8, r = 1, ds = -0.02 => last param == -1; same rows as palc.py"""
    P = g.load_package().palc
    F, J = _fold()
    ls = krylov.DefaultLS()
    bls = BlsAdapter(obls.MatrixBLS())
    cp = P.ContinuationPar(dsmin=0.001, dsmax=0.07, ds=-0.02, p_max=4.1, p_min=-1.0, max_steps=150,
                           newton_options=P.NewtonPar(tol=1e-8, linsolver=ls))
    rows_py, st_py = P.continuation(NumpyProblem(F, J, np.array([0.8]), 1.0), P.PALC(tangent=tangent, bls=bls), cp)
    st, rows, uf, res = native_run(harness, F, J, ls, bls, np.array([0.8]), 1.0, cp, tangent=tangent)
    assert st == 0 and rows[-1, 0] == -1.0
    assert len(rows) == len(rows_py) and res.steps == st_py.step and res.nfail == st_py.nfail
    for r, o in zip(rows, rows_py):
        # same expressions in the same order; the host backends differ by fused multiply-adds (BLAS daxpy vs a plain loop), the
        # device backends not at all (tests/test_gpu_native_loop.py asserts bit-identical rows there)
        assert abs(r[0] - o["param"]) < 1e-13 and abs(r[1] - o["x"]) < 1e-13 and abs(r[4] - o["ds"]) < 1e-15
        assert r[2] == o["itnewton"] and r[3] == o["itlinear"] and r[5] == o["step"]
    assert res.work_newton == st_py.work_newton and res.work_linear == st_py.work_linear
    assert abs(uf[0] - st_py.z_u[0]) < 1e-13 and abs(res.p_final - st_py.z_p) < 1e-13


@pytest.mark.parametrize("normc", [0, 1])
def test_native_loop_matches_host_loop_on_chan(harness, normc):
    """Chan / Bratu branch up to p_max (examples/chan.jl:5-19, 85-95, n = 31, dense Jacobian): rows agree to rounding"""
    P = g.load_package().palc
    n, beta = 31, 0.01
    F = lambda x, a: problems.chan_F(x, a, beta)
    E = np.eye(n)
    J = lambda x, a: np.column_stack([problems.chan_dF(x, E[:, k], a, beta) for k in range(n)])
    ls = krylov.DefaultLS()
    bls = BlsAdapter(obls.MatrixBLS())
    cp = P.ContinuationPar(dsmin=0.01, dsmax=0.2, ds=0.01, p_max=4.2, p_min=-1.0, max_steps=90,
                           newton_options=P.NewtonPar(tol=1e-10, max_iterations=10, linsolver=ls))
    normC = P.norm2 if normc == 0 else P.norminf
    rows_py, st_py = P.continuation(NumpyProblem(F, J, problems.chan_sol0(n), 3.3), P.PALC(bls=bls), cp, normC=normC)
    st, rows, uf, res = native_run(harness, F, J, ls, bls, problems.chan_sol0(n), 3.3, cp, normc=normc)
    assert st == 0 and len(rows) == len(rows_py) > 40
    assert rows[-1, 0] == 4.2 and rows_py[-1]["param"] == 4.2  # the last point sits ON p_max: Natural corrector at the bound (Palc.jl:157-160)
    for r, o in zip(rows, rows_py):
        assert abs(r[0] - o["param"]) < 1e-12 and abs(r[1] - o["x"]) < 1e-12 and r[2] == o["itnewton"] and abs(r[4] - o["ds"]) < 1e-14
    assert np.linalg.norm(uf - st_py.z_u) < 1e-11 and res.work_newton == st_py.work_newton


def test_native_loop_rejected_steps_and_dsmin_stop(harness):
    """a corrector that cannot converge in 3 iterations: steps are rejected, ds is halved down to dsmin, the run stops
    (src/continuation/Contbase.jl:80-86) -- same sequence as the host loop"""
    P = g.load_package().palc
    F, J = _fold()
    ls = krylov.DefaultLS()
    bls = BlsAdapter(obls.MatrixBLS())
    cp = P.ContinuationPar(dsmin=0.01, dsmax=0.5, ds=-0.4, p_max=4.1, p_min=-1.0, max_steps=40,
                           newton_options=P.NewtonPar(tol=1e-10, max_iterations=3, linsolver=ls))
    x0 = np.array([1.3247179572447460])  # root of 1 + x - x^3
    rows_py, st_py = P.continuation(NumpyProblem(F, J, x0, 1.0), P.PALC(bls=bls), cp)
    st, rows, uf, res = native_run(harness, F, J, ls, bls, x0, 1.0, cp)
    assert st == 0 and st_py.nfail > 0
    assert res.nfail == st_py.nfail and res.steps == st_py.step and len(rows) == len(rows_py)
    assert res.stopped == (1 if st_py.stop else 0)
    for r, o in zip(rows, rows_py):
        assert abs(r[0] - o["param"]) < 1e-13 and abs(r[4] - o["ds"]) < 1e-15 and r[2] == o["itnewton"]
    assert abs(res.ds_final - st_py.ds) < 1e-15 and res.work_newton == st_py.work_newton


def test_native_loop_step_size_control(harness):
    """src/continuation/Contbase.jl:77-102, every (converged, itnewton) case against the host loop's function"""
    P = g.load_package().palc
    for nmax in (6, 10, 15, 25, 41):
        cp = P.ContinuationPar(dsmin=1e-3, dsmax=0.1, a=0.5, newton_options=P.NewtonPar(max_iterations=nmax))
        ho = HostOpts(0.0, cp.dsmin, cp.dsmax, cp.a, 0, 0, 0.5, 150.0, 0, 0, 0, nmax, 0, 0)
        for ds in (0.01, -0.01, 1e-3, -1e-3, 0.09, 5e-4):
            for conv in (True, False):
                for it in range(nmax + 1):
                    stop = C.c_int32()
                    got = harness.palc_loop_host_step_size(C.byref(ho), ds, int(conv), it, C.byref(stop))
                    want, wstop = P.step_size_control(ds, conv, it, cp)
                    assert got == want and bool(stop.value) == wstop, (nmax, ds, conv, it)


def test_native_loop_startup_failure_callback_stop_and_full_buffer(harness):
    P = g.load_package().palc
    ls = krylov.DefaultLS()
    bls = BlsAdapter(obls.MatrixBLS())
    # no root: the reference throws "Newton failed to converge for the initial guess" (src/Continuation.jl:375-379) -> status
    Fn, Jn = (lambda x, p: np.exp(x) + 0 * p), (lambda x, p: np.diag(np.exp(x)))
    cp = P.ContinuationPar(newton_options=P.NewtonPar(tol=1e-10, max_iterations=5, linsolver=ls))
    st, rows, _, res = native_run(harness, Fn, Jn, ls, bls, np.array([1.0]), 0.0, cp)
    assert st == -3 and res.nrows == 0
    F, J = _fold()
    cp = P.ContinuationPar(dsmin=0.001, dsmax=0.07, ds=-0.02, p_max=4.1, p_min=-1.0, max_steps=150,
                           newton_options=P.NewtonPar(tol=1e-8, linsolver=ls))
    seen = []
    st, rows, _, res = native_run(harness, F, J, ls, bls, np.array([0.8]), 1.0, cp,
                                  on_step=lambda k, row, z, p: seen.append((k, row[0], z[0], p)) or k < 7)
    assert st == 0 and res.stopped == 2 and len(rows) == 8 and [s[0] for s in seen] == list(range(8))
    assert all(abs(F(np.array([z]), p)[0]) < 1e-8 and p == r0 for _, r0, z, p in seen)  # the callback sees converged points
    st, rows, _, res = native_run(harness, F, J, ls, bls, np.array([0.8]), 1.0, cp, max_rows=5)
    assert st == 0 and res.stopped == 3 and res.nrows == 5
    # p0 outside [p_min, p_max]
    st, _, _, _ = native_run(harness, F, J, ls, bls, np.array([0.8]), 5.0, cp)
    assert st == -3


def test_native_loop_two_point_start(harness):
    """iterate_from_two_points (src/Continuation.jl:408-456): seeded from two consecutive points of a branch, the native
    loop continues it exactly as the host loop does"""
    bk = g.load_package()
    P = bk.palc
    F, J = _fold()
    ls = krylov.DefaultLS()
    bls = BlsAdapter(obls.MatrixBLS())
    mk = lambda ms: P.ContinuationPar(dsmin=0.001, dsmax=0.07, ds=-0.02, p_max=4.1, p_min=-1.0, max_steps=ms,
                                      newton_options=P.NewtonPar(tol=1e-10, linsolver=ls))
    grab = bk.segments.SeedGrabber(1, 5, lambda v: v.copy())
    P.continuation(NumpyProblem(F, J, np.array([0.8]), 1.0), P.PALC(bls=bls), mk(6), callback=grab)
    u0, p0, u1, p1 = grab.pair()
    rows_py, st_py = P.continuation(NumpyProblem(F, J, u0, p0), P.PALC(bls=bls), mk(10), u1=u1, p1=p1)
    st, rows, uf, res = native_run(harness, F, J, ls, bls, u0, p0, mk(10), u1=u1, p1=p1)
    assert st == 0 and len(rows) == len(rows_py) == 11 and rows[0, 0] == p0 == rows_py[0]['param']
    for r, o in zip(rows, rows_py):
        assert abs(r[0] - o["param"]) < 1e-13 and abs(r[1] - o["x"]) < 1e-13 and r[2] == o["itnewton"]


def test_native_entry_is_bound_everywhere():
    """bk_palc_run is declared in the header, exported by the library, bound by lib.py and by the Julia adapter"""
    bk = g.load_package()
    assert "bk_palc_run" in bk.lib.SYMBOLS
    hdr = open(os.path.join(ROOT, "include", "bk200.h")).read()
    assert "bk_palc_run(" in hdr and "bk_palc_opts" in hdr
    lib = C.CDLL(bk.lib.LIB_PATH)
    assert hasattr(lib, "bk_palc_run")
    # the ctypes mirror of bk_palc_opts / bk_palc_result has the size the C compiler gives the structs
    assert C.sizeof(bk.lib.PalcOpts) == 11 * 8 + 8 * 4 and C.sizeof(bk.lib.PalcResult) == 4 * 4 + 2 * 8 + 2 * 8
    jl = open(os.path.join(ROOT, "julia", "BK200.jl")).read()
    assert "bk_palc_run" in jl
