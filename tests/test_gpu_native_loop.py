"""GPU parity of bk_palc_run (include/bk200.h, SURVEY.md 8(b) optional entry): the all-native PALC loop against the
plugin-surface loop (palc.continuation: one C-ABI call per kernel group, the way the Julia adapter drives the library) on the
same context -- same kernels in the same order, so the branches must agree to the last bit -- and, through that loop's own
parity tests (test_gpu_palc.py), against the oracle."""
import numpy as np
import pytest

import __graft_entry__ as g
from oracle import problems

pytestmark = pytest.mark.gpu
LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)
KEYS = ("param", "x", "itnewton", "itlinear", "ds", "step")


@pytest.fixture(scope="module")
def bk():
    return g.load_package()


def _same_rows(rows, ref, exact=True):
    assert len(rows) == len(ref), (len(rows), len(ref))
    for r, o in zip(rows, ref):
        for k in KEYS:
            if exact:
                assert r[k] == o[k], (k, r, o)
            else:
                assert abs(r[k] - o[k]) <= 1e-12 * max(1.0, abs(o[k])), (k, r, o)


def _chan(bk, n, matrixfree=False):
    """the two Chan wirings the other GPU tests already run: config 1 (examples/chan.jl:108-118: GMRES restart 20, maxiter 10,
    reltol 1e-5, Pl = lu(P)) and the N + 1 system of MatrixFreeBLS on a full Krylov space"""
    m = n + 9 if matrixfree else 20
    ctx = bk.Context(bk.BK_CHAN, (n,), (1.0,), krylov_m=m, params=(3.3, 0.01))
    ctx.precond_setup(bk.BK_PC_CHAN_TRIDIAG)
    if matrixfree:
        ls = bk.GMRESB200(reltol=1e-10, restart=m, maxiter=2 * m, Pl=True, orth="cgs2")
    else:
        ls = bk.GMRESB200(reltol=1e-5, N=n, restart=20, maxiter=10, Pl=True)
    return ctx, ls


@pytest.mark.parametrize("tangent,bls_kind", [("secant", "bordering"), ("bordered", "bordering"), ("secant", "matrixfree")])
def test_native_loop_chan_bit_identical_to_the_plugin_loop(bk, tangent, bls_kind):
    """config 1 (examples/chan.jl:97-118 at N = 1e3): both tangents, both bordered solvers, norminf"""
    P = bk.palc
    n = 101 if bls_kind == "matrixfree" else 1000
    ctx, ls = _chan(bk, n, bls_kind == "matrixfree")
    bls = bk.BorderingBLSB200(ls) if bls_kind == "bordering" else bk.MatrixFreeBLSB200(ls)
    cp = P.ContinuationPar(dsmin=0.01, dsmax=0.5, ds=0.01, p_max=4.2, p_min=-1.0, max_steps=25,
                           newton_options=P.NewtonPar(tol=1e-9, max_iterations=10, linsolver=ls))
    u0 = problems.chan_sol0(n)
    alg = P.PALC(tangent=tangent, bls=bls)
    ref, st = P.continuation(P.BifurcationProblemB200(ctx, ctx.to_device(u0), (3.3, 0.01), lens=0), alg, cp, normC=P.norminf)
    rows, info = P.continuation_native(P.BifurcationProblemB200(ctx, ctx.to_device(u0), (3.3, 0.01), lens=0), alg, cp, normC=P.norminf)
    assert len(ref) > 10
    _same_rows(rows, ref)
    assert info["steps"] == st.step and info["nfail"] == st.nfail and info["work_newton"] == st.work_newton and info["work_linear"] == st.work_linear
    assert np.array_equal(info["u"].numpy(), st.z_u.numpy()) and info["p"] == st.z_p
    # host start vector (uploaded once inside the call): same branch
    rows_h, _ = P.continuation_native(P.BifurcationProblemB200(ctx, u0, (3.3, 0.01), lens=0), alg, cp, normC=P.norminf)
    _same_rows(rows_h, ref)


def test_native_loop_sh2d_front_branch(bk):
    """examples/SH2d-fronts.jl:44-86 at 128 x 64: the localized-front branch through MatrixFreeBLS + GMRES(Pr = DCT) on the fused
    JVP+Arnoldi kernels (the benchmark's wiring), 12 steps, against the plugin loop; two-point restart from a callback's seeds"""
    P = bk.palc
    dims = (128, 64)
    ctx = bk.Context(bk.BK_SH2D, dims, (LX, LY), krylov_m=100, params=(-0.1, 1.3))
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    ls = bk.GMRESB200(reltol=1e-5, restart=100, maxiter=100, N=ctx.N, Pr=True)
    mk = lambda u: P.BifurcationProblemB200(ctx, u, (-0.1, 1.3), lens=0)
    hexa = P.newton(mk(ctx.to_device(problems.sh2d_sol0(*dims, LX, LY))), ctx.to_device(problems.sh2d_sol0(*dims, LX, LY)), -0.1,
                    P.NewtonPar(tol=1e-8, max_iterations=20, linsolver=ls), P.norminf)
    assert hexa.converged
    front = problems.sh2d_front_guess(hexa.u.numpy(), *dims, LX, LY)
    fr = P.newton(mk(ctx.to_device(front)), ctx.to_device(front), -0.1, P.NewtonPar(tol=1e-9, max_iterations=30, linsolver=ls), P.norminf)
    assert fr.converged
    cp = P.ContinuationPar(dsmin=1e-4, dsmax=5e-3, ds=-1e-3, p_min=-1.0, p_max=0.0, max_steps=12,
                           newton_options=P.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ls))
    alg = P.PALC(bls=bk.MatrixFreeBLSB200(ls))
    ref, st = P.continuation(mk(fr.u), alg, cp, normC=P.norminf)
    seeds = {}

    def cb(step, row, z_u, z_p):
        if step in (5, 6):
            v = ctx.zeros()
            ctx.lib.bk_vec_copy(ctx.handle, v.dptr, z_u, ctx.N)
            seeds[step] = (v, z_p)
        return True

    rows, info = P.continuation_native(mk(fr.u), alg, cp, normC=P.norminf, callback=cb)
    assert len(ref) == 13 and abs(ref[-1]["param"] + 0.1) > 1e-3
    _same_rows(rows, ref)
    assert np.array_equal(info["u"].numpy(), st.z_u.numpy())
    # the corrected points solve F(u, lambda) = 0 with the oracle's sparse-matrix residual
    sh = problems.SwiftHohenberg(dims, (LX, LY), l=info["p"], nu=1.3)
    assert np.max(np.abs(sh.F(info["u"].numpy(), info["p"]))) < 1e-8
    # iterate_from_two_points (src/Continuation.jl:408-456) inside the library: seeded at steps 5, 6 the run goes on from there
    (u5, p5), (u6, p6) = seeds[5], seeds[6]
    cp2 = P.ContinuationPar(dsmin=1e-4, dsmax=5e-3, ds=ref[5]["ds"], p_min=-1.0, p_max=0.0, max_steps=4,
                            newton_options=P.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ls))
    prob2 = mk(u5)
    prob2.p0 = p5
    prob2.params[0] = p5
    ref2b, _ = P.continuation(prob2, alg, cp2, normC=P.norminf, u1=u6, p1=p6)
    rows2, _ = P.continuation_native(prob2, alg, cp2, normC=P.norminf, u1=u6, p1=p6)
    _same_rows(rows2, ref2b)
    assert rows2[0]["param"] == p5 and len(rows2) == 5


def test_native_loop_reports_startup_failure_and_stops_on_request(bk):
    P = bk.palc
    ctx, _ = _chan(bk, 200)
    ls = bk.GMRESB200(reltol=1e-8, N=200, restart=20, maxiter=20, Pl=True)
    cp = P.ContinuationPar(dsmin=0.01, dsmax=0.5, ds=0.01, p_max=4.2, p_min=-1.0, max_steps=25,
                           newton_options=P.NewtonPar(tol=1e-9, max_iterations=10, linsolver=ls))
    alg = P.PALC(bls=bk.BorderingBLSB200(ls))
    bad = np.full(200, 1e3)  # far from any solution: the start-up Newton fails -> the reference throws (src/Continuation.jl:375-379)
    with pytest.raises(RuntimeError, match="Newton failed"):
        P.continuation_native(P.BifurcationProblemB200(ctx, bad, (3.3, 0.01), lens=0), alg,
                              P.ContinuationPar(p_max=4.2, newton_options=P.NewtonPar(tol=1e-12, max_iterations=2, linsolver=ls)), normC=P.norminf)
    u0 = problems.chan_sol0(200)
    rows, info = P.continuation_native(P.BifurcationProblemB200(ctx, u0, (3.3, 0.01), lens=0), alg, cp, normC=P.norminf,
                                       callback=lambda step, row, z, p: step < 4)
    assert info["stopped"] == 2 and len(rows) == 5
    rows, info = P.continuation_native(P.BifurcationProblemB200(ctx, u0, (3.3, 0.01), lens=0), alg, cp, normC=P.norminf, max_rows=3)
    assert info["stopped"] == 3 and len(rows) == 3
