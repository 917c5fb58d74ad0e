"""The C++/OpenMP CPU baseline (oracle/c) is itself checked against the NumPy oracle (which is pinned to the
reference's known answers): same PALC rows on the same start point.  CPU-only test."""
import numpy as np

from oracle import problems, krylov, bls as obls, palc as opalc, precond as oprecond, cbaseline

LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)


def test_cpp_baseline_rows_match_numpy_oracle():
    cbaseline.build()
    dims, L = (128, 64), (LX, LY)
    N = dims[0] * dims[1]
    sh = problems.SwiftHohenberg(dims, L, l=-0.1, nu=1.3)
    Pinv = oprecond.dct_precond(dims, L, 1.0)
    Pb = lambda r: Pinv(r) if len(r) == N else np.concatenate([Pinv(r[:N]), r[N:]])
    ols = krylov.GMRESIterativeSolvers(reltol=1e-5, restart=100, maxiter=100, N=N, Pr=Pb)
    mk = lambda u0: opalc.Problem(F=lambda u, l: sh.F(u, l), J=lambda u, l: (lambda v: sh.dF(u, v, l)), u0=u0, p0=-0.1)
    u0 = problems.sh2d_sol0(*dims, *L)
    # Newton to the hexagons, then to the localized front (examples/SH2d-fronts.jl:57-80): C++ vs NumPy
    co = cbaseline.make_opts(max_steps=4, nthreads=4)
    hexa = opalc.newton(mk(u0), u0, -0.1, opalc.NewtonPar(tol=1e-8, max_iterations=20, linsolver=ols), opalc.norminf)
    uc, okc, itn, itl = cbaseline.newton(dims, L, -0.1, 1.3, u0, 1e-8, 20, co)
    assert hexa.converged and okc and itn == hexa.itnewton
    assert np.max(np.abs(uc - hexa.u)) < 1e-7
    front = problems.sh2d_front_guess(hexa.u, *dims, *L)
    fr = opalc.newton(mk(front), front, -0.1, opalc.NewtonPar(tol=1e-9, max_iterations=30, linsolver=ols), opalc.norminf)
    assert fr.converged
    cp = opalc.ContinuationPar(dsmin=1e-4, dsmax=5e-3, ds=-1e-3, p_min=-1.0, p_max=0.0, max_steps=4,
                               newton_options=opalc.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ols))
    orows, _ = opalc.continuation(mk(fr.u), opalc.PALC(bls=obls.MatrixFreeBLS(ols)), cp, normC=opalc.norminf)
    crows, secs, tstep, ufin, work = cbaseline.palc(dims, L, 1.3, fr.u, -0.1, co)
    assert len(crows) == len(orows) == 5
    for c, o in zip(crows, orows):
        assert abs(c["param"] - o["param"]) < 1e-9, (c, o)
        assert abs(c["x"] - o["x"]) < 1e-8 * o["x"], (c, o)
        assert c["itnewton"] == o["itnewton"] and abs(c["itlinear"] - o["itlinear"]) <= 1, (c, o)
    assert secs > 0 and np.all(np.diff(tstep[1:]) > 0)
