"""GPU: codim-2 curves (continuation_fold / continuation_hopf of src/codim2/MinAugFold.jl:366-452 and
MinAugHopf.jl:425-522) with every linear solve on the device -- PALC on the minimally augmented systems, the bordered solves
through bk_bls_matrixfree (Fold, device-resident state) and complex bk_gmres on a BK_COMPLEX context (Hopf)."""
import numpy as np
import pytest

import __graft_entry__ as g
from oracle import problems

pytestmark = pytest.mark.gpu
LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)


@pytest.fixture(scope="module")
def bk():
    return g.load_package()


def test_fold_curve_of_the_sh2d_hexagons_on_device(bk):
    """The hexagon branch of examples/SH2d-fronts.jl:88-92 turns near l = -0.215; its Fold point continued in the quadratic
    coefficient nu (second parameter) on 128 x 64: every point of the curve has F = 0 (oracle's sparse residual) and a singular
    Jacobian (smallest |eigenvalue| by shift-invert), and a fresh newton_fold at the last nu lands on the same l."""
    P, C2 = bk.palc, bk.codim2
    dims = (128, 64)
    ctx = bk.Context(bk.BK_SH2D, dims, (LX, LY), krylov_m=100, params=(-0.1, 1.3))
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    ls = bk.GMRESB200(reltol=1e-9, restart=100, maxiter=300, Pr=True, orth="cgs2")
    bls = bk.MatrixFreeBLSB200(ls)
    prob = P.BifurcationProblemB200(ctx, ctx.to_device(problems.sh2d_sol0(*dims, LX, LY)), (-0.1, 1.3), lens=0)
    pts = []
    cp = P.ContinuationPar(dsmin=1e-4, dsmax=5e-3, ds=-1e-3, p_min=-1.0, p_max=0.0, max_steps=70,
                           newton_options=P.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ls))
    P.continuation(prob, P.PALC(bls=bls), cp, normC=P.norminf, callback=lambda st: pts.append((st.z_u.copy(), st.z_p, st.tau_u.copy())) or True)
    ps = [p for _, p, _ in pts]
    k = next((i for i in range(1, len(ps) - 1) if ps[i] < ps[i + 1]), None)
    assert k is not None, ps
    x0, p0, tau = pts[k]
    tau.scale_(1.0 / tau.norm())
    nopt = P.NewtonPar(tol=1e-8, max_iterations=12, linsolver=ls)
    f0 = C2.newton_fold(prob, x0, p0, tau, tau, nopt, bls)
    assert f0.converged, f0.residuals
    cpf = P.ContinuationPar(dsmin=1e-4, dsmax=0.02, ds=0.01, p_min=1.0, p_max=1.6, max_steps=4, newton_options=nopt)
    seen = []
    curve = C2.continuation_fold(prob, f0.u, f0.p, 1, tau, tau, cpf, bls, normC=P.norminf,
                                 callback=lambda st: seen.append((st.z_u.u.copy(), st.z_u.p, st.z_p)) or True)
    assert prob.params == [-0.1, 1.3]
    l, nu = np.array(curve.p1), np.array(curve.p2)
    assert len(nu) == 5 and nu[0] == 1.3 and np.all(np.diff(nu) > 0) and abs(l[0] - f0.p) < 1e-7
    assert np.all(np.diff(l) < 0) and l[-1] < l[0] - 1e-3          # a stronger quadratic term moves the fold to smaller l
    assert max(r["itnewton"] for r in curve.rows) <= 8
    eig = bk.ShiftInvertB200(0.0, ls, krylovdim=30, tol=1e-8, maxrestart=10)
    for xk, lk, nuk in (seen[2], seen[-1]):
        sh = problems.SwiftHohenberg(dims, (LX, LY), l=lk, nu=nuk)
        assert np.max(np.abs(sh.F(xk.numpy(), lk))) < 1e-6
        prob.params[1] = nuk
        vals, _, cv, _ = eig(prob.J(xk, lk), 1)
        prob.params[1] = 1.3
        assert abs(vals[0]) < 1e-4, (nuk, vals)
    # an independent refinement at the last nu, started from the previous point of the curve
    xa, la, _ = seen[-2]
    prob.params[1] = nu[-1]
    b = curve.ma.b.copy()
    f1 = C2.newton_fold(prob, xa, la, b, b, nopt, bls)
    prob.params[1] = 1.3
    assert f1.converged and abs(f1.p - l[-1]) < 1e-6, (f1.p, l[-1], f1.residuals)


def test_hopf_curve_of_cgl2d_on_device(bk):
    """cGL2d 24 x 12 (examples/cGL2d.jl): the Hopf point of the trivial state, r = -lambda_1(Delta), omega = nu, continued in nu
    (third parameter): r stays at -lambda_1(Delta) and omega follows nu -- every complex shifted solve and J' application of the
    minimally augmented Hopf system on the BK_COMPLEX context, inside the PALC loop."""
    P, C2 = bk.palc, bk.codim2
    Nx, Ny = 24, 12
    gl = problems.GinzburgLandau2D(Nx, Ny, np.pi, np.pi / 2)
    n = gl.N
    rH, nu0 = gl.r_hopf(), gl.nu
    par = [rH + 0.05, gl.mu, nu0, gl.c3, gl.c5]
    rctx = bk.Context(bk.BK_CGL2D, (Nx, Ny), (np.pi, np.pi / 2), krylov_m=200, params=par)
    cctx = bk.Context(bk.BK_CGL2D, (Nx, Ny), (np.pi, np.pi / 2), krylov_m=300, params=par, complex=True)
    rctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)
    cctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)
    ls = bk.GMRESB200(reltol=1e-11, restart=200, maxiter=600, Pr=True, orth="cgs2")
    cls = bk.ComplexGMRESB200(reltol=1e-11, restart=300, maxiter=900, Pr=True, orth="cgs2")
    rng = np.random.default_rng(21)
    phi = gl.phi11()
    zeta = np.concatenate([phi, -1j * phi]) + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    u0 = 1e-4 * rng.standard_normal(n)
    prob = P.BifurcationProblemB200(rctx, u0, par, lens=0)
    cprob = C2.ComplexProblemB200(cctx, par, lens=0)
    cp = P.ContinuationPar(dsmin=1e-3, dsmax=0.1, ds=0.05, p_min=0.5, p_max=3.0, max_steps=4,
                           newton_options=P.NewtonPar(tol=1e-8, max_iterations=15, linsolver=ls))
    curve = C2.continuation_hopf(prob, cprob, u0, rH + 0.05, nu0 + 0.03, 2, zeta, zeta.copy(), cp, ls, cls)
    r, nu, om = np.array(curve.p1), np.array(curve.p2), np.array(curve.omega)
    assert len(nu) == 5 and nu[0] == nu0 and np.all(np.diff(nu) > 0) and not curve.stopped_at_bt
    assert np.max(np.abs(r - rH)) < 1e-6 and np.max(np.abs(om - nu)) < 1e-6, (r - rH, om - nu)
    assert np.linalg.norm(curve.state.z_u.u) < 1e-6
    assert prob.params[2] == nu0 and cprob.params[2] == nu0
