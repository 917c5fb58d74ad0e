"""GPU parity tests of the complexified path (BK_COMPLEX contexts, include/bk200.h): complex shifts and J' through the C ABI
against dense NumPy algebra on the oracle's Jacobians, the complex bordered known answer of the reference
(test/linear_solvers/test_linear.jl:324-351 pattern) and the Hopf minimally augmented Newton (src/codim2/MinAugHopf.jl).
Tolerances: operator applications 1e-12 relative; solves 1e-8 relative to the dense solution."""
import numpy as np
import pytest

import __graft_entry__ as g
from oracle import problems

pytestmark = pytest.mark.gpu

LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)
CGL = (1.2, 0.1, 1.0, -1.0, 1.0)


@pytest.fixture(scope="module")
def bk():
    return g.load_package()


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def _dense(dF, u, n):
    return np.column_stack([dF(u, np.eye(n)[:, j]) for j in range(n)])


def test_complex_apply_shift_and_transpose(bk):
    """((a0 + i a0i) I + a1 J) z and J' z on split complex vectors: cGL2d 24 x 12 (non-symmetric reaction block) and SH2d 32 x 16
    (self-adjoint: J' = J, served by the TMA tile kernel on both halves)."""
    rng = np.random.default_rng(11)
    gl = problems.GinzburgLandau2D(24, 12, np.pi, np.pi / 2, r=1.2)
    u = 0.4 * rng.standard_normal(gl.N)
    z = rng.standard_normal(gl.N) + 1j * rng.standard_normal(gl.N)
    Jd = _dense(gl.dF, u, gl.N)
    ctx = bk.Context(bk.BK_CGL2D, (24, 12), (np.pi, np.pi / 2), krylov_m=8, params=CGL, complex=True)
    assert ctx.N == 2 * gl.N and ctx.N0 == gl.N
    assert _rel(ctx.residual(u), gl.F(u)) < 1e-12                      # F stays the real functional
    assert _rel(ctx.cjacobian(u)(z), Jd @ z) < 1e-12
    assert _rel(ctx.cjacobian(u, transpose=True)(z), Jd.T @ z) < 1e-12
    ctx.set_transpose(False)
    ctx.set_shift_imag(-0.7)
    out = bk.core.cjoin(ctx.jvp(bk.core.csplit(z), a0=0.3, a1=0.9))
    assert _rel(out, (0.3 - 0.7j) * z + 0.9 * (Jd @ z)) < 1e-12
    ctx.set_shift_imag(0.0)
    # the real context refuses an imaginary shift, Chan refuses J'
    rc = bk.Context(bk.BK_CHAN, (31,), (1.0,), krylov_m=4, params=(3.3, 0.01))
    with pytest.raises(bk.BK200Error):
        rc.set_shift_imag(0.5)
    with pytest.raises(bk.BK200Error):
        rc.set_transpose(True)
    sh = problems.SwiftHohenberg((32, 16), (LX, LY), l=-0.1, nu=1.3)
    us = problems.sh2d_sol0(32, 16, LX, LY) + 0.1 * rng.standard_normal(sh.N)
    zs = rng.standard_normal(sh.N) + 1j * rng.standard_normal(sh.N)
    cs = bk.Context(bk.BK_SH2D, (32, 16), (LX, LY), krylov_m=8, params=(-0.1, 1.3), complex=True)
    ref = sh.dF(us, zs.real) + 1j * sh.dF(us, zs.imag)
    assert _rel(cs.cjacobian(us)(zs), ref) < 1e-12
    assert _rel(cs.cjacobian(us, transpose=True)(zs), ref) < 1e-12


@pytest.mark.parametrize("transpose", [False, True])
def test_complex_gmres_vs_dense(bk, transpose):
    """(a0 I + J) x = rhs with complex a0, rhs: cGL2d unpreconditioned and with the Helmholtz DST preconditioner on both halves."""
    rng = np.random.default_rng(12)
    gl = problems.GinzburgLandau2D(24, 12, np.pi, np.pi / 2, r=1.2)
    n = gl.N
    u = 0.4 * rng.standard_normal(n)
    Jd = _dense(gl.dF, u, n)
    A = Jd.T if transpose else Jd
    rhs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a0 = 0.25 - 0.9j
    xd = np.linalg.solve(a0 * np.eye(n) + A, rhs)
    ctx = bk.Context(bk.BK_CGL2D, (24, 12), (np.pi, np.pi / 2), krylov_m=300, params=CGL, complex=True)
    Jc = ctx.cjacobian(u, transpose)
    ls = bk.ComplexGMRESB200(reltol=1e-11, restart=300, maxiter=900, orth="cgs2")
    x, cv, it = ls(Jc, rhs, a0=a0)
    assert cv and _rel(x, xd) < 1e-8, (cv, it, _rel(x, xd))
    ctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)               # (Delta - I)^-1 on every component of both halves
    lp = bk.ComplexGMRESB200(reltol=1e-11, restart=300, maxiter=900, Pr=True, orth="cgs2")
    xp, cvp, itp = lp(Jc, rhs, a0=a0)
    assert cvp and _rel(xp, xd) < 1e-8 and itp < it, (cvp, itp, it)


def test_complex_gmres_sh2d_preconditioned(bk):
    """SH2d 64 x 32 with the DCT preconditioner: (J - i omega) x = rhs."""
    rng = np.random.default_rng(13)
    dims = (64, 32)
    sh = problems.SwiftHohenberg(dims, (LX, LY), l=-0.1, nu=1.3)
    us = problems.sh2d_sol0(*dims, LX, LY)
    n = sh.N
    Jd = _dense(sh.dF, us, n)
    rhs = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    xd = np.linalg.solve(Jd - 0.6j * np.eye(n), rhs)
    ctx = bk.Context(bk.BK_SH2D, dims, (LX, LY), krylov_m=120, params=(-0.1, 1.3), complex=True)
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    ls = bk.ComplexGMRESB200(reltol=1e-11, restart=120, maxiter=600, Pr=True, orth="cgs2")
    x, cv, it = ls(ctx.cjacobian(us), rhs, a0=-0.6j)
    assert cv and _rel(x, xd) < 1e-8, (cv, it, _rel(x, xd))


def test_hopf_border_and_newton_on_device(bk):
    """cGL2d 24 x 12 (examples/cGL2d.jl: Hopf bifurcation of the trivial state at r = -lambda_1(Delta), omega = nu):
    (i) the bordered solve [J - i omega, a; b^H, 0] [v; sigma] = [0; 1] by bordering over complex bk_gmres against the explicit
    dense solve; (ii) newton_hopf from a perturbed guess."""
    P = bk.palc
    Nx, Ny = 24, 12
    gl = problems.GinzburgLandau2D(Nx, Ny, np.pi, np.pi / 2)
    n = gl.N
    rH, nu = gl.r_hopf(), gl.nu
    par = (rH + 0.3, gl.mu, gl.nu, gl.c3, gl.c5)
    rctx = bk.Context(bk.BK_CGL2D, (Nx, Ny), (np.pi, np.pi / 2), krylov_m=200, params=par)
    cctx = bk.Context(bk.BK_CGL2D, (Nx, Ny), (np.pi, np.pi / 2), krylov_m=300, params=par, complex=True)
    rctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)
    cctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)
    ls = bk.GMRESB200(reltol=1e-11, restart=200, maxiter=600, Pr=True, orth="cgs2")
    cls = bk.ComplexGMRESB200(reltol=1e-11, restart=300, maxiter=900, Pr=True, orth="cgs2")
    rng = np.random.default_rng(14)
    phi = gl.phi11()
    zeta = np.concatenate([phi, -1j * phi]) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    u0 = 1e-3 * rng.standard_normal(n)
    prob = P.BifurcationProblemB200(rctx, u0, par, lens=0)
    cprob = bk.codim2.ComplexProblemB200(cctx, par, lens=0)
    # (i)
    ma = bk.codim2.HopfMinAug(prob, cprob, zeta, zeta, ls, cls)
    om = nu + 0.2
    v, sigma = ma._border(cprob.J(u0, rH + 0.3), complex(0.0, -om), zeta, zeta)
    Jd = _dense(lambda u, d: gl.dF(u, d, rH + 0.3), u0, n)
    J0 = np.block([[Jd - 1j * om * np.eye(n), zeta[:, None]], [np.conj(zeta)[None, :], np.zeros((1, 1))]])
    rhs = np.zeros(n + 1, dtype=complex)
    rhs[-1] = 1
    ex = np.linalg.solve(J0, rhs)
    assert _rel(v, ex[:-1]) < 1e-8 and abs(sigma - ex[-1]) < 1e-8 * abs(ex[-1])
    # (ii)
    sol = bk.codim2.newton_hopf(prob, cprob, u0, rH + 0.3, om, zeta, zeta.copy(),
                                P.NewtonPar(tol=1e-8, max_iterations=15, linsolver=ls), ls, cls)
    assert sol.converged, sol.residuals
    assert abs(sol.p - rH) < 1e-6 and abs(sol.omega - nu) < 1e-6 and np.linalg.norm(sol.u) < 1e-7, (sol.p - rH, sol.omega - nu)
