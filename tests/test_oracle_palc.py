"""Pins the oracle's Newton / newton_palc / continuation / Trapeze / problem restatements against the
reference's known-answer tests (test/newton, test/continuation, test/periodic_orbits_function_fd)
and the structural identities of SURVEY section 8c.10."""
import numpy as np
import scipy.sparse as sp
import pytest

from oracle import krylov, bls, palc, problems, precond, potrap


def test_newton_root():
    """test/newton/test_newton.jl:4-21: root 1.234 of x^3 - 1.234^3."""
    rng = np.random.default_rng(0)
    x0 = np.ones(5) + 0.01 * rng.random(5)
    prob = palc.Problem(F=lambda x, p: x**3 - 1.234**3, J=lambda x, p: np.diag(3 * x**2), u0=x0, p0=0.0)
    sol = palc.newton(prob, x0, 0.0, palc.NewtonPar(tol=1e-13, max_iterations=8, linsolver=krylov.DefaultLS()), palc.norminf)
    assert sol.converged and np.allclose(sol.u, 1.234, rtol=1e-13)


def _simple_problem(n=10, p0=-1.5):
    F = lambda x, p: p * x + x**3 / 3 + 0.01
    J = lambda x, p: np.diag(p + x**2)
    return palc.Problem(F=F, J=J, u0=np.zeros(n), p0=p0)


def test_solve_bls_palc_vs_full_jacobian():
    """test/continuation/simple_continuation.jl:74-103: solve_bls_palc for the three BLS kinds
    equals the solve with the Jacobian of (F, N) -- pins the theta/N scaling of the arclength row."""
    prob = _simple_problem()
    opts = palc.NewtonPar(tol=1e-10, linsolver=krylov.DefaultLS())
    sol = palc.newton(prob, prob.u0, prob.p0, opts)
    n = 10
    theta, ds = 0.5, 1e-2
    tau_u, tau_p = 0.1 * np.ones(n), 1.0
    # analytic Jacobian of X=(u,p) -> (F(u,p), N(u,p))
    Jp = np.zeros((n + 1, n + 1))
    Jp[:n, :n] = prob.J(sol.u, prob.p0)
    Jp[:n, n] = sol.u  # dF/dp
    Jp[n, :n] = theta * tau_u / n
    Jp[n, n] = (1 - theta) * tau_p
    # check the row against finite differences of arc_length_eq
    Nf = lambda u, p: palc.arc_length_eq(u, sol.u, p - prob.p0, tau_u, tau_p, theta, ds)
    e = 1e-6
    for k in range(n):
        d = np.zeros(n); d[k] = e
        assert abs((Nf(sol.u + d, prob.p0) - Nf(sol.u - d, prob.p0)) / (2 * e) - Jp[n, k]) < 1e-9
    assert abs((Nf(sol.u, prob.p0 + e) - Nf(sol.u, prob.p0 - e)) / (2 * e) - Jp[n, n]) < 1e-9
    rng = np.random.default_rng(1)
    rhs = rng.random(n + 1)
    ref = np.linalg.solve(Jp, rhs)
    J0 = prob.J(sol.u, prob.p0)
    for s in (bls.MatrixBLS(), bls.BorderingBLS(krylov.DefaultLS()),
              bls.MatrixFreeBLS(krylov.GMRESIterativeSolvers(reltol=1e-12))):
        u, up, ok, it = palc.solve_bls_palc(s, theta, tau_u, tau_p, J0, sol.u.copy(), rhs[:n], rhs[n])
        assert np.allclose(u, ref[:n], rtol=1e-8, atol=1e-12) and abs(up - ref[n]) < 1e-8


def test_newton_palc_types_case():
    """test/newton/test_newton.jl:23-52: x^3 - 13x - p with delta=0.01, converges."""
    n = 10
    rng = np.random.default_rng(2)
    F = lambda x, p: x**3 - 13 * x - p
    J = lambda x, p: np.diag(3 * x**2 - 13)
    x0 = -0.04 * np.ones(n)
    prob = palc.Problem(F=F, J=J, u0=x0, p0=0.5, delta=0.01)
    cp = palc.ContinuationPar(newton_options=palc.NewtonPar(tol=1e-6, linsolver=krylov.DefaultLS()), ds=0.01, eta=10.0)
    sol = palc.newton_palc(prob, x0, 0.5, rng.random(n), 0.2, x0, 0.3, 0.01, 0.5, cp, bls.MatrixBLS())
    assert sol.converged


def test_continuation_fold_hits_pmin():
    """test/continuation/test-cont-non-vector.jl:22-45: r + x - x^3 from x=.8, r=1, ds=-0.02
    => br0.param[end] == -1 exactly (clamped Natural corrector on the boundary)."""
    prob = palc.Problem(F=lambda x, r: r + x - x**3, J=lambda x, r: np.diag(1 - 3 * x**2), u0=np.array([0.8]), p0=1.0,
                        record=lambda x: x[0])
    cp = palc.ContinuationPar(dsmin=0.001, dsmax=0.07, ds=-0.02, p_max=4.1, p_min=-1.0, max_steps=150,
                              newton_options=palc.NewtonPar(tol=1e-8, linsolver=krylov.DefaultLS()))
    rows, st = palc.continuation(prob, palc.PALC(bls=bls.MatrixBLS()), cp)
    assert rows[-1]["param"] == -1.0
    assert len(rows) > 10
    # the branch passes two folds of r = x^3 - x at x = +-1/sqrt(3): param is not monotone
    ps = np.array([r["param"] for r in rows])
    assert np.sum(np.diff(np.sign(np.diff(ps))) != 0) == 2
    # all points on the curve
    for r in rows:
        assert abs(r["param"] + r["x"] - r["x"] ** 3) < 1e-7


@pytest.mark.parametrize("tangent", ["secant", "bordered"])
@pytest.mark.parametrize("kind", ["matrix", "bordering", "matrixfree"])
def test_continuation_simple_all_bls(tangent, kind):
    """test/continuation/simple_continuation.jl:108-432 (length(br) > 10 for every predictor/BLS)."""
    prob = _simple_problem(n=10, p0=-1.5)
    ls = krylov.GMRESIterativeSolvers(reltol=1e-12)
    b = {"matrix": bls.MatrixBLS(), "bordering": bls.BorderingBLS(krylov.DefaultLS()), "matrixfree": bls.MatrixFreeBLS(ls)}[kind]
    cp = palc.ContinuationPar(dsmax=0.05, ds=0.01, p_min=-3.0, p_max=0.2, max_steps=60,
                              newton_options=palc.NewtonPar(tol=1e-9, linsolver=krylov.DefaultLS()))
    rows, st = palc.continuation(prob, palc.PALC(tangent=tangent, bls=b), cp)
    assert len(rows) > 10
    for r in rows[1:]:
        assert r["itnewton"] <= 25


def test_step_size_control_rule():
    cp = palc.ContinuationPar(dsmin=1e-3, dsmax=0.1, a=0.5, newton_options=palc.NewtonPar(max_iterations=10))
    ds, stop = palc.step_size_control(0.01, True, 2, cp)
    assert np.isclose(ds, 0.01 * (1 + 0.5 * 0.8**2)) and not stop
    ds, stop = palc.step_size_control(-0.01, False, 10, cp)
    assert np.isclose(ds, -0.005) and not stop
    ds, stop = palc.step_size_control(1e-3, False, 10, cp)
    assert stop
    ds, _ = palc.step_size_control(0.09, True, 0, cp)
    assert ds == 0.1


# ----------------------------------------------------------------------------- problems
def test_sh2d_two_pass_clamp_stencil_identity():
    """(I+Lap)^2 built by kron == two passes of the 5-pt stencil with clamp-to-edge."""
    Nx, Ny, lx, ly = 7, 5, 1.3, 0.9
    sh = problems.SwiftHohenberg((Nx, Ny), (lx, ly))
    rng = np.random.default_rng(0)
    u = rng.standard_normal(Nx * Ny)
    hx, hy = 2 * lx / Nx, 2 * ly / Ny

    def lap_clamp(a):
        A = a.reshape(Ny, Nx)
        P = np.pad(A, 1, mode="edge")
        return ((P[1:-1, :-2] - 2 * A + P[1:-1, 2:]) / hx**2 + (P[:-2, 1:-1] - 2 * A + P[2:, 1:-1]) / hy**2).reshape(-1)

    w = u + lap_clamp(u)
    L1u = w + lap_clamp(w)
    assert np.allclose(sh.L1 @ u, L1u, rtol=1e-13, atol=1e-10)


def test_dct_symbol_diagonalises_L1():
    """DCT-II diagonalises the Neumann-closure operator: (L1 + I)^-1 by DCT == sparse LU."""
    for dims, lens in (((7, 5), (1.3, 0.9)), ((6, 4, 5), (1.0, 1.2, 0.8))):
        sh = problems.SwiftHohenberg(dims, lens)
        rng = np.random.default_rng(1)
        r = rng.standard_normal(sh.N)
        a = precond.dct_precond(dims, lens, 1.0)(r)
        b = precond.sparse_lu_precond(sh.L1, 1.0)(r)
        assert np.allclose(a, b, rtol=1e-11, atol=1e-13)


def test_sh_jvp_is_derivative():
    sh = problems.SwiftHohenberg((9, 6), (2.0, 1.5))
    rng = np.random.default_rng(2)
    u, v = rng.standard_normal(sh.N), rng.standard_normal(sh.N)
    e = 1e-6
    fd = (sh.F(u + e * v) - sh.F(u - e * v)) / (2 * e)
    assert np.allclose(sh.dF(u, v), fd, rtol=1e-6, atol=1e-6)
    assert np.allclose(sh.jac_sparse(u) @ v, sh.dF(u, v), rtol=1e-12)


def test_chan_and_cgl_jvp_are_derivatives():
    rng = np.random.default_rng(3)
    x, dx = problems.chan_sol0(31), rng.standard_normal(31)
    e = 1e-6
    fd = (problems.chan_F(x + e * dx, 3.3, 0.01) - problems.chan_F(x - e * dx, 3.3, 0.01)) / (2 * e)
    assert np.allclose(problems.chan_dF(x, dx, 3.3, 0.01), fd, rtol=1e-6, atol=1e-5)
    g = problems.GinzburgLandau2D(6, 5, np.pi, np.pi / 2, r=1.2)
    u, du = 0.3 * rng.standard_normal(g.N), rng.standard_normal(g.N)
    fd = (g.F(u + e * du) - g.F(u - e * du)) / (2 * e)
    assert np.allclose(g.dF(u, du), fd, rtol=1e-6, atol=1e-6)
    # analytic Hopf point: largest eigenvalue of the linearisation at 0 crosses zero at r_hopf
    Lam = np.linalg.eigvalsh(g.lap.toarray())
    assert np.isclose(-Lam.max(), g.r_hopf(), rtol=1e-12)


def test_potrap_vs_reference_test_functional():
    """test/periodic_orbits_function_fd/test_potrap.jl:90-157 with F = x^2, J = 2 dx."""
    rng = np.random.default_rng(4)
    N, M = 12, 7
    F = lambda x: x**2
    dF = lambda x, dx: 2 * dx   # (sic) the reference test uses J = dx -> 2 dx
    phi, xpi = rng.random(N * M), rng.random(N * M)
    tr = potrap.Trapeze(F, dF, phi, xpi, M, N)
    x, dx = rng.random(N * M + 1), rng.random(N * M + 1)
    assert np.allclose(tr.residual(x), potrap.functional_ref(F, x, M, N, phi, xpi), rtol=1e-13)
    assert np.allclose(tr.jvp(x, dx), potrap.dfunctional_ref(F, dF, x, dx, M, N, phi), rtol=1e-13)
    # with a consistent Jacobian the jvp is the derivative of the residual
    dF2 = lambda x, dx: 2 * x * dx
    tr2 = potrap.Trapeze(F, dF2, phi, xpi, M, N)
    e = 1e-6
    fd = (tr2.residual(x + e * dx) - tr2.residual(x - e * dx)) / (2 * e)
    assert np.allclose(tr2.jvp(x, dx), fd, rtol=1e-6, atol=1e-7)


def test_chan_continuation_matrixfree_iterativesolvers():
    """Config 1 (plumbing): examples/chan.jl:97-118 -- matrix-free J, GMRESIterativeSolvers with
    Pl = lu(P), PALC(tangent=Bordered(), bls=BorderingBLS(lsp)); branch passes the fold near
    alpha ~ 3.9 and matches the dense-Jacobian branch."""
    n = 101
    beta = 0.01
    F = lambda x, a: problems.chan_F(x, a, beta)
    Jmf = lambda x, a: (lambda dx: problems.chan_dF(x, dx, a, beta))
    lsp = krylov.GMRESIterativeSolvers(reltol=1e-5, N=n, restart=20, maxiter=10, Pl=precond.chan_lu_precond(n))
    prob = palc.Problem(F=F, J=Jmf, u0=problems.chan_sol0(n), p0=3.3)
    cp = palc.ContinuationPar(dsmin=0.01, dsmax=0.5, ds=0.01, p_max=4.2, p_min=-1.0, max_steps=40,
                              newton_options=palc.NewtonPar(tol=1e-9, max_iterations=10, linsolver=lsp))
    rows, st = palc.continuation(prob, palc.PALC(tangent="bordered", bls=bls.BorderingBLS(lsp)), cp, normC=palc.norminf)
    assert len(rows) > 10
    # dense reference branch
    def Jd(x, a):
        E = np.eye(n)
        return np.column_stack([problems.chan_dF(x, E[:, k], a, beta) for k in range(n)])
    prob_d = palc.Problem(F=F, J=Jd, u0=problems.chan_sol0(n), p0=3.3)
    cpd = palc.ContinuationPar(dsmin=0.01, dsmax=0.5, ds=0.01, p_max=4.2, p_min=-1.0, max_steps=40,
                               newton_options=palc.NewtonPar(tol=1e-9, max_iterations=10, linsolver=krylov.DefaultLS()))
    rows_d, _ = palc.continuation(prob_d, palc.PALC(tangent="bordered", bls=bls.MatrixBLS()), cpd, normC=palc.norminf)
    pa = np.array([r["param"] for r in rows]); xa = np.array([r["x"] for r in rows])
    pd_ = np.array([r["param"] for r in rows_d]); xd = np.array([r["x"] for r in rows_d])
    m = min(len(pa), len(pd_))
    assert np.allclose(pa[:m], pd_[:m], rtol=1e-5, atol=1e-6)
    assert np.allclose(xa[:m], xd[:m], rtol=1e-5, atol=1e-6)
    assert pa.max() > 3.9  # reached the fold region


def test_potrap_circulant_preconditioner_makes_gmres_converge():
    """The time-circulant / DST preconditioner inverts the trivial-state PO Jacobian exactly and brings GMRES on the real
    PO Jacobian from 'no convergence in 60' down to ~10 iterations (stand-in for the ILU of examples/cGL2d.jl:209-213)."""
    Nx, Ny, M = 16, 8, 12
    L = (np.pi, np.pi / 2)
    gl = problems.GinzburgLandau2D(Nx, Ny, *L, r=1.0)
    gl.r = gl.r_hopf() + 0.05
    Ns = gl.N
    ph = gl.phi11()
    xs = np.concatenate([np.concatenate([0.3 * ph * np.cos(2 * np.pi * k / M), 0.3 * ph * np.sin(2 * np.pi * k / M)]) for k in range(M)] + [np.array([2 * np.pi])])
    N = len(xs)
    f1 = gl.F(xs[:Ns])
    phi = np.zeros(N - 1)
    phi[:Ns] = f1 / np.linalg.norm(f1)
    tr = potrap.Trapeze(gl.F, gl.dF, phi, np.zeros(N - 1), M, Ns)
    P = precond.potrap_circulant_precond(Nx, Ny, *L, M, xs[-1], gl.r, gl.nu)
    # exact inverse of the trivial-state PO Jacobian (first N-1 rows/cols)
    zero = np.concatenate([np.zeros(N - 1), [xs[-1]]])
    tr0 = potrap.Trapeze(gl.F, gl.dF, phi, np.zeros(N - 1), M, Ns)
    v = np.random.default_rng(0).standard_normal(N)
    v[-1] = 0.0
    Jv = tr0.jvp(zero, v)
    Jv[-1] = 0.0
    w = P(Jv)
    assert np.allclose(w[:-1], v[:-1], rtol=1e-9, atol=1e-10)
    rhs = tr.residual(xs)
    x0, ok0, it0 = krylov.gmres(lambda q: tr.jvp(xs, q), rhs, reltol=1e-6, restart=60, maxiter=60)
    x1, ok1, it1 = krylov.gmres(lambda q: tr.jvp(xs, q), rhs, Pr=P, reltol=1e-6, restart=60, maxiter=60)
    assert (not ok0) and ok1 and it1 <= 15
