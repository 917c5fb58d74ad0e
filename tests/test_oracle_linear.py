"""Pins the oracle against the reference's own known-answer tests for the linear-solver
surfaces (test/linear_solvers/test_linear.jl).  Julia's RNG stream (Random.seed!(1234),
test_linear.jl:3) cannot be reproduced, so randomised cases use their own seeds and compare
against LAPACK dense solves exactly as the originals do."""
import numpy as np
import pytest

from oracle import krylov, bls


def _isapprox(a, b, rtol=np.sqrt(np.finfo(float).eps)):
    # Julia isapprox default: norm(a-b) <= rtol*max(norm(a), norm(b))
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def test_gmres_vs_dense_and_shift():
    """test_linear.jl:106-160"""
    rng = np.random.default_rng(1234)
    J0 = rng.random((100, 100)) * 0.1 - np.eye(100)
    rhs = rng.random(100)
    ls = krylov.GMRESIterativeSolvers(N=100, reltol=1e-16)
    sol, ok, it = ls(J0, rhs)
    assert _isapprox(sol, np.linalg.solve(J0, rhs))
    ref = np.linalg.solve(0.9 * J0 + 0.1 * np.eye(100), rhs)
    sol, _, _ = ls(J0, rhs, a0=0.1, a1=0.9)
    assert _isapprox(sol, ref)
    sol, _, _ = ls(lambda x: J0 @ x, rhs, a0=0.1, a1=0.9)
    assert _isapprox(sol, ref)
    for orth in ("cgs", "cgs2"):
        sol, ok, _ = krylov.GMRESIterativeSolvers(N=100, reltol=1e-12, orth=orth)(J0, rhs)
        assert _isapprox(sol, np.linalg.solve(J0, rhs))


def test_axpy_op_cases():
    """test_linear.jl:51-68"""
    rng = np.random.default_rng(0)
    J = rng.random((10, 10))
    v = rng.random(10)
    for a0, a1 in [(0.0, 1.0), (0.0, 0.3), (1.0, 1.0), (1.0, 0.3), (0.2, 0.3)]:
        assert np.allclose(krylov.axpy_op(J, a0, a1)(v), a0 * v + a1 * (J @ v), rtol=1e-14)
        assert np.allclose(krylov.axpy_op(lambda x: J @ x, a0, a1)(v), a0 * v + a1 * (J @ v), rtol=1e-14)


def test_gmres_preconditioned_left_right_restart():
    rng = np.random.default_rng(7)
    n = 60
    A = np.diag(np.linspace(1, 50, n)) + 0.1 * rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    Pinv = lambda r: r / np.diag(A)
    xref = np.linalg.solve(A, b)
    for kw in (dict(Pl=Pinv), dict(Pr=Pinv), dict(Pl=Pinv, restart=5, maxiter=400), dict(restart=7, maxiter=2000)):
        x, ok, it = krylov.gmres(A, b, reltol=1e-12, **{**dict(restart=200, maxiter=200), **kw})
        assert ok, kw
        assert np.linalg.norm(x - xref) <= 1e-9 * np.linalg.norm(xref), kw
    # maxiter cap -> not converged, iters == maxiter
    x, ok, it = krylov.gmres(A, b, reltol=1e-14, restart=200, maxiter=3)
    assert (not ok) and it == 3


def test_matrixfree_bls_map():
    """test_linear.jl:71-85: map == explicit bordered matrix * vector, with shifts."""
    rng = np.random.default_rng(3)
    n = 100
    J0 = rng.random((n, n))
    a, b, c = rng.random(n), rng.random(n), rng.random()
    x = rng.random(n + 1)
    for shift in (None, 0.3):
        A = np.zeros((n + 1, n + 1))
        A[:n, :n] = J0 + (0 if shift is None else shift * np.eye(n))
        A[:n, n] = a
        A[n, :n] = b
        A[n, n] = c
        m = bls.MatrixFreeBLSmap(J0, a, b, c, shift, np.dot)
        assert np.allclose(m(x), A @ x, rtol=1e-13)


@pytest.mark.parametrize("xiu,xip", [(1.0, 1.0), (0.37, 0.81)])
def test_bordered_solvers_vs_dense(xiu, xip):
    """test_linear.jl:172-244 incl. the (xiu, xip) cross-check :233-243."""
    rng = np.random.default_rng(11)
    n = 100
    J0 = rng.random((n, n)) * 0.1 - np.eye(n)
    dR, dzu, R = rng.random(n), rng.random(n), rng.random(n)
    dzp, nn = rng.random(), rng.random()
    for shift in (None, 0.12):
        A = np.zeros((n + 1, n + 1))
        A[:n, :n] = J0 + (0 if shift is None else shift * np.eye(n))
        A[:n, n] = dR
        A[n, :n] = xiu * dzu
        A[n, n] = xip * dzp
        ref = np.linalg.solve(A, np.concatenate([R, [nn]]))
        solvers = [bls.MatrixBLS(), bls.BorderingBLS(krylov.DefaultLS()),
                   bls.BorderingBLS(krylov.DefaultLS(), check_precision=True, k=2),
                   bls.BorderingBLS(krylov.GMRESIterativeSolvers(N=n, reltol=1e-14))]
        if shift is None:
            solvers.append(bls.MatrixFreeBLS(krylov.GMRESIterativeSolvers(N=n + 1, reltol=1e-14)))
        for s in solvers:
            dX, dl, ok, it = s(J0, dR, dzu, dzp, R, nn, xiu, xip, shift=shift)
            assert ok
            assert _isapprox(dX, ref[:n]) and abs(dl - ref[n]) <= 1e-8 * abs(ref[n]), type(s)


J5 = np.array([[0.688714, 0.363181, 0.956579, 0.967328, 0.950136],
               [0.860239, 0.0481349, 0.705687, 0.236736, 0.921345],
               [0.740663, 0.659207, 0.365235, 0.123933, 0.810514],
               [0.998672, 0.717179, 0.609523, 0.907024, 0.307781],
               [0.259797, 0.0059453, 0.105637, 0.218516, 0.356943]])
GOLD5 = np.array([2.750124876460063 + 0.0j,
                  0.2338099902832191 - 0.3203002738693372j, 0.2338099902832191 + 0.3203002738693372j,
                  -0.42584697851325004 - 0.17961985097997188j, -0.42584697851325004 + 0.17961985097997188j])


def test_golden_5x5_eigenvalues():
    """test_linear.jl:595-614 golden spectrum; Arnoldi oracle must reproduce it."""
    vals, vecs, cv, _ = krylov.arnoldi_eigs(lambda v: J5 @ v, 5, 5, krylovdim=5, seed=1)
    vals, vecs = krylov.sort_spectrum(vals, vecs)
    assert cv
    key = lambda z: (-round(z.real, 9), z.imag)
    assert np.allclose(sorted(vals, key=key), sorted(GOLD5, key=key), atol=1e-10)
    # eigenvector residuals
    for k in range(5):
        assert np.linalg.norm(J5 @ vecs[:, k] - vals[k] * vecs[:, k]) < 1e-9
    # sortedness (test_linear.jl:5)
    assert np.all(np.diff(vals.real) <= 1e-12)


def test_shift_invert_vs_eigvals():
    """test_linear.jl:666-673: ShiftInvert(0.1, DefaultLS) on I + 0.1 rand(10,10), |.|_inf < 1e-9"""
    rng = np.random.default_rng(5)
    J = np.eye(10) + 0.1 * rng.random((10, 10))
    eig = krylov.ShiftInvert(0.1, krylov.DefaultLS(), krylovdim=10, tol=1e-12)
    lam, vecs, cv, _ = eig(J, 10)
    ref = np.linalg.eigvals(J)
    ref = ref[np.lexsort((-ref.imag, -ref.real))]
    lam_s = lam[np.lexsort((-lam.imag, -lam.real))]
    assert np.max(np.abs(ref - lam_s)) < 1e-9
    assert np.all(np.diff(lam.real) <= 1e-12)
    # matrix-free inner solver (GMRES) gives the same spectrum
    eig2 = krylov.ShiftInvert(0.1, krylov.GMRESIterativeSolvers(N=10, reltol=1e-13), krylovdim=10, tol=1e-12)
    lam2 = eig2(lambda v: J @ v, 10, n=10)[0]
    lam2 = lam2[np.lexsort((-lam2.imag, -lam2.real))]
    assert np.max(np.abs(ref - lam2)) < 1e-8


def test_is_stable():
    ok, nu, ni = krylov.is_stable(np.array([0.1 + 0.2j, 0.1 - 0.2j, -1.0, 0.05]))
    assert (not ok) and nu == 3 and ni == 2


def test_dst_helmholtz_precond_inverts_shifted_dirichlet_laplacian():
    """oracle.precond.dst_helmholtz_precond == sparse solve with a0 I + a1 Lap (Dirichlet Laplacian of examples/cGL2d.jl:6-22)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from oracle import precond, problems
    nx, ny = 24, 17
    lap = problems.laplacian2d(nx, ny, np.pi, np.pi / 2, "dirichlet")
    v = np.random.default_rng(0).standard_normal(nx * ny)
    for a0, a1 in ((1.0, -0.3), (0.584, -0.32)):
        ref = spl.spsolve((a0 * sp.identity(nx * ny) + a1 * lap).tocsc(), v)
        got = precond.dst_helmholtz_precond(nx, ny, np.pi, np.pi / 2, a0, a1)(v)
        assert np.linalg.norm(got - ref) < 1e-13 * np.linalg.norm(ref)


def test_block_bordered_solvers_oracle_vs_dense():
    """solve_bls_block, both forms (src/LinearBorderSolver.jl:173-206, 440-450), as exercised by test/linear_solvers/test_linear.jl:300-320
    (random J, a, b, c): against the explicit (N + 2) x (N + 2) matrix; the tuple map (:366-389) against the same matrix."""
    obls = bls
    rng = np.random.default_rng(5)
    N = 40
    J = rng.standard_normal((N, N)) + 6 * np.eye(N)
    a = (rng.standard_normal(N), rng.standard_normal(N))
    b = (rng.standard_normal(N), rng.standard_normal(N))
    c = rng.standard_normal((2, 2))
    rhst, rhsb = rng.standard_normal(N), rng.standard_normal(2)
    for shift in (None, 0.4):
        A = np.block([[J + (0 if shift is None else shift) * np.eye(N), np.column_stack(a)], [np.vstack(b), c]])
        ex = np.linalg.solve(A, np.concatenate([rhst, rhsb]))
        u, p, cv, it = obls.solve_bls_block_bordering(krylov.DefaultLS(), J, a, b, c, rhst, rhsb, shift=shift)
        assert cv and np.allclose(u, ex[:N], atol=1e-10) and np.allclose(p, ex[N:], atol=1e-10)
        gm = krylov.GMRESIterativeSolvers(reltol=1e-12, restart=N + 2, maxiter=200)
        u, p, cv, it = obls.solve_bls_block_matrixfree(gm, J, a, b, c, rhst, rhsb, shift=shift)
        assert cv and np.allclose(u, ex[:N], atol=1e-8) and np.allclose(p, ex[N:], atol=1e-8)
        x = rng.standard_normal(N + 2)
        assert np.allclose(obls.MatrixFreeBLSmapBlock(J, a, b, c, shift, np.dot)(x), A @ x, atol=1e-12)
    with pytest.raises(ValueError):
        obls.solve_bls_block_bordering(krylov.DefaultLS(), J, a, b[:1], c, rhst, rhsb)
