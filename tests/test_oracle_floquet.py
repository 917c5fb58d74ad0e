"""Oracle for the Floquet "quick and dirty" monodromy, pinned to the reference's known answer
test/periodic_orbits_function_fd/stuartLandauTrap.jl:84-93 (exponents {0, -2 r T}, atol 5e-2 at M = 100)."""
import numpy as np

from oracle import floquet, potrap, problems


def _stuart_landau(r, mu=0.0, nu=1.0, c3=1.0):
    def F(u):
        ua = u[0] ** 2 + u[1] ** 2
        return np.array([r * u[0] - nu * u[1] - ua * (c3 * u[0] - mu * u[1]), r * u[1] + nu * u[0] - ua * (c3 * u[1] + mu * u[0])])

    def jac(u):
        u1, u2 = u
        ua = u1 * u1 + u2 * u2
        return np.array([[r - 2 * u1 * (c3 * u1 - mu * u2) - ua * c3, -nu - 2 * u2 * (c3 * u1 - mu * u2) + ua * mu],
                         [nu - 2 * u1 * (c3 * u2 + mu * u1) - ua * mu, r - 2 * u2 * (c3 * u2 + mu * u1) - ua * c3]])

    return F, jac


def _newton_po(tr, x, tol=1e-9, maxit=15):
    n = len(x)
    for _ in range(maxit):
        f = tr.residual(x)
        if np.max(np.abs(f)) < tol:
            return x, True
        Jm = np.column_stack([tr.jvp(x, e) for e in np.eye(n)])
        x = x - np.linalg.solve(Jm, f)
    return x, np.max(np.abs(tr.residual(x))) < tol


def test_stuart_landau_floquet_exponents_known_answer():
    r, M, N = 0.1, 100, 2     # par_hopf, Trapeze(prob2, [1, 0], zeros(2), 100, 2)   stuartLandauTrap.jl:24,36-40
    F, jac = _stuart_landau(r)
    phi = np.zeros(N * M); phi[0] = 1.0
    tr = potrap.Trapeze(F, lambda u, du: jac(u) @ du, phi, np.zeros(N * M), M, N)
    th = np.linspace(0, 2 * np.pi, M)
    x0 = np.concatenate([np.sqrt(r) * np.column_stack([np.cos(th), np.sin(th)]).reshape(-1), [2 * np.pi]])  # :49-50
    x, ok = _newton_po(tr, x0)
    assert ok and abs(x[-1] - 2 * np.pi) < 0.1
    T = x[-1]
    mono = floquet.monodromy_dense(jac, x, M, N)
    sig, _ = floquet.floquet_exponents(np.linalg.eigvals(mono))
    assert abs(sig[0]) < 1e-6                                   # the trivial exponent
    assert abs(sig.real.min() - (-2 * r * T)) < 5e-2            # reference tolerance (:93), "~1/M"
    # matrix-free application == dense product, column by column
    apply_J = lambda u, v: jac(u) @ v
    solve = lambda u, rhs, a0, a1: np.linalg.solve(a0 * np.eye(N) + a1 * jac(u), rhs)
    mf = np.column_stack([floquet.monodromy_matrix_free(apply_J, solve, x, M, N, e) for e in np.eye(N)])
    assert np.allclose(mf, mono, rtol=1e-12, atol=1e-13)
    # eigenvector extraction: slice k of the spatio-temporal vector is the partial product applied to zeta; the last one
    # (ii = M) applies one more factor than the monodromy
    vals, vecs = np.linalg.eig(mono)
    z = np.real(vecs[:, np.argmax(np.abs(vals))])
    sl = floquet.extract_eigenvector(apply_J, solve, x, M, N, z)
    assert len(sl) == M and np.allclose(sl[M - 2], mono @ z, rtol=1e-10)


def test_cgl_monodromy_and_arnoldi_largest_modulus():
    """cGL2d (config 4's vector field) on a small grid: matrix-free monodromy with sparse solves == dense product, and the
    Arnoldi used for `which = :LM` (Floquet.jl:4-17) finds the dominant multipliers."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    nx, ny, M = 8, 6, 12
    g = problems.GinzburgLandau2D(nx, ny, np.pi, np.pi / 2, r=1.3, mu=0.1, nu=1.0, c3=-1.0, c5=1.0)
    N = g.N
    rng = np.random.default_rng(3)
    i = np.arange(1, nx + 1); j = np.arange(1, ny + 1)
    phi11 = (np.sin(np.pi * i / (nx + 1))[None, :] * np.sin(np.pi * j / (ny + 1))[:, None]).reshape(-1)
    x = np.concatenate([np.concatenate([0.4 * phi11 * np.cos(2 * np.pi * k / M), 0.4 * phi11 * np.sin(2 * np.pi * k / M)]) for k in range(M)] + [np.array([6.5])])
    jac = lambda u: np.column_stack([g.dF(u, e) for e in np.eye(N)])
    apply_J = lambda u, v: g.dF(u, v)
    solve = lambda u, rhs, a0, a1: spl.spsolve(sp.csc_matrix(a0 * np.eye(N) + a1 * jac(u)), rhs)
    mono = floquet.monodromy_dense(jac, x, M, N)
    v = rng.standard_normal(N)
    assert np.allclose(floquet.monodromy_matrix_free(apply_J, solve, x, M, N, v), mono @ v, rtol=1e-9, atol=1e-12)
    ref = np.linalg.eigvals(mono)
    ref = ref[np.argsort(-np.abs(ref))]
    vals, vecs, cv, nops = floquet.arnoldi_largest_modulus(lambda q: mono @ q, N, 4, krylovdim=30, tol=1e-10)
    assert cv
    assert np.allclose(np.sort(np.abs(vals))[::-1], np.abs(ref[:4]), rtol=1e-8)
    for k in range(4):
        assert np.linalg.norm(mono @ vecs[:, k] - vals[k] * vecs[:, k]) < 1e-7 * abs(vals[k])
