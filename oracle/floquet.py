"""Floquet multipliers of a Trapeze periodic orbit, "quick and dirty" monodromy.  Test infrastructure only.

monodromy_matrix_free : src/periodicorbit/Floquet.jl:285-316  `MonodromyQaD_matrix_free(trap, u0, par, du)`
monodromy_dense       : src/periodicorbit/Floquet.jl:358-381  `MonodromyQaD(trap, J, u0, par)`
extract_eigenvector   : src/periodicorbit/Floquet.jl:319-355  `fl(Val(:ExtractEigenVector), ...)`
floquet_exponents     : src/periodicorbit/Floquet.jl:70-85    `compute_eigenvalues(fl::FloquetQaD, ...)`

Unknown x = [x_1; ...; x_M; T]; uniform mesh step 1/M (src/TimeMesh.jl:20-21), so h = T / M at every slice.
The linear solve has the reference's shifted contract  ls(J, rhs; a0, a1) -> (a0 I + a1 J)^-1 rhs
(src/LinearSolver.jl:8-12) with a0 = 1, a1 = -h/2.

Pinned by the reference's known answer test/periodic_orbits_function_fd/stuartLandauTrap.jl:84-93: on the
Stuart-Landau orbit the exponents are {0, -2 r T} (atol 5e-2 at M = 100) -- tests/test_oracle_floquet.py.
"""
import numpy as np


def _slices(x, M, N):
    return x[:-1].reshape(M, N)


def monodromy_matrix_free(apply_J, solve, x, M, N, du):
    """apply_J(u, v) = J(u) v;  solve(u, rhs, a0, a1) = (a0 I + a1 J(u))^-1 rhs.  0-based slice i = reference slice i+1."""
    T = x[-1]
    u = _slices(x, M, N)
    h = T / M
    out = np.array(du, dtype=float)
    out = out + (h / 2) * apply_J(u[M - 2], out)          # :300  slice M-1 (x_0 == x_{M-1})
    out = solve(u[0], out, 1.0, -h / 2)                   # :303
    for ii in range(1, M - 1):                            # :306-312  ii = 2..M-1
        out = out + (h / 2) * apply_J(u[ii - 1], out)
        out = solve(u[ii], out, 1.0, -h / 2)
    return out


def monodromy_dense(jac, x, M, N):
    """jac(u) -> dense J(u).  Product of (I - h/2 J_i)^-1 (I + h/2 J_{i-1}), i = 1..M-1 (:370-379)."""
    T = x[-1]
    u = _slices(x, M, N)
    h = T / M
    I = np.eye(N)
    mono = np.linalg.solve(I - h / 2 * jac(u[0]), I + h / 2 * jac(u[M - 2]))
    for ii in range(1, M - 1):
        mono = np.linalg.solve(I - h / 2 * jac(u[ii]), I + h / 2 * jac(u[ii - 1])) @ mono
    return mono


def extract_eigenvector(apply_J, solve, x, M, N, zeta):
    """Spatio-temporal eigenvector (M slices) from the Floquet eigenvector zeta (:319-355; note ii runs to M there)."""
    T = x[-1]
    u = _slices(x, M, N)
    h = T / M
    out = np.array(zeta)
    out = out + (h / 2) * apply_J(u[M - 2], out)
    out = solve(u[0], out, 1.0, -h / 2)
    res = [out.copy()]
    for ii in range(1, M):
        out = out + (h / 2) * apply_J(u[ii - 1], out)
        out = solve(u[ii], out, 1.0, -h / 2)
        res.append(out.copy())
    return res


def floquet_exponents(multipliers):
    """sigma = log(mu) sorted by decreasing real part (:77-82); the reference does not divide by T."""
    lv = np.log(np.asarray(multipliers, dtype=complex))
    order = np.argsort(-lv.real, kind="stable")
    return lv[order], order


def arnoldi_largest_modulus(op, n, nev, krylovdim=30, tol=1e-10, maxrestart=20, v0=None):
    """Explicitly restarted Arnoldi for the `nev` eigenvalues of largest modulus of a real operator (what the reference
    asks of its eigensolvers for Floquet multipliers, Floquet.jl:4-17 `which = :LM`).  Returns (vals, vecs, converged, nops)."""
    rng = np.random.default_rng(0)
    v = rng.standard_normal(n) if v0 is None else np.array(v0, dtype=float)
    nops = 0
    m = min(krylovdim, n)
    for _ in range(maxrestart + 1):
        Q = np.zeros((n, m + 1))
        H = np.zeros((m + 1, m))
        Q[:, 0] = v / np.linalg.norm(v)
        k_eff = m
        for k in range(m):
            w = op(Q[:, k]); nops += 1
            for _pass in range(2):
                c = Q[:, :k + 1].T @ w
                w = w - Q[:, :k + 1] @ c
                H[:k + 1, k] += c
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] < 1e-14 * max(1.0, np.abs(H[:k + 1, k]).max()):
                k_eff = k + 1
                break
            Q[:, k + 1] = w / H[k + 1, k]
        Hm = H[:k_eff, :k_eff]
        vals, Y = np.linalg.eig(Hm)
        order = np.argsort(-np.abs(vals), kind="stable")
        vals, Y = vals[order], Y[:, order]
        beta = H[k_eff, k_eff - 1] if k_eff < m + 1 and k_eff <= m else 0.0
        res = np.abs(beta * Y[k_eff - 1, :])
        want = min(nev, k_eff)
        if k_eff < m or np.all(res[:want] <= tol * np.maximum(np.abs(vals[:want]), 1e-300)):
            return vals[:want], Q[:, :k_eff] @ Y[:, :want], True, nops
        v = np.real(Q[:, :k_eff] @ Y[:, :want].sum(axis=1))  # restart with the sum of the wanted Ritz vectors
    return vals[:want], Q[:, :k_eff] @ Y[:, :want], False, nops
