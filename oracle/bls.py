"""Bordered linear solvers (src/LinearBorderSolver.jl).  Test infrastructure only.

Solve   [ shift*I + J      dR     ] [dX]   [R]
        [ xiu*dzu'       xip*dzp  ] [dl] = [n]
"""
import numpy as np


def _apply(J, v):
    return J(v) if callable(J) else J @ v


class BorderingBLS:
    """src/LinearBorderSolver.jl:59-166 (BEC + k refinement rounds)."""

    def __init__(self, solver, tol=1e-12, check_precision=True, k=1):
        assert k > 0
        self.solver, self.tol, self.check_precision, self.k = solver, tol, check_precision, k

    def BEC(self, J, dR, dzu, dzp, R, n, xiu, xip, shift, dotp):
        # :125-144
        if shift is None:
            x1, dx, ok, it = self.solver(J, R, dR)
        else:
            x1, dx, ok, it = self.solver(J, R, dR, a0=shift)
        dl = (n - dotp(dzu, x1) * xiu) / (dzp * xip - dotp(dzu, dx) * xiu)
        x1 = x1 - dl * dx
        return x1, dl, ok, it

    def residualBEC(self, J, dR, dzu, dzp, R, n, dX, dl, xiu, xip, shift, dotp):
        # :146-166
        dXr = _apply(J, dX)
        if shift is not None:
            dXr = dXr + shift * dX
        dXr = dXr + dl * dR
        dXr = R - dXr
        dlr = n - xip * dzp * dl - xiu * dotp(dzu, dX)
        return dXr, dlr

    def __call__(self, J, dR, dzu, dzp, R, n, xiu=1.0, xip=1.0, shift=None, dotp=np.dot, apply_xiu=None):
        # :88-123
        dX, dl, cv, it = self.BEC(J, dR, dzu, dzp, R, n, xiu, xip, shift, dotp)
        k = 0
        fail = True
        while self.check_precision and k < self.k and fail:
            rX, rl = self.residualBEC(J, dR, dzu, dzp, R, n, dX, dl, xiu, xip, shift, dotp)
            fail = np.linalg.norm(rX) > self.tol or abs(rl) > self.tol
            if fail:
                dX1, dl1, cv, it = self.BEC(J, dR, dzu, dzp, rX, rl, xiu, xip, shift, dotp)
                dX = dX + dX1
                dl = dl + dl1
                k += 1
        return dX, dl, cv, it


class MatrixFreeBLSmap:
    """src/LinearBorderSolver.jl:299-335: x=[xu; xp] -> [J xu + xp a (+ shift xu); dot(b,xu) + c xp]."""

    def __init__(self, J, a, b, c, shift, dot):
        self.J, self.a, self.b, self.c, self.shift, self.dot = J, a, b, c, shift, dot

    def __call__(self, x):
        xu, xp = x[:-1], x[-1]
        out = np.empty_like(x)
        out[:-1] = _apply(self.J, xu) + xp * self.a
        if self.shift is not None:
            out[:-1] += self.shift * xu
        out[-1] = self.dot(self.b, xu) + self.c * xp
        return out


class MatrixFreeBLS:
    """src/LinearBorderSolver.jl:404-437 (use_bordered_array=false path: rhs = vcat(R, n))."""

    def __init__(self, solver):
        self.solver = solver

    def __call__(self, J, dR, dzu, dzp, R, n, xiu=1.0, xip=1.0, shift=None, dotp=np.dot, apply_xiu=None):
        lmap = MatrixFreeBLSmap(J, dR, dzu * xiu, dzp * xip, shift, dotp)
        rhs = np.concatenate([R, [n]])
        sol, cv, it = self.solver(lmap, rhs)
        return sol[:-1], sol[-1], cv, it


class MatrixBLS:
    """src/LinearBorderSolver.jl:217-264: assemble the (N+1)x(N+1) matrix, backslash."""

    def __call__(self, J, dR, dzu, dzp, R, n, xiu=1.0, xip=1.0, shift=None, dotp=None, apply_xiu=None):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spl
        N = len(R)
        row = xiu * dzu
        if apply_xiu is not None:  # :255-257 (e.g. row /= N for the PALC normalised dot)
            row = apply_xiu(row.copy())
        if sp.issparse(J):
            Js = J if shift is None else J + shift * sp.identity(N)
            A = sp.bmat([[Js, sp.csr_matrix(dR.reshape(-1, 1))],
                         [sp.csr_matrix(row.reshape(1, -1)), sp.csr_matrix([[xip * dzp]])]]).tocsc()
            sol = spl.spsolve(A, np.concatenate([R, [n]]))
        else:
            A = np.zeros((N + 1, N + 1))
            A[:N, :N] = np.asarray(J) + (0 if shift is None else shift * np.eye(N))
            A[:N, N] = dR
            A[N, :N] = row
            A[N, N] = xip * dzp
            sol = np.linalg.solve(A, np.concatenate([R, [n]]))
        return sol[:-1], sol[-1], True, 1


class MatrixFreeBLSmapBlock:
    """Tuple / block form, src/LinearBorderSolver.jl:366-389: x = [xu; xp], m = len(a) entries in xp;
    out.u = J xu + sum_i xp[i] a[i] (+ shift xu);  out.p = c xp + [dot(b[i], xu)]."""

    def __init__(self, J, a, b, c, shift, dot):
        self.J, self.a, self.b, self.c, self.shift, self.dot = J, tuple(a), tuple(b), np.atleast_2d(np.asarray(c, float)), shift, dot

    def __call__(self, x):
        m = len(self.a)
        xu, xp = x[:-m], x[-m:]
        out = np.empty_like(x)
        out[:-m] = _apply(self.J, xu)
        for i in range(m):
            out[:-m] += self.a[i] * xp[i]
        if self.shift is not None:
            out[:-m] += self.shift * xu
        out[-m:] = self.c @ xp
        for i in range(m):
            out[len(xu) + i] += self.dot(self.b[i], xu)
        return out


def solve_bls_block_bordering(solver, J, b, c, d, rhst, rhsb, shift=None):
    """solve_bls_block(lbs::BorderingBLS, J, b, c, d, rhst, rhsb), src/LinearBorderSolver.jl:173-206: b columns, c rows, d corner."""
    m = np.atleast_2d(d).shape[0]
    if not (len(b) == len(c) == m):
        raise ValueError("Linear bordered solver, wrong sizes!")
    kw = {} if shift is None else {"a0": shift}
    x1, cv, it = solver(J, rhst, **kw)
    x2s, its = [], []
    for bi in b:
        x2, flag, i2 = solver(J, bi, **kw)
        x2s.append(x2)
        its.append(i2)
        cv = cv and flag
    d = np.atleast_2d(np.asarray(d, float))
    S = np.array([[d[i, j] - np.dot(c[i], x2s[j]) for j in range(m)] for i in range(m)])
    h = np.array([rhsb[i] - np.dot(c[i], x1) for i in range(m)])
    u2 = np.linalg.solve(S, h)
    u1 = x1.copy()
    for i in range(m):
        u1 -= u2[i] * x2s[i]
    return u1, u2, cv, (it, *its)


def solve_bls_block_matrixfree(solver, J, a, b, c, rhst, rhsb, shift=None, dotp=np.dot):
    """solve_bls_block(lbs::MatrixFreeBLS, J, a, b, c, rhst, rhsb; shift, dotp), src/LinearBorderSolver.jl:440-450."""
    lmap = MatrixFreeBLSmapBlock(J, a, b, c, shift, dotp)
    m = len(a)
    sol, cv, it = solver(lmap, np.concatenate([rhst, np.atleast_1d(rhsb)]))
    return sol[:-m], sol[-m:], cv, it
