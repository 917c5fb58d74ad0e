"""GMRES (IterativeSolvers.jl semantics) and Arnoldi eigensolver restatements.
Test infrastructure only (see oracle/__init__.py).

The reference calls ``IterativeSolvers.gmres`` at src/LinearSolver.jl:198-201;
that package is not vendored (Project.toml:53, compat 0.8.4/0.8.5/^0.9, no
Manifest).  Its published algorithm (v0.9 ``gmres.jl``) restated here:
  * left/right preconditioners Pl, Pr: Krylov vectors v_{k+1} = Pl \\ (A (Pr \\ v_k));
  * modified Gram-Schmidt;
  * tolerance = max(reltol * ||Pl \\ r0||, abstol) on the *preconditioned* residual
    tracked through Givens rotations;
  * ``iters`` = total number of inner iterations, capped by ``maxiter``;
  * ``initially_zero=true`` skips the first mat-vec;
  * x is updated at restart and at the end: x += Pr \\ (V y).
Iterates/iteration counts: parity unpinned (reference tests pin only x ~ A\\b).
"""
import numpy as np


def _ident(x):
    return x


def axpy_op(J, a0=0.0, a1=1.0):
    """v -> a0 v + a1 J v  (src/LinearSolver.jl:46-62 _axpy_op)."""
    apply = J if callable(J) else (lambda v: J @ v)
    if a0 == 0.0 and a1 == 1.0:
        return apply
    return lambda v: a0 * v + a1 * apply(v)


def gmres(A, b, *, Pl=None, Pr=None, abstol=0.0, reltol=1e-8, restart=200, maxiter=100,
          initially_zero=True, x0=None, orth="mgs", history=None):
    """Returns (x, isconverged, iters).  A, Pl, Pr are callables (Pl/Pr apply the
    INVERSE of the preconditioner, i.e. ``Pl(r) = Pl \\ r``)."""
    Pl = Pl or _ident
    Pr = Pr or _ident
    if not callable(A):
        M = A
        A = lambda v: M @ v
    n = b.shape[0]
    x = np.zeros_like(b) if (x0 is None or initially_zero) else x0.copy()
    restart = min(restart, n)
    V = np.empty((restart + 1, n), dtype=b.dtype)
    H = np.zeros((restart + 1, restart), dtype=b.dtype)

    def init_residual(first):
        r = b.copy() if (first and initially_zero) else b - A(x)
        r = Pl(r)
        beta = np.linalg.norm(r)
        V[0] = r / beta if beta > 0 else r
        return beta

    beta = init_residual(True)
    tol = max(reltol * beta, abstol)
    res = beta
    total = 0
    if history is not None:
        history.append(res)
    while total < maxiter and res > tol:
        # one restart cycle
        g = np.zeros(restart + 1, dtype=b.dtype)
        g[0] = beta
        cs = np.zeros(restart, dtype=b.dtype)
        sn = np.zeros(restart, dtype=b.dtype)
        k = 0
        while k < restart and total < maxiter and res > tol:
            w = Pl(A(Pr(V[k])))
            if orth == "mgs":
                for i in range(k + 1):
                    H[i, k] = np.dot(V[i], w)
                    w = w - H[i, k] * V[i]
            elif orth == "cgs":
                h = V[: k + 1] @ w
                w = w - V[: k + 1].T @ h
                H[: k + 1, k] = h
            elif orth == "cgs2":
                h = V[: k + 1] @ w
                w = w - V[: k + 1].T @ h
                h2 = V[: k + 1] @ w
                w = w - V[: k + 1].T @ h2
                H[: k + 1, k] = h + h2
            else:
                raise ValueError(orth)
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] != 0:
                V[k + 1] = w / H[k + 1, k]
            # apply previous Givens rotations to the new column
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k] = H[k, k] / d
            sn[k] = H[k + 1, k] / d
            H[k, k] = d
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            res = abs(g[k + 1])
            k += 1
            total += 1
            if history is not None:
                history.append(res)
        # solve the triangular system and update x
        if k > 0:
            y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k > 1 else g[:1] / H[0, 0]
            x = x + Pr(V[:k].T @ y)
        if total < maxiter and res > tol:
            beta = init_residual(False)
            res = beta
    return x, bool(res <= tol), total


class GMRESIterativeSolvers:
    """Mirror of src/LinearSolver.jl:149-206: call ``ls(J, rhs; a0, a1) -> (x, ok, iters)``;
    two-rhs form (src/LinearSolver.jl:15-19) -> (x1, x2, ok1&ok2, (it1, it2))."""

    def __init__(self, abstol=0.0, reltol=1e-8, restart=200, maxiter=100, N=0,
                 initially_zero=True, Pl=None, Pr=None, orth="mgs"):
        self.abstol, self.reltol, self.restart, self.maxiter = abstol, reltol, restart, maxiter
        self.N, self.initially_zero, self.Pl, self.Pr, self.orth = N, initially_zero, Pl, Pr, orth

    def __call__(self, J, rhs, rhs2=None, a0=0.0, a1=1.0):
        if rhs2 is not None:
            x1, ok1, it1 = self(J, rhs, a0=a0, a1=a1)
            x2, ok2, it2 = self(J, rhs2, a0=a0, a1=a1)
            return x1, x2, ok1 and ok2, (it1, it2)
        op = axpy_op(J, a0, a1)
        return gmres(op, rhs, Pl=self.Pl, Pr=self.Pr, abstol=self.abstol, reltol=self.reltol,
                     restart=self.restart, maxiter=self.maxiter, initially_zero=self.initially_zero,
                     orth=self.orth)


class DefaultLS:
    """src/LinearSolver.jl:94-117 backslash solver (dense or sparse matrix J)."""

    def __call__(self, J, rhs, rhs2=None, a0=0.0, a1=1.0):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spl
        if sp.issparse(J):
            M = (a0 * sp.identity(J.shape[0]) + a1 * J).tocsc()
            lu = spl.splu(M)
            solve = lu.solve
        else:
            M = a0 * np.eye(J.shape[0]) + a1 * np.asarray(J)
            solve = lambda r: np.linalg.solve(M, r)
        if rhs2 is not None:
            return solve(rhs), solve(rhs2), True, (1, 1)
        return solve(rhs), True, 1


# --------------------------------------------------------------------------- eigen
def arnoldi_eigs(op, n, nev, *, krylovdim=None, tol=1e-10, maxrestart=20, v0=None, which="LM", seed=0):
    """Explicitly restarted Arnoldi (restart vector = sum of wanted Ritz vectors) with
    full re-orthogonalisation (CGS2).  Stand-in for ArnoldiMethod.partialschur /
    KrylovKit.eigsolve (src/EigSolver.jl:157-160,204-225; external packages).
    Returns (vals, vecs[n, nev], converged, n_opapplies)."""
    m = krylovdim or max(30, nev + 30)
    m = min(m, n)
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n) if v0 is None else np.array(v0, dtype=float)
    nops = 0
    vals = vecs = None
    for _ in range(maxrestart):
        V = np.zeros((m + 1, n))
        H = np.zeros((m + 1, m))
        V[0] = v / np.linalg.norm(v)
        k_eff = m
        for k in range(m):
            w = op(V[k])
            nops += 1
            h = V[: k + 1] @ w
            w = w - V[: k + 1].T @ h
            h2 = V[: k + 1] @ w
            w = w - V[: k + 1].T @ h2
            H[: k + 1, k] = h + h2
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] < 1e-14 * max(1.0, np.linalg.norm(H[: k + 1, k])):
                k_eff = k + 1
                break
            V[k + 1] = w / H[k + 1, k]
        Hm = H[:k_eff, :k_eff]
        theta, S = np.linalg.eig(Hm)
        order = np.argsort(-np.abs(theta)) if which == "LM" else np.argsort(-theta.real)
        theta, S = theta[order], S[:, order]
        nv = min(nev, k_eff)
        resid = np.abs(H[k_eff, k_eff - 1] * S[-1, :nv]) if k_eff < m + 1 and k_eff == m else np.zeros(nv)
        vals = theta[:nv]
        vecs = V[:k_eff].T @ S[:, :nv]
        if np.all(resid <= tol * np.maximum(np.abs(vals), 1e-300)) or k_eff < m:
            return vals, vecs, True, nops
        v = np.real(vecs @ np.ones(nv))
    return vals, vecs, False, nops


def sort_spectrum(vals, vecs):
    """src/EigSolver.jl:16-19: sort by decreasing real part."""
    idx = np.argsort(-np.real(vals), kind="stable")
    return vals[idx], (vecs[:, idx] if vecs is not None else None)


class ShiftInvert:
    """src/EigSolver.jl:246-266: eigen-elements of (J - sigma I)^-1 via ls(J, rhs; a0=-sigma, a1=1),
    mapped back lambda = sigma + 1/theta, sorted by decreasing real part."""

    def __init__(self, sigma, ls, krylovdim=None, tol=1e-10, maxrestart=20):
        self.sigma, self.ls, self.krylovdim, self.tol, self.maxrestart = sigma, ls, krylovdim, tol, maxrestart

    def __call__(self, J, nev, n=None, v0=None, seed=0):
        if n is None:
            n = J.shape[0]
        Jmap = lambda rhs: self.ls(J, rhs, a0=-self.sigma, a1=1.0)[0]
        vals, vecs, cv, nops = arnoldi_eigs(Jmap, n, nev, krylovdim=self.krylovdim, tol=self.tol,
                                            maxrestart=self.maxrestart, v0=v0, seed=seed)
        lam = 1.0 / vals + self.sigma
        lam, vecs = sort_spectrum(lam, vecs)
        return lam, vecs, cv, nops


def is_stable(eigvalues, tol_stability=1e-10):
    """src/Bifurcations.jl:5-18 -> (isstable, n_unstable, n_imag)."""
    ev = np.asarray(eigvalues)
    n_unstable = int(np.sum(ev.real > tol_stability))
    n_imag = int(np.sum((np.abs(ev.imag) > tol_stability) & (ev.real > tol_stability)))
    return n_unstable == 0, n_unstable, n_imag
