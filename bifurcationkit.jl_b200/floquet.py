"""Floquet multipliers of a Trapeze periodic orbit, matrix-free -- host orchestration over the same C ABI.

* ``FloquetQaDB200``  <-> ``FloquetQaD`` (src/periodicorbit/Floquet.jl:46-85):  ``fl(x, nev) -> (sigma, vecs, converged, info)``
  with ``sigma = log(mu)`` sorted by decreasing real part (:77-82).  The monodromy operator is the reference's
  ``MonodromyQaD_matrix_free`` (:285-316): M-1 factors ``(I - h/2 J_i)^-1 (I + h/2 J_{i-1})``, each one shifted JVP
  (``bk_jvp`` with a0 = 1, a1 = h/2) and one shifted solve through the *same* linear-solver contract
  ``ls(J, rhs; a0 = 1, a1 = -h/2)`` (:303,310) -> ``bk_gmres``.  ``extract_eigenvector`` is :319-355.
* ``ArnoldiLMB200``   <-> the eigensolver the reference requires for this job, "largest modulus" (``_check_floquet_options``
  :4-17): Arnoldi on a user operator, CGS2 with the library's vector kernels, small Hessenberg problem through
  ``bk_hessenberg_eig``; contract ``eig(op, x0, nev) -> (vals, vecs, converged, nops)`` (src/EigSolver.jl:4-19).

The vector field lives in its own context (``BK_CGL2D``: one Jacobian per context, re-linearised at every slice as the
reference does, ``jacobian(trap.prob_vf, u0c[:, ii], par)``); the orbit ``x = [x_1 .. x_M; T]`` may be a NumPy array or a
``DeviceVec`` of the ``BK_POTRAP_CGL2D`` context -- slices are passed as raw pointers, nothing is copied.
"""
import numpy as np

from . import core as _core
from . import lib as _l
from .core import DeviceVec
from .palc import V


class _Slice:
    """Non-owning view of ``n`` entries of a DeviceVec starting at ``offset`` (what ``get_time_slices`` returns, a view)."""

    def __init__(self, vec, offset, n):
        self.ctx, self.n, self.dptr, self._keep = vec.ctx, int(n), vec.dptr + 8 * int(offset), vec

    def __len__(self):
        return self.n


def time_slice(x, i, N):
    """u0c[:, i+1] of ``get_time_slices(u0, N, M)`` (0-based i)."""
    if isinstance(x, DeviceVec):
        return _Slice(x, i * N, N)
    return x[i * N:(i + 1) * N]


def period(x):
    """getperiod(trap, u0) = u0[end]."""
    if isinstance(x, DeviceVec):
        out = np.empty(1)
        _core._chk(x.ctx, x.ctx.lib.bk_vec_download(x.ctx.handle, out.ctypes.data, x.dptr + 8 * (x.n - 1), 1))
        return float(out[0])
    return float(x[-1])


class ArnoldiLMB200:
    """Explicitly restarted Arnoldi for the eigenvalues of largest modulus of a real operator given as a closure."""

    def __init__(self, krylovdim=30, tol=1e-8, maxrestart=10):
        self.krylovdim, self.tol, self.maxrestart = krylovdim, tol, maxrestart

    def __call__(self, op, x0, nev):
        n = len(x0)
        m = max(2, min(self.krylovdim, n))
        v = V.copy(x0)
        nops, want = 0, min(nev, m)
        for _ in range(self.maxrestart + 1):
            V.scale(v, 1.0 / V.norm2(v))
            Q = [v]
            H = np.zeros((m + 1, m))
            k_eff = m
            for k in range(m):
                w = op(Q[k]); nops += 1
                for _pass in range(2):              # CGS2
                    c = [V.dot(q, w) for q in Q]
                    for ci, q in zip(c, Q):
                        V.axpby(w, -ci, q, 1.0)
                    H[:k + 1, k] += c
                H[k + 1, k] = V.norm2(w)
                if H[k + 1, k] < 1e-14 * max(1.0, np.abs(H[:k + 1, k]).max()):
                    k_eff = k + 1                   # invariant subspace
                    break
                Q.append(V.scale(w, 1.0 / H[k + 1, k]))
            vals, Y = _core.hessenberg_eig(H[:k_eff, :k_eff])
            order = np.argsort(-np.abs(vals), kind="stable")
            vals, Y = vals[order], Y[:, order]
            Y = Y / np.linalg.norm(Y, axis=0)
            want = min(nev, k_eff)
            beta = H[k_eff, k_eff - 1] if k_eff == m else 0.0
            res = np.abs(beta * Y[k_eff - 1, :want])
            done = bool(np.all(res <= self.tol * np.maximum(np.abs(vals[:want]), 1e-300)))
            if done or _ == self.maxrestart:
                vecs = [(self._lincomb(Q, Y[:, j].real), self._lincomb(Q, Y[:, j].imag)) for j in range(want)]
                return vals[:want], vecs, done, nops
            v = self._lincomb(Q, np.real(Y[:, :want].sum(axis=1)))  # restart with the sum of the wanted Ritz vectors
        raise AssertionError("unreachable")

    @staticmethod
    def _lincomb(Q, y):
        out = V.zeros_like(Q[0])
        for q, yi in zip(Q, y):
            if yi != 0.0:
                V.axpby(out, float(yi), q, 1.0)
        return out


class FloquetQaDB200:
    """``FloquetQaD(eigsolver)`` for ``Trapeze`` with a matrix-free vector field (Floquet.jl:46-85, 285-355).

    ctx_vf : context of the vector field (same grid/params as the orbit's); provides ``jacobian(u)`` and ``jvp(v, a0, a1)``.
    ls     : linear solver with the reference contract ``ls(J, rhs, a0=.., a1=..) -> (x, ok, its)``  (trap.linsolver).
    M      : number of time slices of the Trapeze discretisation; N = len(x) // M.
    """

    def __init__(self, ctx_vf, ls, M, eigsolver=None):
        self.ctx, self.ls, self.M = ctx_vf, ls, int(M)
        self.eigsolver = eigsolver or ArnoldiLMB200()
        self.solves = self.linear_its = 0
        self.all_converged = True

    def _factor(self, x, N, i_prev, i_cur, h, v):
        """v <- (I - h/2 J(x_cur))^-1 (I + h/2 J(x_prev)) v"""
        self.ctx.jacobian(time_slice(x, i_prev, N))
        rhs = self.ctx.jvp(v, a0=1.0, a1=h / 2)                      # out .+ h/2 .* apply(J_{i-1}, out)
        J = self.ctx.jacobian(time_slice(x, i_cur, N))
        res, ok, it = self.ls(J, rhs, a0=1.0, a1=-h / 2)              # trap.linsolver(J_i, out; a0 = 1, a1 = -h/2)
        self.solves += 1
        self.linear_its += it if np.isscalar(it) else sum(it)
        self.all_converged &= bool(ok)
        return res

    def _prepare(self, x):
        N = (len(x) - 1) // self.M
        if isinstance(x, DeviceVec) and x.ctx is not self.ctx:
            x.ctx.sync()  # the orbit was produced on another context's stream
        return N, period(x) / self.M  # uniform mesh: T * get_time_step(trap, i) = T / M

    def monodromy(self, x, du):
        """MonodromyQaD_matrix_free(trap, u0, par, du) (:285-316)"""
        N, h = self._prepare(x)
        M = self.M
        out = self._factor(x, N, M - 2, 0, h, du)                     # first factor uses slice M-1 (x_0 == x_{M-1})
        for ii in range(1, M - 1):
            out = self._factor(x, N, ii - 1, ii, h, out)
        return out

    def extract_eigenvector(self, x, zeta):
        """fl(Val(:ExtractEigenVector), wrap, u0, par, zeta) (:319-355): the M time slices of the eigenfunction."""
        N, h = self._prepare(x)
        M = self.M
        out = self._factor(x, N, M - 2, 0, h, zeta)
        res = [V.copy(out)]
        for ii in range(1, M):
            out = self._factor(x, N, ii - 1, ii, h, out)
            res.append(V.copy(out))
        return res

    def __call__(self, x, nev, x0=None):
        """compute_eigenvalues(fl, ...) (:59-85) -> (sigma, vecs, converged, info); sigma = log(mu) by decreasing real part."""
        N, _ = self._prepare(x)
        if x0 is None:
            rng = np.random.default_rng(0)
            r = rng.standard_normal(N)
            x0 = self.ctx.to_device(r) if isinstance(x, DeviceVec) else r  # Krylov vectors live in the vector field's context
        vals, vecs, cv, nops = self.eigsolver(lambda v: self.monodromy(x, v), x0, nev)
        logvals = np.log(vals.astype(complex))
        order = np.argsort(-logvals.real, kind="stable")
        return logvals[order], [vecs[i] for i in order], cv and self.all_converged, {"multipliers": vals[order], "monodromy_applications": nops,
                                                                                     "solves": self.solves, "linear_its": self.linear_its}


def cgl_shifted_precond(ctx_vf, T, M, r):
    """Preconditioner for the M-1 solves with I - h/2 J(x_i) on the cGL vector field: J = Lap + r + (rotation, nonlinear
    terms), so (1 - h/2 r) I - h/2 Lap is inverted exactly by the DST (BK_PC_CGL_DST).  Stand-in for the reference's
    per-slice factorisations (`jacobian_block_diag`, PeriodicOrbitTrapeze.jl:619-643)."""
    h = T / M
    ctx_vf.precond_setup(_l.BK_PC_CGL_DST, 1.0 - 0.5 * h * r, -0.5 * h)
