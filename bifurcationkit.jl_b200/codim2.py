"""Fold and Hopf points by the minimally augmented (MA) formulation, as host orchestration over the
same C ABI -- every linear solve is a (bordered) solve through `bk_bls_*` / `bk_gmres`, every operator application a `bk_jvp`.

Mirror of src/codim2/MinAugFold.jl:
  FoldMinAug.residual      <->  (F::FoldMinimallyAugmentedFormulation)(x, p, params)            :15-39
  FoldMinAug.bordered_terms <-> _compute_bordered_vectors / _get_bordered_terms                   :55-104
  FoldMinAug.solve         <->  foldMALinearSolver, finite-difference branch (usehessian = false) :122-146
  newton_fold              <->  newton_fold(prob, foldpointguess, par, eigenvec, eigenvec_ad, options; bdlinsolver)  :201-222
Swift-Hohenberg is self-adjoint (is_symmetric = true, examples/SH3d.jl:123), so J' = J and no adjoint kernel is needed (for
the Chan problem J' = J only up to the two boundary rows: the left null vector, hence sigma_x and sigma_p, are then approximate
and Newton on the MA system degrades to a quasi-Newton iteration that still converges to the same fold).

Mirror of src/codim2/MinAugHopf.jl:
  HopfMinAug.residual       <->  (H::HopfMinimallyAugmentedFormulation)(x, p, omega, params)      :19-40
  HopfMinAug.bordered_terms <->  __compute_bordered_vectors / _get_bordered_terms                 :59-104
  HopfMinAug.solve          <->  _hopf_MA_linear_solver, finite-difference branch                 :122-188
  newton_hopf               <->  newton_hopf(prob, hopfpointguess, par, eigenvec, eigenvec_ad, options)  :258-283
The complex shifts (J - i omega, (J - i omega)^H = J' + i omega) are solved on a BK_COMPLEX context (include/bk200.h): split
complex vectors, GMRES on the real-equivalent system, J' from bk_jac_set_transpose.  The complex bordered systems
[A a; b^H 0] are eliminated by bordering on the host (one complex solve each, since their right-hand side is (0, 1));
complex vectors are NumPy arrays (a Hopf refinement is a handful of solves, not a hot loop), real ones keep their container.
"""
from dataclasses import dataclass
import math

import numpy as np

from .palc import V


def _apply(J, v):
    """apply(J, v) (src/Utils.jl:192): J may be a callable Jacobian object or a matrix"""
    return J(v) if callable(J) else J @ v


@dataclass
class FoldSolution:
    u: object
    p: float
    residuals: list
    converged: bool
    itnewton: int
    itlinear: int
    sigma: float


class FoldMinAug:
    """[F(x, p); sigma(x, p)] with sigma from  [J a; b' 0] [v; sigma] = [0; 1]  (Govaerts 2000: a ~ left, b ~ right null vector)."""

    def __init__(self, prob, a, b, bls, symmetric=True):
        assert symmetric, "the adjoint Jacobian is not built: only self-adjoint problems (J' = J)"
        self.prob, self.a, self.b, self.bls = prob, V.copy(a), V.copy(b), bls
        self.zero = V.zeros_like(a)
        self.itlinear = 0

    def _border(self, J, a, b):
        """linbdsolver(J, a, b, 0, zero, 1) -> (v, sigma): J v + a sigma = 0, <b, v> = 1"""
        v, sig, cv, it = self.bls(J, a, b, 0.0, self.zero, 1.0)
        self.itlinear += int(np.sum(it))
        return v, sig

    def residual(self, x, p):
        J = self.prob.J(x, p)
        _, sigma = self._border(J, self.a, self.b)
        return self.prob.F(x, p), sigma

    def bordered_terms(self, x, p):
        prob = self.prob
        eps = prob.delta
        J = prob.J(x, p)
        v, _ = self._border(J, self.a, self.b)
        w, _ = self._border(J, self.b, self.a)          # adjoint system with J' = J
        # d_p F and sigma_p = -<w, d_p(J v)> by centred differences (MinAugFold.jl:92-101)
        dpF = prob.F(x, p + eps)
        V.axpby(dpF, -1.0 / (2 * eps), prob.F(x, p - eps), 1.0 / (2 * eps))
        jp = _apply(prob.J(x, p + eps), v)
        jm = _apply(prob.J(x, p - eps), v)
        V.axpby(jp, -1.0 / (2 * eps), jm, 1.0 / (2 * eps))
        sigma_p = -V.dot(w, jp)
        return v, w, dpF, sigma_p

    def solve(self, x, p, rhsu, rhsp):
        """foldMALinearSolver: [J d_pF; sigma_x' sigma_p] [dX; dp] = [rhsu; rhsp] with
        sigma_x = (J'(x) w - J'(x + eps v) w) / eps  (MinAugFold.jl:135-141)"""
        prob = self.prob
        eps = prob.delta
        v, w, dpF, sigma_p = self.bordered_terms(x, p)
        xs = V.copy(x)
        V.axpby(xs, eps, v, 1.0)
        u1 = _apply(prob.J(xs, p), w)
        J = prob.J(x, p)                                 # back to the linearisation at x (one state per context)
        sigma_x = _apply(J, w)
        V.axpby(sigma_x, -1.0 / eps, u1, 1.0 / eps)      # (u2 - u1) / eps
        dX, dp, cv, it = self.bls(J, dpF, sigma_x, sigma_p, rhsu, rhsp)
        self.itlinear += int(np.sum(it))
        return dX, dp, cv


def newton_fold(prob, x0, p0, eigenvec, eigenvec_ad, opts, bls, normN=V.norm2):
    """Newton on the MA system from the guess (x0, p0) with guesses for the right / left null vectors
    (newton_fold, MinAugFold.jl:201-222 + src/Newton.jl:66-114 on the bordered state)."""
    ma = FoldMinAug(prob, eigenvec_ad, eigenvec, bls)
    x, p = V.copy(x0), float(p0)
    F, sigma = ma.residual(x, p)
    res = math.hypot(normN(F), abs(sigma))
    residuals = [res]
    step = 0
    while step < opts.max_iterations and res > opts.tol:
        dX, dp, _ = ma.solve(x, p, F, sigma)
        V.axpby(x, -1.0, dX, 1.0)
        p -= dp
        F, sigma = ma.residual(x, p)
        res = math.hypot(normN(F), abs(sigma))
        residuals.append(res)
        step += 1
    return FoldSolution(x, p, residuals, residuals[-1] < opts.tol, step, ma.itlinear, sigma)


# ------------------------------------------------------------------------------------------------ Hopf
def _np(x):
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def _shifted(x, eps, d):
    """x + eps d for a real state x (NumPy array or DeviceVec) and a real NumPy direction d"""
    if hasattr(x, "ctx"):
        t = x.copy()
        t.axpby_(eps, x.ctx.to_device(d), 1.0)
        return t
    return x + eps * d


class ComplexProblemB200:
    """Complexified twin of a BifurcationProblemB200: the same stencil, grid and parameters on a BK_COMPLEX context.
    J(x, p, transpose) -> callable on complex vectors, consumable by ComplexGMRESB200."""

    def __init__(self, cctx, params, lens=0):
        assert cctx.complex
        self.ctx, self.params, self.lens = cctx, list(params), lens

    def J(self, x, p, transpose=False):
        q = list(self.params)
        q[self.lens] = p
        self.ctx.set_params(q)
        return self.ctx.cjacobian(x, transpose)


@dataclass
class HopfSolution:
    u: object
    p: float
    omega: float
    residuals: list
    converged: bool
    itnewton: int
    itlinear: int


class HopfMinAug:
    """[F(x, p); Re sigma; Im sigma](x, p, omega) with  [J - i omega, a; b^H, 0] [v; sigma] = [0; 1]
    (a ~ null vector of (J - i omega)^H, b ~ null vector of J - i omega)."""

    def __init__(self, prob, cprob, a, b, ls, cls):
        self.prob, self.cprob, self.ls, self.cls = prob, cprob, ls, cls
        self.a, self.b = np.array(a, dtype=complex), np.array(b, dtype=complex)
        self.itlinear = 0

    def _border(self, Jc, shift, a, b):
        """(Jc + shift) v + a sigma = 0, <b, v> = 1 by bordering: with y = (Jc + shift)^-1 a, sigma = -1 / <b, y>, v = -sigma y
        (linbdsolver(J, a, b, 0, zero, 1; shift), MinAugHopf.jl:17)"""
        y, cv, it = self.cls(Jc, a, a0=shift)
        self.itlinear += int(np.sum(it))
        sigma = -1.0 / np.vdot(b, y)
        return -sigma * y, sigma

    def residual(self, x, p, om):
        _, sigma = self._border(self.cprob.J(x, p), complex(0.0, -om), self.a, self.b)
        return self.prob.F(x, p), sigma.real, sigma.imag

    def bordered_terms(self, x, p, om):
        prob, cprob = self.prob, self.cprob
        eps = prob.delta
        v, _ = self._border(cprob.J(x, p), complex(0.0, -om), self.a, self.b)
        w, _ = self._border(cprob.J(x, p, transpose=True), complex(0.0, om), self.b, self.a)
        dpF = prob.F(x, p + eps)
        V.axpby(dpF, -1.0 / (2 * eps), prob.F(x, p - eps), 1.0 / (2 * eps))
        dpJv = (cprob.J(x, p + eps)(v) - cprob.J(x, p - eps)(v)) / (2 * eps)
        sigma_p = -np.vdot(w, dpJv)
        sigma_om = 1j * np.vdot(w, v)
        return v, w, dpF, sigma_p, sigma_om

    def solve(self, x, p, om, duu, dup, duom):
        """_hopf_MA_linear_solver: [J dpF 0; sigma_x sigma_p sigma_om] [dX; dp; dom] = [duu; dup; duom]"""
        prob, cprob = self.prob, self.cprob
        eps = prob.delta
        v, w, dpF, sigma_p, sigma_om = self.bordered_terms(x, p, om)
        x1, x2, cv, it = self.ls(prob.J(x, p), duu, dpF)
        self.itlinear += int(np.sum(it))
        cw = np.conj(w)
        u1r = cprob.J(_shifted(x, eps, np.ascontiguousarray(v.real)), p, transpose=True)(cw)
        u1i = cprob.J(_shifted(x, eps, np.ascontiguousarray(v.imag)), p, transpose=True)(cw)
        u2 = cprob.J(x, p, transpose=True)(cw)
        sigma_x = -(u1r - u2) / eps + 1j * (-(u1i - u2) / eps)
        sxx1 = np.vdot(sigma_x, _np(x1))
        sxx2 = np.vdot(sigma_x, _np(x2))
        # the inner product conjugates its first argument: hence + Im(sxx2) and + Im(sxx1) (MinAugHopf.jl:180-186)
        LS = np.array([[(sigma_p - sxx2).real, sigma_om.real], [(sigma_p + sxx2).imag, sigma_om.imag]])
        rhs = np.array([dup - sxx1.real, duom + sxx1.imag])
        dp, dom = np.linalg.solve(LS, rhs)
        V.axpby(x1, -dp, x2, 1.0)
        return x1, float(dp), float(dom), cv


def newton_hopf(prob, cprob, x0, p0, omega0, eigenvec, eigenvec_ad, opts, ls, cls, normN=V.norm2):
    """Newton on the Hopf MA system from (x0, p0, omega0) with guesses for the i omega eigenvector and its adjoint
    (newton_hopf, MinAugHopf.jl:258-283 + src/Newton.jl:66-114 on the state (x, [p, omega]))."""
    ma = HopfMinAug(prob, cprob, eigenvec_ad, eigenvec, ls, cls)
    x, p, om = V.copy(x0), float(p0), float(omega0)
    F, sr, si = ma.residual(x, p, om)
    res = math.sqrt(normN(F) ** 2 + sr * sr + si * si)
    residuals = [res]
    step = 0
    while step < opts.max_iterations and res > opts.tol:
        dX, dp, dom, _ = ma.solve(x, p, om, F, sr, si)
        V.axpby(x, -1.0, dX, 1.0)
        p -= dp
        om -= dom
        F, sr, si = ma.residual(x, p, om)
        res = math.sqrt(normN(F) ** 2 + sr * sr + si * si)
        residuals.append(res)
        step += 1
    return HopfSolution(x, p, om, residuals, residuals[-1] < opts.tol, step, ma.itlinear)
