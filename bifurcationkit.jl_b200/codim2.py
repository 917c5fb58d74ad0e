"""SURVEY section 8f.3 (first half): Fold points by the minimally augmented (MA) formulation, as host orchestration over the
same C ABI -- every linear solve is a bordered solve through `bk_bls_*`, every operator application a `bk_jvp`.

Mirror of src/codim2/MinAugFold.jl:
  FoldMinAug.residual      <->  (F::FoldMinimallyAugmentedFormulation)(x, p, params)            :15-39
  FoldMinAug.bordered_terms <-> _compute_bordered_vectors / _get_bordered_terms                   :55-104
  FoldMinAug.solve         <->  foldMALinearSolver, finite-difference branch (usehessian = false) :122-146
  newton_fold              <->  newton_fold(prob, foldpointguess, par, eigenvec, eigenvec_ad, options; bdlinsolver)  :201-222
Swift-Hohenberg is self-adjoint (is_symmetric = true, examples/SH3d.jl:123), so J' = J and no adjoint kernel is needed (for
the Chan problem J' = J only up to the two boundary rows: the left null vector, hence sigma_x and sigma_p, are then approximate
and Newton on the MA system degrades to a quasi-Newton iteration that still converges to the same fold).  The Hopf MA formulation (MinAugHopf.jl) needs complex shifts a0 = i omega, i.e. a complex
bk_gmres: not built.
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
