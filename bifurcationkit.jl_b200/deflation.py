"""Deflated Newton -- host orchestration over the same C ABI.

* ``DeflationOperator``       <-> src/DeflationOperator.jl:50-170:  M(u) = prod_i (1 / <u - r_i, u - r_i>^p + alpha)
  (or the mean, ``accumulator="mean"``); ``dM(u, du)`` by the reference's forward difference (delta = 1e-8, :158-166).
* ``DeflatedProblem``         <-> :172-232:  residual M(u) F(u); its Jacobian handle is the triple (u, p, problem) (:215-217).
* ``DeflatedProblemCustomLS`` <-> :247-310:  Sherman-Morrison-type solve of  M(u) J h + F(u) dM(u).h = rhs  with the *two-rhs*
  call ``linsolve(Ju, rhs, Fu)`` (src/LinearSolver.jl:15-19 -> ``GMRESB200(J, rhs, rhs2)``), h = (h1 - z h2) / M(u),
  z = dM.h1 / (M(u) + dM.h2).
* ``newton_deflated``         <-> ``solve(prob, defOp, options)`` (:340-356);  ``newton_two_guesses`` <-> ``newton(prob, x0, x1, p0, options)``
  (:392-420), the variant the reference uses for branch switching seeds (examples/SH2d-fronts.jl:70-80 builds the localized
  fronts this way).

Vectors are ``DeviceVec`` or ndarray through ``palc.V``; one extra device vector of scratch, like the reference's ``tmp``.
"""
import numpy as np

from .palc import V, NewtonPar, newton
from dataclasses import replace as _replace


class DeflationOperator:
    def __init__(self, power, alpha, roots, dot=None, delta=1e-8, accumulator="prod"):
        assert accumulator in ("prod", "mean")
        self.power, self.alpha, self.roots, self.delta, self.accumulator = power, float(alpha), list(roots), delta, accumulator
        self.dot = dot or V.dot

    def __len__(self):
        return len(self.roots)

    def __getitem__(self, i):
        return self.roots[i]

    def push(self, r):
        self.roots.append(r)

    def pop(self):
        return self.roots.pop()

    def __call__(self, u, tmp=None):
        """M(u) (:118-131)"""
        if not self.roots:
            return 1.0
        tmp = V.copy(u) if tmp is None else tmp
        out = None
        for r in self.roots:
            V.copyto(tmp, u)
            V.axpby(tmp, -1.0, r, 1.0)
            m = 1.0 / self.dot(tmp, tmp) ** self.power + self.alpha
            out = m if out is None else (out * m if self.accumulator == "prod" else out + m)
        return out / len(self.roots) if self.accumulator == "mean" else out

    def dM(self, u, du, tmp=None, tmp2=None):
        """dM(u).du by forward difference (:158-166)"""
        if not self.roots:
            return 0.0
        tmp = V.copy(u) if tmp is None else V.copyto(tmp, u)
        V.axpby(tmp, self.delta, du, 1.0)
        return (self(tmp, tmp2) - self(u, tmp2)) / self.delta


class DeflatedProblem:
    """M(u) F(u) = 0 (:172-232), same duck type as BifurcationProblemB200 for palc.newton."""

    def __init__(self, prob, M):
        self.prob, self.M = prob, M
        self.u0, self.p0 = prob.u0, prob.p0
        self.delta = getattr(prob, "delta", 1e-8)
        self.record = getattr(prob, "record", None)

    def F(self, x, p, out=None):
        out = self.prob.F(x, p, out)
        return V.scale(out, self.M(x))

    def J(self, x, p):
        return (x, p, self)   # jacobian(dfp::DeflatedProblem{..., Val{:Custom}}, x, p)

    def jvp(self, x, p, du):
        """dF(u).du M(u) + F(u) dM(u).du (:193-207)"""
        J = self.prob.J(x, p)
        out = V.scale(J(du), self.M(x))
        if len(self.M):
            V.axpby(out, self.M.dM(x, du), self.prob.F(x, p), 1.0)
        return out


class DeflatedProblemCustomLS:
    """(:247-310)"""

    def __init__(self, solver):
        self.solver = solver

    def __call__(self, J, rhs):
        u, p, dp = J
        Fu = dp.prob.F(u, p)
        Mu = dp.M(u)
        Ju = dp.prob.J(u, p)
        if len(dp.M) == 0:
            h1, ok, it1 = self.solver(Ju, rhs)
            return h1, ok, (it1, 0)
        h1, h2, ok, its = self.solver(Ju, rhs, Fu)   # two right-hand sides
        tmp, tmp2 = V.zeros_like(u), V.zeros_like(u)
        z1 = dp.M.dM(u, h1, tmp, tmp2)
        z2 = dp.M.dM(u, h2, tmp, tmp2)
        z = z1 / (Mu + z2)
        V.copyto(tmp, h1)
        V.axpby(tmp, -z, h2, 1.0)
        V.scale(tmp, 1.0 / Mu)
        return tmp, True, its


def newton_deflated(prob, x0, p, defop, opts, normN=V.norm2):
    """solve(prob, defOp, options) (:340-356): Newton on M(u) F(u) with the custom linear solver around opts.linsolver."""
    dprob = DeflatedProblem(prob, defop)
    return newton(dprob, x0, p, _replace(opts, linsolver=DeflatedProblemCustomLS(opts.linsolver)), normN)


def newton_two_guesses(prob, x0, x1, p, opts, defop=None, normN=V.norm2):
    """newton(prob, x0, x1, p0, options, defOp) (:392-420) -> (sol1, sol0, ok): converge from x0, deflate that root, then
    converge from x1 to a different one."""
    defop = defop or DeflationOperator(2, 1.0, [])
    sol0 = newton(prob, x0, p, opts, normN)
    assert sol0.converged, "Newton did not converge to the trivial solution x0."
    defop.push(sol0.u)
    sol1 = newton_deflated(prob, x1, p, defop, _replace(opts, max_iterations=10 * opts.max_iterations), normN)  # (:401)
    return sol1, sol0, sol0.converged and sol1.converged
