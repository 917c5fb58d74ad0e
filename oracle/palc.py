"""Newton, newton_palc and the PALC continuation driver (NumPy restatement).
Test infrastructure only (see oracle/__init__.py).

Follows src/Newton.jl:66-114, src/continuation/Palc.jl:1-56,112-305,
src/continuation/Tangents.jl:8-42,71-104, src/continuation/Contbase.jl:69-102,
src/continuation/Natural.jl:36-58, src/Continuation.jl:254-257,349-504,506-601.
"""
from dataclasses import dataclass, field
import numpy as np

from .problems import SQRT_EPS


def norminf(x):
    return float(np.max(np.abs(x))) if len(x) else 0.0


def norm2(x):
    return float(np.linalg.norm(x))


@dataclass
class NewtonPar:
    """src/Newton.jl:17-33"""
    tol: float = 1e-10
    max_iterations: int = 25
    linsolver: object = None
    eigsolver: object = None


@dataclass
class ContinuationPar:
    """src/ContParameters.jl:44-100 (the fields the hot path reads)."""
    dsmin: float = 1e-4
    dsmax: float = 1e-1
    ds: float = 1e-2
    a: float = 0.5
    p_min: float = -1.0
    p_max: float = 1.0
    max_steps: int = 400
    newton_options: NewtonPar = field(default_factory=NewtonPar)
    eta: float = 150.0
    nev: int = 3
    detect_bifurcation: int = 0
    tol_stability: float = 1e-10


@dataclass
class Problem:
    """F(x, p) -> residual, J(x, p) -> matrix or callable dx -> J dx (src/Problems.jl:98-101)."""
    F: object
    J: object
    u0: np.ndarray
    p0: float
    delta: float = SQRT_EPS
    record: object = norm2  # record_from_solution default = norm(x) (src/Problems.jl:286)


@dataclass
class PALC:
    """src/continuation/Palc.jl:70-84"""
    tangent: str = "secant"  # or "bordered"
    theta: float = 0.5
    bls: object = None


@dataclass
class NonLinearSolution:
    u: object
    p: float
    residuals: list
    converged: bool
    itnewton: int
    itlineartot: int


def newton(prob, x0, p, opts, normN=norm2):
    """src/Newton.jl:66-114 (_newton)."""
    x = x0.copy()
    fx = prob.F(x, p)
    res = normN(fx)
    residuals = [res]
    step = 0
    itlin = 0
    while step < opts.max_iterations and res > opts.tol:
        J = prob.J(x, p)
        u, cv, it = opts.linsolver(J, fx)
        itlin += int(np.sum(it))
        x = x - u
        fx = prob.F(x, p)
        res = normN(fx)
        residuals.append(res)
        step += 1
    return NonLinearSolution(x, p, residuals, residuals[-1] < opts.tol, step, itlin)


def dot_theta(u1, u2, p1, p2, theta):
    """src/continuation/Palc.jl:1-41: theta*dot(u1,u2)/N + (1-theta) p1 p2."""
    return float(np.dot(u1, u2)) / len(u1) * theta + p1 * p2 * (1.0 - theta)


def norm_theta(u, p, theta):
    return np.sqrt(dot_theta(u, u, p, p, theta))


def arc_length_eq(u1, u2, dp, tau_u, tau_p, theta, ds):
    """src/continuation/Palc.jl:44-56 (u1 = x, u2 = z0.u, dp = p - z0.p)."""
    return (dot_theta(u1, tau_u, dp, tau_p, theta) - ds) - (dot_theta(u2, tau_u, dp, 0.0, theta) - 0.0)


def solve_bls_palc(bls, theta, tau_u, tau_p, J, dR, R, n):
    """src/LinearBorderSolver.jl:16-36: xiu=theta, xip=1-theta, dotp=dot/N."""
    N = len(R)
    dotp = lambda a, b: float(np.dot(a, b)) / N
    return bls(J, dR, tau_u, tau_p, R, n, theta, 1.0 - theta, shift=None, dotp=dotp,
               apply_xiu=lambda row: row / N)


def newton_palc(prob, z0u, z0p, tau_u, tau_p, zpred_u, zpred_p, ds, theta, contpar, bls, normN=norm2):
    """src/continuation/Palc.jl:187-305 (no line search)."""
    opts = contpar.newton_options
    eps = prob.delta
    Nfun = lambda u, p: arc_length_eq(u, z0u, p - z0p, tau_u, tau_p, theta, ds)
    x = zpred_u.copy()
    p = zpred_p
    res_f = prob.F(x, p)
    res_n = Nfun(x, p)
    res = max(normN(res_f), abs(res_n))
    residuals = [res]
    step = 0
    itlin = 0
    while step < opts.max_iterations and res > opts.tol:
        dFdp = (prob.F(x, p + eps) - res_f) / eps
        J = prob.J(x, p)
        u, up, flag, it = solve_bls_palc(bls, theta, tau_u, tau_p, J, dFdp, res_f, res_n)
        itlin += int(np.sum(it))
        x = x - u
        p = min(max(p - up, contpar.p_min), contpar.p_max)
        res_f = prob.F(x, p)
        res_n = Nfun(x, p)
        res = max(normN(res_f), abs(res_n))
        residuals.append(res)
        step += 1
    return NonLinearSolution(x, p, residuals, residuals[-1] < opts.tol, step, itlin)


def step_size_control(ds, converged, itnewton, contpar):
    """src/continuation/Contbase.jl:77-102 -> (ds_new, stop)."""
    if not converged:
        if abs(ds) <= contpar.dsmin:
            return ds, True
        dsnew = np.sign(ds) * max(abs(ds) / 2, contpar.dsmin)
    else:
        Nmax = contpar.newton_options.max_iterations
        factor = (Nmax - itnewton) / Nmax
        dsnew = ds * (1 + contpar.a * (factor * factor))  # factor^2 is a literal power in Julia: x * x
    dsnew = np.sign(dsnew) * min(max(abs(dsnew), contpar.dsmin), contpar.dsmax)  # clamp_ds ContParameters.jl:107
    return float(dsnew), False


@dataclass
class ContState:
    z_u: np.ndarray
    z_p: float
    zold_u: np.ndarray
    zold_p: float
    tau_u: np.ndarray
    tau_p: float
    zpred_u: np.ndarray
    zpred_p: float
    ds: float
    step: int = 0
    converged: bool = True
    itnewton: int = 0
    itlinear: int = 0
    stop: bool = False
    n_unstable: tuple = (-1, -1)
    eigvals: object = None


def secant_tangent(z1u, z1p, z0u, z0p, ds, theta):
    """src/continuation/Tangents.jl:28-42"""
    tu = z1u - z0u
    tp = z1p - z0p
    alpha = np.sign(ds) / norm_theta(tu, tp, theta)
    return tu * alpha, tp * alpha


def bordered_tangent(prob, st, theta, bls):
    """src/continuation/Tangents.jl:71-104"""
    eps = prob.delta
    dFdl = (prob.F(st.z_u, st.z_p + eps) - prob.F(st.z_u, st.z_p)) / eps
    J = prob.J(st.z_u, st.z_p)
    tu, tp, flag, it = solve_bls_palc(bls, theta, st.tau_u, st.tau_p, J, dFdl, np.zeros_like(st.z_u), 1.0)
    alpha = 1.0 / np.sqrt(dot_theta(tu, tu, tp, tp, theta))
    alpha *= np.sign(dot_theta(st.tau_u, tu, st.tau_p, tp, theta))
    return tu * alpha, tp * alpha


def continuation(prob, alg, contpar, normC=norm2, u1=None, p1=None, verbose=False, callback=None):
    """src/Continuation.jl:349-504,506-601.  Returns (rows, state) where rows is a list of dicts
    (param, x=record, itnewton, itlinear, ds, step, n_unstable) like ContResult.branch
    (src/Continuation.jl:259-272).  If (u1, p1) are given, starts from two points
    (iterate_from_two_points, src/Continuation.jl:408-456) with prob.u0/p0 as the first."""
    opts = contpar.newton_options
    theta = alg.theta
    bls = alg.bls
    p0 = prob.p0
    if u1 is None:
        assert contpar.p_min <= p0 <= contpar.p_max
        sol0 = newton(prob, prob.u0, p0, opts, normC)
        if not sol0.converged:
            raise RuntimeError("Newton failed to converge for the initial guess")
        p1 = p0 + contpar.ds / contpar.eta
        sol1 = newton(prob, sol0.u, p1, opts, normC)
        if not sol1.converged:
            raise RuntimeError("Newton failed to converge for the initial tangent")
        u0, u1 = sol0.u, sol1.u
    else:
        u0 = prob.u0.copy()
    # iterate_from_two_points: state.z = z1, z_old = z0; initialize! -> secant tangent, z <- z0, z_pred
    tau_u, tau_p = secant_tangent(u1, p1, u0, p0, contpar.ds, theta)
    st = ContState(z_u=u0.copy(), z_p=p0, zold_u=u0.copy(), zold_p=p0, tau_u=tau_u, tau_p=tau_p,
                   zpred_u=u0 + contpar.ds * tau_u, zpred_p=p0 + contpar.ds * tau_p, ds=contpar.ds)
    rows = []

    def eig_update():
        if contpar.detect_bifurcation > 0 and opts.eigsolver is not None:
            nprev = st.n_unstable[0]
            nev_ = max(nprev + 5, contpar.nev) if nprev >= 0 else contpar.nev
            J = prob.J(st.z_u, st.z_p)
            vals = opts.eigsolver(J, nev_)[0]
            nun = int(np.sum(np.real(vals) > contpar.tol_stability))
            st.n_unstable = (nun, st.n_unstable[0])
            st.eigvals = vals

    def save():
        rows.append(dict(param=st.z_p, x=prob.record(st.z_u), itnewton=st.itnewton, itlinear=st.itlinear,
                         ds=st.ds, step=st.step, n_unstable=st.n_unstable[0]))

    eig_update()
    save()  # ContResult(it, state) at step 0 (src/Continuation.jl:322-330)
    if callback is not None and callback(st) is False:  # step 0 hook (marks the start of the continuation! loop)
        st.stop = True

    def done():
        return (st.step <= contpar.max_steps) and ((contpar.p_min < st.z_p < contpar.p_max) or st.step == 0) and not st.stop

    first = True
    while True:
        if not first and st.converged and st.step <= contpar.max_steps and st.step > 0:
            save()
            if callback is not None and callback(st) is False:
                st.stop = True
        first = False
        if not done():
            break
        # corrector! (src/continuation/Palc.jl:153-176)
        if st.zpred_p <= contpar.p_min or st.zpred_p >= contpar.p_max:
            st.zpred_p = min(max(st.zpred_p, contpar.p_min), contpar.p_max)
            sol = newton(prob, st.zpred_u, st.zpred_p, opts, normC)  # Natural corrector
            sol.p = st.zpred_p
        else:
            sol = newton_palc(prob, st.z_u, st.z_p, st.tau_u, st.tau_p, st.zpred_u, st.zpred_p, st.ds,
                              theta, contpar, bls, normC)
        st.converged, st.itnewton, st.itlinear = sol.converged, sol.itnewton, sol.itlineartot
        if sol.converged:
            st.zold_u, st.zold_p = st.z_u, st.z_p
            st.z_u, st.z_p = sol.u.copy(), sol.p
            eig_update()
            st.step += 1
        if verbose:
            print(f"step {st.step} p={st.z_p:.6e} ds={st.ds:.3e} conv={st.converged} itn={st.itnewton} itl={st.itlinear}")
        # step size control
        if not st.stop:
            st.ds, stop = step_size_control(st.ds, st.converged, st.itnewton, contpar)
            st.stop = stop
        # predictor
        if st.converged:
            if alg.tangent == "secant":
                st.tau_u, st.tau_p = secant_tangent(st.z_u, st.z_p, st.zold_u, st.zold_p, st.ds, theta)
            else:
                st.tau_u, st.tau_p = bordered_tangent(prob, st, theta, bls)
        st.zpred_u = st.z_u + st.ds * st.tau_u
        st.zpred_p = st.z_p + st.ds * st.tau_p
    return rows, st
