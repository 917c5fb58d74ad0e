"""Bifurcation detection and location along a PALC branch -- host orchestration only.

What the reference does between two continuation steps when ``detect_bifurcation >= 2`` (src/Continuation.jl:506-560):
count unstable eigenvalues (``is_stable``, src/Bifurcations.jl:5-18), flag a change (``detect_bifurcation`` :21-28),
optionally locate it by bisection on the step size (``locate_bifurcation!`` :159-349, ``detect_bifurcation = 3``), classify
it from the change of (n_unstable, n_imag) (``get_bifurcation_type`` :70-150) and record a special point; folds by
parameter monotony when eigenvalues are not used (``locate_fold!`` :33-66).  The continuation step itself (``iterate``,
src/Continuation.jl:458-504) is assembled from the same pieces as ``palc.continuation`` -- corrector = ``newton_palc`` with
the context's bordered solver, tangent, predictor -- so every linear solve and eigen-solve still goes through the C ABI
(``MatrixFreeBLSB200`` / ``BorderingBLSB200`` / ``ShiftInvertB200``); ``palc.continuation`` is left untouched.

State vectors are ``DeviceVec`` or ndarray through the ``V`` interface of palc.py; the three state copies the bisection keeps
(`before`, `after`, current) are device copies (4 vectors each).
"""
import copy as _copy
from dataclasses import dataclass, field

import numpy as np

from .palc import (V, ContState, newton, newton_palc, step_size_control, _secant, _bordered_tangent, _predict)


# ------------------------------------------------------------------------------------------------ stability bookkeeping
def is_stable(contpar, eigvals):
    """src/Bifurcations.jl:5-18 -> (isstable, n_unstable, n_imag)"""
    if eigvals is None:
        return True, 0, 0
    ev = np.asarray(eigvals, dtype=complex)
    tol = contpar.tol_stability
    n_unstable = int(np.sum(ev.real > tol))
    n_imag = int(np.sum((np.abs(ev.imag) > tol) & (ev.real > tol)))
    return n_unstable == 0, n_unstable, n_imag


def detect_bifurcation(st):
    """src/Bifurcations.jl:21-28"""
    n1, n2 = st.n_unstable
    if n1 == -1 or n2 == -1:
        return False
    return n1 != n2


def detect_fold(p1, p2, p3):
    """src/Bifurcations.jl:31"""
    return (p3 - p2) * (p2 - p1) < 0


def rightmost(ev):
    """src/Utils.jl:31: eigenvalues sorted by |real part|"""
    ev = np.asarray(ev, dtype=complex)
    return ev[np.argsort(np.abs(ev.real), kind="stable")]


def getinterval(a, b):
    return (min(a, b), max(a, b))


@dataclass
class SpecialPoint:
    """src/Results.jl SpecialPoint (fields the detection fills)"""
    type: str
    idx: int            # 0-based row of the branch holding the state recorded with the point
    param: float
    norm: float
    step: int
    status: str         # :guess, :guessL, :converged
    delta: tuple        # (change of n_unstable, change of n_imag)
    ind_ev: int
    interval: tuple
    x: object = None
    tau_p: float = 0.0
    precision: float = -1.0


# ------------------------------------------------------------------------------------------------ state helpers
_VEC = ("z_u", "zold_u", "tau_u", "zpred_u")


def copy_state(st):
    """copy(state): src/Continuation.jl:196-212"""
    new = _copy.copy(st)
    for k in _VEC:
        setattr(new, k, V.copy(getattr(st, k)))
    return new


def copyto_state(dst, src):
    """copyto!(dest, src): src/Continuation.jl:214-240 (vectors copied into dst's own buffers)"""
    for k, v in vars(src).items():
        if k in _VEC:
            V.copyto(getattr(dst, k), v)
        else:
            setattr(dst, k, v)
    return dst


def _done(contpar, st):
    """src/Continuation.jl:254-257"""
    return (st.step <= contpar.max_steps) and ((contpar.p_min < st.z_p < contpar.p_max) or st.step == 0) and not st.stop


def _is_on_boundary(contpar, p):
    return p == contpar.p_min or p == contpar.p_max


class _Iter:
    """ContIterable: everything `iterate` needs (src/Continuation.jl:27-60)."""

    def __init__(self, prob, alg, contpar, normC):
        self.prob, self.alg, self.contpar, self.normC = prob, alg, contpar, normC

    # compute_eigenvalues! (src/Utils.jl:70-104) + update_stability! (src/Continuation.jl:274-278)
    def eigen(self, st):
        cp = self.contpar
        eig = cp.newton_options.eigsolver
        if cp.detect_bifurcation <= 0 or eig is None:
            return
        n = st.n_unstable[1]
        nev_ = max(n + 5, cp.nev)
        out = eig(self.prob.J(st.z_u, st.z_p), nev_)
        vals = np.asarray(out[0])
        _, nu, ni = is_stable(cp, vals)
        st.n_unstable = (nu, st.n_unstable[0])
        st.n_imag = (ni, st.n_imag[0])
        st.eigvals = vals
        st.eigvecs = out[1] if len(out) > 1 else None

    # iterate (src/Continuation.jl:458-504); returns False when the reference returns `nothing`
    def iterate(self, st):
        cp, alg, prob = self.contpar, self.alg, self.prob
        if not _done(cp, st):
            return False
        if st.zpred_p <= cp.p_min or st.zpred_p >= cp.p_max:  # Palc.jl:157-160 -> Natural corrector
            st.zpred_p = min(max(st.zpred_p, cp.p_min), cp.p_max)
            sol = newton(prob, st.zpred_u, st.zpred_p, cp.newton_options, self.normC)
            sol.p = st.zpred_p
        else:
            sol = newton_palc(prob, st.z_u, st.z_p, st.tau_u, st.tau_p, st.zpred_u, st.zpred_p, st.ds, alg.theta, cp, alg.bls,
                              self.normC)
        st.converged, st.itnewton, st.itlinear = sol.converged, sol.itnewton, sol.itlineartot
        st.work_newton += sol.itnewton
        st.work_linear += sol.itlineartot
        st.nfail += 0 if sol.converged else 1
        if sol.converged:
            st.zold_u, st.z_u = st.z_u, st.zold_u
            st.zold_p = st.z_p
            V.copyto(st.z_u, sol.u)
            st.z_p = sol.p
            self.eigen(st)
            st.step += 1
        if not st.stop and st.stepsizecontrol:            # step_size_control! (Contbase.jl:69-76)
            st.ds, st.stop = step_size_control(st.ds, st.converged, st.itnewton, cp)
        if st.converged:                                  # getpredictor! (Palc.jl:133-146)
            if alg.tangent == "secant":
                _secant(st, alg.theta)
            else:
                _bordered_tangent(prob, st, alg.theta, alg.bls)
        _predict(st)
        return True


# ------------------------------------------------------------------------------------------------ classification
def get_bifurcation_type(it, st, status, interval, floquet=False):
    """src/Bifurcations.jl:70-150 -> SpecialPoint (raises as the reference `throw`s when nothing changed)"""
    n_unstable, n_unstable_prev = st.n_unstable
    n_imag, n_imag_prev = st.n_imag
    ind_ev = n_unstable_prev if n_unstable < n_unstable_prev else n_unstable
    tp, known = "none", False
    dn, di = abs(n_unstable - n_unstable_prev), abs(n_imag - n_imag_prev)
    if dn == 1:
        tp = "bp" if di == 0 else (("pd" if floquet else "hopf") if di == 1 else "nd")
        known = True
    elif dn == 2:
        tp = ("ns" if floquet else "hopf") if di == 2 else "nd"
        known = True
    elif dn > 2:
        tp, known = "nd", True
    if dn < di:
        tp, known = "nd", True
    if st.n_unstable[0] * st.n_unstable[1] < 0 or st.n_imag[0] * st.n_imag[1] < 0:
        tp, known = "nd", True
    if not known:
        raise RuntimeError(f"We could not detect/identify the bifurcation point. (dn_unstable, dn_imag) = ({dn}, {di})")
    return SpecialPoint(type=tp, idx=st.step, param=st.z_p, norm=it.normC(st.z_u), step=st.step, status=status,
                        delta=(n_unstable - n_unstable_prev, n_imag - n_imag_prev), ind_ev=ind_ev, interval=tuple(interval),
                        x=V.copy(st.z_u), tau_p=st.tau_p, precision=abs(interval[1] - interval[0]))


def locate_fold(rows, specialpoints, it, st):
    """src/Bifurcations.jl:33-66 (called before the current state is saved: rows[-1] is the previous point)"""
    cp = it.contpar
    if cp.detect_fold and len(rows) > 2 and detect_fold(rows[-3]["param"], rows[-2]["param"], rows[-1]["param"]):
        specialpoints.append(SpecialPoint(type="fold", idx=len(rows) - 2, param=st.z_p, norm=it.normC(st.z_u), step=len(rows) - 2,
                                          status="guess", delta=(0, 0), ind_ev=0, interval=(rows[-2]["param"], rows[-2]["param"]),
                                          x=V.copy(st.z_u), tau_p=st.tau_p))
        return True
    return False


# ------------------------------------------------------------------------------------------------ bisection
def locate_bifurcation(it, _st):
    """locate_bifurcation!(iter, state) (src/Bifurcations.jl:159-349): bisection on ds; on return `_st` sits just after
    the bifurcation point (or is restored to `after`), status in {guess, guessL, converged, none}."""
    assert detect_bifurcation(_st), "No bifurcation detected for the state"
    cp = it.contpar
    n2, n1 = _st.n_unstable
    if n1 == -1 or n2 == -1 or abs(_st.ds) < cp.dsmin:
        return "none", (0.0, 0.0)
    after, st, before = copy_state(_st), copy_state(_st), copy_state(_st)
    st.in_bisection = True
    before.n_unstable = (before.n_unstable[1], before.n_unstable[0])
    before.n_imag = (before.n_imag[1], before.n_imag[0])
    before.zold_p, before.z_p = before.z_p, before.zold_p
    st.ds *= -1
    st.step = 0
    st.stepsizecontrol = False
    alive = True                       # `next !== nothing`
    nunstbls, nimags = [n2], [st.n_imag[0]]
    interval = list(getinterval(st.z_p, st.zold_p))
    indinterval = 0 if interval[0] == st.z_p else 1
    n_inversion = 0
    while True:
        if not st.converged:
            break                      # Newton failed to fully locate the point with the bisection parameters
        if not alive:
            break
        nunstbls.append(st.n_unstable[0])
        nimags.append(st.n_imag[0])
        if nunstbls[-1] == nunstbls[-2]:
            st.ds /= 2                 # bifurcation point still after the current state, keep going
        else:
            st.ds /= -2                # passed it: reverse
            n_inversion += 1
            indinterval = 1 - indinterval
        _predict(st)                   # update_predictor!
        copyto_state(after if n_inversion % 2 == 0 else before, st)
        if st.step > 0:
            interval[indinterval] = st.z_p
        ev = rightmost(st.eigvals)
        biflocated = abs(ev.real[0]) < cp.tol_bisection_eigenvalue
        if not (abs(st.ds) >= cp.dsmin_bisection and st.step < cp.max_bisection_steps and n_inversion < cp.n_inversion
                and not biflocated):
            break
        alive = it.iterate(st)
    if n_inversion % 2 == 0:
        status = "converged" if n_inversion >= cp.n_inversion else "guess"
        src = st
        _st.n_unstable = (st.n_unstable[0], before.n_unstable[0])
        _st.n_imag = (st.n_imag[0], before.n_imag[0])
        interval = (st.z_p, before.z_p)
    else:
        status = "guessL"
        src = after
        _st.n_unstable = (after.n_unstable[0], st.n_unstable[0])
        _st.n_imag = (after.n_imag[0], st.n_imag[0])
        interval = (st.z_p, after.z_p)
    for k in _VEC:
        V.copyto(getattr(_st, k), getattr(src, k))
    _st.z_p, _st.zold_p, _st.tau_p, _st.zpred_p = src.z_p, src.zold_p, src.tau_p, src.zpred_p
    _st.eigvals, _st.eigvecs = src.eigvals, getattr(src, "eigvecs", None)
    _st.work_newton, _st.work_linear = st.work_newton, st.work_linear   # the bisection's corrector work is real work
    _predict(_st)                      # update_predictor!(_state, iter) with the outer ds
    return status, getinterval(*interval)


# ------------------------------------------------------------------------------------------------ driver
@dataclass
class Branch:
    rows: list = field(default_factory=list)
    specialpoint: list = field(default_factory=list)
    eig: list = field(default_factory=list)
    state: object = None


def continuation(prob, alg, contpar, normC=V.norm2, verbose=False, callback=None, floquet=False):
    """continuation(prob, PALC(...), ContinuationPar(detect_bifurcation = 0..3)) with special points
    (src/Continuation.jl:349-400 start-up, :506-575 loop).  Returns a Branch (rows as palc.continuation + `stable`,
    `n_imag`; specialpoint list ends with the :endpoint)."""
    cp, opts = contpar, contpar.newton_options
    it = _Iter(prob, alg, cp, normC)
    p0 = prob.p0
    assert cp.p_min <= p0 <= cp.p_max
    sol0 = newton(prob, prob.u0, p0, opts, normC)
    if not sol0.converged:
        raise RuntimeError(f"Newton failed to converge for the initial guess: {sol0.residuals}")
    p1 = p0 + cp.ds / cp.eta
    sol1 = newton(prob, sol0.u, p1, opts, normC)
    if not sol1.converged:
        raise RuntimeError("Newton failed to converge for the initial tangent")
    u0, u1 = sol0.u, sol1.u
    st = ContState(z_u=u1, z_p=p1, zold_u=u0, zold_p=p0, tau_u=V.zeros_like(u0), tau_p=0.0, zpred_u=V.zeros_like(u0),
                   zpred_p=0.0, ds=cp.ds)
    st.eigvecs = None
    _secant(st, alg.theta)
    st.z_u, st.z_p = V.copy(u0), p0
    _predict(st)
    br = Branch(state=st)

    def save():
        stable, _, _ = is_stable(cp, st.eigvals)
        br.rows.append(dict(param=st.z_p, x=prob.record(st.z_u), itnewton=st.itnewton, itlinear=st.itlinear, ds=st.ds,
                            step=st.step, n_unstable=st.n_unstable[0], n_imag=st.n_imag[0], stable=stable))
        if st.eigvals is not None:
            br.eig.append(dict(eigenvals=np.array(st.eigvals), step=st.step))

    it.eigen(st)
    save()
    if callback is not None and callback(st) is False:
        st.stop = True
    status = "guess"
    alive = True
    first = True
    while alive:
        if not first and st.converged and st.step <= cp.max_steps and st.step > 0:
            if cp.detect_fold and cp.detect_bifurcation < 2:
                locate_fold(br.rows, br.specialpoint, it, st)
            if cp.detect_bifurcation > 1 and detect_bifurcation(st):
                interval = getinterval(st.zold_p, st.z_p)
                if cp.detect_bifurcation > 2 and not _is_on_boundary(cp, st.z_p):
                    status, interval = locate_bifurcation(it, st)
                if detect_bifurcation(st):   # the bisection may have moved the state before the point
                    bp = get_bifurcation_type(it, st, status, interval, floquet)
                    if bp.type != "none":
                        br.specialpoint.append(bp)
                    if verbose:
                        print(f"--> {bp.type} bifurcation point at p ~ {bp.param:.8g} in {bp.interval}, delta = {bp.delta}, {bp.status}", flush=True)
            save()
            if callback is not None and callback(st) is False:
                st.stop = True
        first = False
        alive = it.iterate(st)
        if verbose and alive:
            print(f"step {st.step} p={st.z_p:.6e} ds={st.ds:.3e} conv={st.converged} itn={st.itnewton} n_unstable={st.n_unstable}", flush=True)
    br.specialpoint.append(SpecialPoint(type="endpoint", idx=len(br.rows) - 1, param=st.z_p, norm=normC(st.z_u), step=st.step,
                                        status="converged", delta=(0, 0), ind_ev=0, interval=(st.z_p, st.z_p)))
    return br
