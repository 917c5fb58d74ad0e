"""CPU tests of the codim-2 host logic (bifurcationkit.jl_b200/codim2.py: continuation_fold / continuation_hopf) on
host arrays with the oracle's dense solvers as the backend -- pinned to the reference's own golden values:
test/fold_codim_2/codim2.jl (CO-oxidation model): the Fold point `sn.u.u`, `sn.u.p` (:66-69), the Bogdanov-Takens point on
the Fold curve at k = 0.9716038596420551 (:88-90), and the null-vector checks of the updated border vectors (:93-106)."""
import numpy as np
import pytest

import __graft_entry__ as g
from oracle import krylov, bls as obls
from tests.test_host_logic_cpu import BlsAdapter


class NumpyProblem2:
    """Duck-typed BifurcationProblemB200 on host arrays with a parameter tuple and a lens (index): F(x, params), J(x, params)."""

    def __init__(self, F, J, u0, params, lens, record=None):
        self.F_, self.J_, self.u0, self.params, self.lens = F, J, u0, list(params), lens
        self.p0 = float(params[lens])
        self.delta = float(np.sqrt(np.finfo(float).eps))
        self.record = record or (lambda x: float(np.linalg.norm(x)))

    def _par(self, p):
        q = list(self.params)
        q[self.lens] = p
        return q

    def F(self, x, p, out=None):
        r = self.F_(x, self._par(p))
        if out is not None:
            out[...] = r
            return out
        return r

    def J(self, x, p):
        return self.J_(x, self._par(p))

    def Jt(self, x, p):  # jacobian_adjoint
        return self.J_(x, self._par(p)).T


def COm(u, q):
    """test/fold_codim_2/codim2.jl:8-17"""
    q1, q2, q3, q4, q5, q6, k = q
    x, y, s = u
    z = 1 - x - y - s
    return np.array([2 * q1 * z**2 - 2 * q5 * x**2 - q3 * x * y, q2 * z - q6 * y - q3 * x * y, q4 * z - k * q4 * s])


def COmJ(u, q):
    q1, q2, q3, q4, q5, q6, k = q
    x, y, s = u
    z = 1 - x - y - s
    return np.array([[-4 * q1 * z - 4 * q5 * x - q3 * y, -4 * q1 * z - q3 * x, -4 * q1 * z],
                     [-q2 - q3 * y, -q2 - q6 - q3 * x, -q2],
                     [-q4, -q4, -q4 - k * q4]])


PAR_COM = [2.5, 2.0, 10.0, 0.0675, 1.0, 0.1, 0.4]          # (q1, q2, q3, q4, q5, q6, k), codim2.jl:20
SN_U = np.array([0.05402941507127516, 0.3022414400400177, 0.45980653206336225])  # codim2.jl:67
SN_P = 1.0522002878699546                                   # codim2.jl:68
BT_K = 0.9716038596420551                                   # codim2.jl:90


@pytest.fixture(scope="module")
def com_fold():
    """the branch of codim2.jl:22-30 in q2 (downwards from q2 = 2), its second turning point refined by newton_fold"""
    bk = g.load_package()
    P, C2 = bk.palc, bk.codim2
    ls, bls = krylov.DefaultLS(), BlsAdapter(obls.MatrixBLS())
    prob = NumpyProblem2(COm, COmJ, np.array([0.07, 0.2, 5.0]), PAR_COM, 1)
    cp = P.ContinuationPar(p_min=0.6, p_max=2.5, ds=-0.002, dsmax=0.01, dsmin=1e-4, max_steps=3000,
                           newton_options=P.NewtonPar(tol=1e-10, max_iterations=25, linsolver=ls))
    pts = []
    rows, _ = P.continuation(prob, P.PALC(bls=bls), cp, normC=P.norminf,
                             callback=lambda s: pts.append((s.z_u.copy(), s.z_p, s.tau_u.copy())) or True)
    ps = [r["param"] for r in rows]
    turns = [i for i in range(1, len(ps) - 1) if (ps[i] - ps[i - 1]) * (ps[i + 1] - ps[i]) < 0]
    assert len(turns) == 2 and rows[-1]["param"] == 0.6   # the S-shaped branch of the CO model: two folds
    x0, p0, tau = pts[turns[1]]
    t = tau / np.linalg.norm(tau)
    sol = C2.newton_fold(prob, x0, p0, t, t, P.NewtonPar(tol=1e-10, max_iterations=10, linsolver=ls), bls, symmetric=False)
    return bk, prob, ls, bls, sol, t


def test_fold_point_of_the_co_model_matches_the_reference(com_fold):
    """codim2.jl:64-69: sn = newton(br, 3; bdlinsolver = MatrixBLS()) -> sn.u.u, sn.u.p (rtol 1e-4 there)"""
    _, prob, _, _, sol, _ = com_fold
    assert sol.converged and sol.itnewton <= 4
    assert np.allclose(sol.u, SN_U, rtol=1e-9) and abs(sol.p - SN_P) < 1e-10
    J = COmJ(sol.u, prob._par(sol.p))
    assert np.min(np.abs(np.linalg.eigvals(J))) < 1e-9 and np.linalg.norm(COm(sol.u, prob._par(sol.p))) < 1e-12


def test_fold_curve_of_the_co_model(com_fold):
    """codim2.jl:75-106: continuation of the Fold in (q2, k), k from 0.4 to 1 in at most 50 steps: the Bogdanov-Takens point (zero
    of the test function <a, b> of the updated border vectors; reported at k = 0.97160386 by the reference), a singular Jacobian
    at the last point and a, b equal to the null vectors of J', J there."""
    bk, prob, ls, bls, sol, t = com_fold
    P, C2 = bk.palc, bk.codim2
    cpf = P.ContinuationPar(p_min=0.0, p_max=1.0, ds=0.002, dsmax=0.01, dsmin=1e-4, max_steps=50,
                            newton_options=P.NewtonPar(tol=1e-10, max_iterations=25, linsolver=ls))
    curve = C2.continuation_fold(prob, sol.u, sol.p, 6, t, t, cpf, bls, symmetric=False, normC=P.norminf)
    assert prob.params == PAR_COM                       # the caller's parameters are restored
    assert curve.p2[0] == 0.4 and abs(curve.p1[0] - SN_P) < 1e-10 and curve.p2[-1] == 1.0 and len(curve.rows) <= 51
    assert max(r["itnewton"] for r in curve.rows) <= 3  # Newton on the MA system converges quadratically from the predictor
    # every point of the curve is a Fold point: F = 0 and a singular Jacobian
    z = curve.state.z_u
    par = list(PAR_COM)
    par[1], par[6] = z.p, curve.state.z_p
    Jl = COmJ(z.u, par)
    ev, evec = np.linalg.eig(Jl)
    i0 = int(np.argmin(np.abs(ev)))
    assert abs(ev[i0]) < 1e-10 and np.linalg.norm(COm(z.u, par)) < 1e-9          # codim2.jl:96-98
    zeta = evec[:, i0].real
    b = curve.ma.b
    assert np.allclose(b / np.linalg.norm(b), zeta * np.sign(zeta[0]) * np.sign(b[0]), atol=1e-7)   # :99-100
    ev2, evec2 = np.linalg.eig(Jl.T)
    zs = evec2[:, int(np.argmin(np.abs(ev2)))].real
    a = curve.ma.a
    assert np.allclose(a / np.linalg.norm(a), zs * np.sign(zs[0]) * np.sign(a[0]), atol=1e-7)       # :102-105
    # Bogdanov-Takens: sign change of the test function; located by a second pass with small steps from the bracketing point
    bt, k = np.array(curve.BT), np.array(curve.p2)
    cross = [j for j in range(len(bt) - 1) if bt[j] * bt[j + 1] < 0]
    assert len(cross) == 1
    j = cross[0]
    assert abs(k[j] + (k[j + 1] - k[j]) * bt[j] / (bt[j] - bt[j + 1]) - BT_K) < 1e-3
    pts = []
    C2.continuation_fold(prob, sol.u, sol.p, 6, t, t, cpf, bls, symmetric=False, normC=P.norminf,
                         callback=lambda st: pts.append((st.z_u.u.copy(), st.z_u.p, st.z_p)) or True)
    xs, q2s, ks = pts[j]
    prob.params[6] = ks
    fine = P.ContinuationPar(p_min=0.0, p_max=1.0, ds=2e-4, dsmax=2e-4, dsmin=1e-5, max_steps=40,
                             newton_options=P.NewtonPar(tol=1e-11, max_iterations=25, linsolver=ls))
    c2 = C2.continuation_fold(prob, xs, q2s, 6, curve.ma.b, curve.ma.a, fine, bls, symmetric=False, normC=P.norminf)
    prob.params[6] = PAR_COM[6]
    bt2, k2 = np.array(c2.BT), np.array(c2.p2)
    jj = [i for i in range(len(bt2) - 1) if bt2[i] * bt2[i + 1] < 0]
    assert len(jj) == 1
    i = jj[0]
    k_bt = k2[i] + (k2[i + 1] - k2[i]) * bt2[i] / (bt2[i] - bt2[i + 1])
    # independent location: the Fold at fixed k by a dense root solve (F = 0, det J = 0), then the zero of <left, right null vector>
    from scipy.optimize import brentq, fsolve

    def wv(kk):
        def H(zz):
            q = list(PAR_COM)
            q[1], q[6] = zz[3], kk
            return np.concatenate([COm(zz[:3], q), [1e3 * np.linalg.det(COmJ(zz[:3], q))]])
        zz = fsolve(H, np.array([xs[0], xs[1], xs[2], q2s]), xtol=1e-12)
        q = list(PAR_COM)
        q[1], q[6] = zz[3], kk
        Jk = COmJ(zz[:3], q)
        e1, v1 = np.linalg.eig(Jk)
        e2, w1 = np.linalg.eig(Jk.T)
        v, w = v1[:, np.argmin(np.abs(e1))].real, w1[:, np.argmin(np.abs(e2))].real
        return (w * np.sign(w[2])) @ (v * np.sign(v[1]))
    k_true = brentq(wv, 0.9712, 0.9716, xtol=1e-13)
    assert abs(k_bt - k_true) < 1e-6, (k_bt, k_true)
    # the reference's value (codim2.jl:90, pinned there at rtol 1e-5 across its own Jacobian variants) is the end point of its event
    # bisection (n_inversion = 4 halvings of dsmax = 0.01 steps): it sits 2e-4 from the zero of the test function
    assert abs(k_bt - BT_K) < 5e-4, k_bt


def test_fold_curve_known_answer_cusp_family():
    """F = r + s x - x^3: Folds at 3 x^2 = s, r = -2 x^3 -- the curve (r, s) = (-2 (s/3)^(3/2), s); continued in s from s = 1
    upwards and downwards (towards the cusp at the origin, where the p1-component of the tangent vanishes)."""
    bk = g.load_package()
    P, C2 = bk.palc, bk.codim2
    ls, bls = krylov.DefaultLS(), BlsAdapter(obls.MatrixBLS())
    F = lambda x, q: q[0] + q[1] * x - x**3
    J = lambda x, q: np.diag(q[1] - 3 * x**2)
    s0 = 1.0
    x0 = np.array([np.sqrt(s0 / 3)])
    prob = NumpyProblem2(F, J, x0 + 0.01, [-2 * x0[0] ** 3 + 0.01, s0], 0)
    one = np.array([1.0])
    sol = C2.newton_fold(prob, prob.u0, prob.p0, one, one, P.NewtonPar(tol=1e-12, max_iterations=10, linsolver=ls), bls)
    assert sol.converged and abs(sol.u[0] - x0[0]) < 1e-8 and abs(sol.p + 2 * x0[0] ** 3) < 1e-8
    for ds, smax, smin in ((0.01, 2.0, 0.0), (-0.01, 2.0, 0.05)):
        cp = P.ContinuationPar(p_min=smin, p_max=smax, ds=ds, dsmax=0.05, dsmin=1e-4, max_steps=200,
                               newton_options=P.NewtonPar(tol=1e-11, max_iterations=15, linsolver=ls))
        curve = C2.continuation_fold(prob, sol.u, sol.p, 1, one, one, cp, bls, normC=P.norminf)
        r, s = np.array(curve.p1), np.array(curve.p2)
        assert s[-1] == (smax if ds > 0 else smin) and len(s) > 10
        assert np.max(np.abs(r + 2 * (s / 3) ** 1.5)) < 1e-8
        cpv = np.abs(np.array(curve.CP))
        if ds < 0:
            assert cpv[-1] < cpv[0]                      # dr/ds -> 0 towards the cusp


def test_bordered_vec_algebra():
    """BorderedArray semantics (src/BorderedArrays.jl:49-62, 86-217) of the MA state vector"""
    bk = g.load_package()
    BV, V = bk.codim2.BorderedVec, bk.palc.V
    rng = np.random.default_rng(0)
    a, b, c = (BV(rng.standard_normal(5), rng.standard_normal()) for _ in range(3))
    cat = lambda z: np.concatenate([z.u, [z.p]])
    assert len(a) == 6
    assert np.isclose(V.dot(a, b), cat(a) @ cat(b)) and np.isclose(V.norm2(a), np.linalg.norm(cat(a)))
    assert np.isclose(V.norminf(a), np.abs(cat(a)).max()) and np.isclose(V.diffdot(a, b, c), (cat(a) - cat(b)) @ cat(c))
    y = V.copy(a)
    V.axpby(y, 2.0, b, -0.5)
    assert np.allclose(cat(y), 2 * cat(b) - 0.5 * cat(a))
    V.scale(y, 3.0)
    assert np.allclose(cat(y), 3 * (2 * cat(b) - 0.5 * cat(a)))
    z = V.zeros_like(a)
    assert np.all(cat(z) == 0) and np.all(cat(a) != 0)
    V.copyto(z, b)
    assert np.array_equal(cat(z), cat(b)) and z.u is not b.u
    nan = BV(np.array([1.0, np.nan]), 0.0)
    assert np.isnan(V.norminf(nan))


# ------------------------------------------------------------------------------------------------ Hopf curves
from tests.test_host_logic_cpu import DenseComplexProblem, _dense_cls, _dense_ls2  # noqa: E402


class DenseComplexProblem2:
    """complexified twin of a NumpyProblem2 (cprob of codim2.HopfMinAug): J(x, p, transpose) -> callable on complex vectors"""

    def __init__(self, prob):
        self.prob, self.params, self.lens = prob, prob.params, prob.lens  # the SAME parameter list: one problem, two views

    def J(self, x, p, transpose=False):
        M = np.asarray(self.prob.J(x, p), dtype=float)
        return DenseComplexProblem.Jc(M.T.copy() if transpose else M)


def test_hopf_curve_of_the_brusselator_is_analytic():
    """x' = a - (b+1) x + x^2 y, y' = b x - x^2 y: Hopf points at b = 1 + a^2, omega = a, equilibrium (a, b / a) --
    continuation_hopf in (b; a) from a = 1.3 up to a = 2 and down to a = 0.6 reproduces the curve"""
    bk = g.load_package()
    P, C2 = bk.palc, bk.codim2
    F = lambda x, q: np.array([q[0] - (q[1] + 1) * x[0] + x[0] ** 2 * x[1], q[1] * x[0] - x[0] ** 2 * x[1]])
    J = lambda x, q: np.array([[-(q[1] + 1) + 2 * x[0] * x[1], x[0] ** 2], [q[1] - 2 * x[0] * x[1], -x[0] ** 2]])
    a0 = 1.3
    b0 = 1 + a0 * a0
    x0 = np.array([a0, b0 / a0])
    prob = NumpyProblem2(F, J, x0, [a0, b0], 1)
    cprob = DenseComplexProblem2(prob)
    vals, vecs = np.linalg.eig(J(x0, [a0, b0]))
    k = int(np.argmax(vals.imag))
    valt, vect = np.linalg.eig(J(x0, [a0, b0]).T)
    kt = int(np.argmin(valt.imag))
    opts = P.NewtonPar(tol=1e-10, max_iterations=15, linsolver=krylov.DefaultLS())
    for ds, amin, amax in ((0.01, 0.5, 2.0), (-0.01, 0.6, 2.5)):
        cp = P.ContinuationPar(p_min=amin, p_max=amax, ds=ds, dsmax=0.05, dsmin=1e-4, max_steps=200, newton_options=opts)
        curve = C2.continuation_hopf(prob, cprob, x0, b0, a0, 0, vecs[:, k], vect[:, kt], cp, _dense_ls2, _dense_cls)
        a, b, om = np.array(curve.p2), np.array(curve.p1), np.array(curve.omega)
        assert a[-1] == (amax if ds > 0 else amin) and len(a) > 10 and not curve.stopped_at_bt
        assert np.max(np.abs(b - (1 + a * a))) < 1e-8 and np.max(np.abs(np.abs(om) - a)) < 1e-8
        assert prob.params == [a0, b0]
        xs = curve.state.z_u.u
        assert np.allclose(xs, [a[-1], b[-1] / a[-1]], atol=1e-8)
        assert max(r["itnewton"] for r in curve.rows) <= 4


def test_hopf_curve_of_the_co_model_ends_on_the_bogdanov_takens_point(com_fold):
    """test/fold_codim_2/codim2.jl:118-146 continues the Hopf point of the CO branch in (q2, k).  Towards larger k that curve
    ends where its frequency vanishes -- on the Bogdanov-Takens point of the Fold curve (test_fold_curve_of_the_co_model:
    the one at k = 0.7223 on the curve through the branch's first fold): two independent continuations (real bordered systems
    there, complex shifted ones here) meet."""
    bk, prob, ls, bls, sol, t = com_fold
    P, C2 = bk.palc, bk.codim2
    # a Hopf point on the q2-branch at k = 0.4: scan the branch for a complex pair crossing the imaginary axis
    cp = P.ContinuationPar(p_min=0.6, p_max=2.5, ds=-0.002, dsmax=0.01, dsmin=1e-4, max_steps=3000,
                           newton_options=P.NewtonPar(tol=1e-10, max_iterations=25, linsolver=ls))
    pts = []
    P.continuation(prob, P.PALC(bls=bls), cp, normC=P.norminf, callback=lambda s: pts.append((s.z_u.copy(), s.z_p)) or True)
    lead = []
    for x, p in pts:
        ev = np.linalg.eigvals(COmJ(x, prob._par(p)))
        c = ev[np.abs(ev.imag) > 1e-6]
        lead.append(c.real.max() if len(c) else np.nan)
    lead = np.array(lead)
    idx = [i for i in range(len(lead) - 1) if np.isfinite(lead[i]) and np.isfinite(lead[i + 1]) and lead[i] * lead[i + 1] < 0]
    assert idx, "no Hopf point on the branch"
    x0, p0 = pts[idx[0]]
    Jh = COmJ(x0, prob._par(p0))
    vals, vecs = np.linalg.eig(Jh)
    kk = int(np.argmax(vals.imag))
    valt, vect = np.linalg.eig(Jh.T)
    kt = int(np.argmin(valt.imag))
    cprob = DenseComplexProblem2(prob)
    opts = P.NewtonPar(tol=1e-10, max_iterations=15, linsolver=ls)
    hp = C2.newton_hopf(prob, cprob, x0, p0, vals[kk].imag, vecs[:, kk], vect[:, kt], opts, _dense_ls2, _dense_cls)
    assert hp.converged, hp.residuals
    ev = np.linalg.eigvals(COmJ(hp.u, prob._par(hp.p)))
    assert abs(ev[np.argmax(ev.imag)].real) < 1e-8 and abs(ev.imag.max() - abs(hp.omega)) < 1e-8   # codim2.jl:123-131
    cph = P.ContinuationPar(p_min=0.0, p_max=1.0, ds=0.002, dsmax=0.01, dsmin=1e-6, max_steps=400, newton_options=opts)
    curve = C2.continuation_hopf(prob, cprob, hp.u, hp.p, hp.omega, 6, vecs[:, kk], vect[:, kt], cph, _dense_ls2, _dense_cls)
    om, k = np.abs(np.array(curve.omega)), np.array(curve.p2)
    assert curve.stopped_at_bt and k[0] == 0.4 and om[-1] < 1e-8 and prob.params == PAR_COM
    # omega^2 vanishes linearly at a Bogdanov-Takens point: extrapolate the last genuine Hopf points to omega = 0
    m = np.where(om > 1e-3)[0][-4:]
    k_hopf_end = np.polyval(np.polyfit(om[m] ** 2, k[m], 1), 0.0)
    # the Fold curve through the FIRST fold of the branch (q2 = 1.0420) and the zero of its Bogdanov-Takens test function
    pts2 = []
    rows = P.continuation(prob, P.PALC(bls=bls), cp, normC=P.norminf, callback=lambda s: pts2.append((s.z_u.copy(), s.z_p, s.tau_u.copy())) or True)[0]
    ps = [r["param"] for r in rows]
    i1 = next(i for i in range(1, len(ps) - 1) if (ps[i] - ps[i - 1]) * (ps[i + 1] - ps[i]) < 0)
    xf, pf, tau = pts2[i1]
    tf = tau / np.linalg.norm(tau)
    f1 = C2.newton_fold(prob, xf, pf, tf, tf, P.NewtonPar(tol=1e-10, max_iterations=10, linsolver=ls), bls, symmetric=False)
    assert f1.converged and abs(f1.p - 1.042048505) < 1e-8
    cpf = P.ContinuationPar(p_min=0.0, p_max=1.0, ds=0.002, dsmax=0.005, dsmin=1e-4, max_steps=200,
                            newton_options=P.NewtonPar(tol=1e-10, max_iterations=25, linsolver=ls))
    fc = C2.continuation_fold(prob, f1.u, f1.p, 6, tf, tf, cpf, bls, symmetric=False, normC=P.norminf)
    bt, kf = np.array(fc.BT), np.array(fc.p2)
    j = next(i for i in range(len(bt) - 1) if bt[i] * bt[i + 1] < 0)
    k_bt = kf[j] + (kf[j + 1] - kf[j]) * bt[j] / (bt[j] - bt[j + 1])
    assert 0.70 < k_bt < 0.74 and abs(k_hopf_end - k_bt) < 1e-3, (k_hopf_end, k_bt)
    # ... and the (q2, k) of the two curves agree there as well
    q2_hopf_end = np.polyval(np.polyfit(om[m] ** 2, np.array(curve.p1)[m], 1), 0.0)
    q2_bt = fc.p1[j] + (fc.p1[j + 1] - fc.p1[j]) * bt[j] / (bt[j] - bt[j + 1])
    assert abs(q2_hopf_end - q2_bt) < 1e-3, (q2_hopf_end, q2_bt)


@pytest.mark.parametrize("vector_p", [False, True])
def test_bordered_vec_reference_cases(vector_p):
    """test/linear_solvers/bordered_arrays.jl:17-60 (x = BorderedArray([1, 2], 3), y = BorderedArray([4, 5], 6), scalar and
    one-element-vector p): zerovector, scale, add!, inner on the concatenated entries; test_linear.jl:13-34: length == 11"""
    bk = g.load_package()
    BV, V = bk.codim2.BorderedVec, bk.palc.V
    mk = lambda u, p: BV(np.array(u, dtype=float), np.array([p]) if vector_p else p)
    cat = lambda z: np.concatenate([z.u, np.atleast_1d(z.p)])
    x, y = mk([1.0, 2.0], 3.0), mk([4.0, 5.0], 6.0)
    z = V.zeros_like(x)
    assert np.all(cat(z) == 0.0) and np.array_equal(cat(x), [1, 2, 3])           # zerovector leaves x alone
    alpha = -0.7364
    xs = V.copy(x)
    V.scale(xs, alpha)
    assert np.allclose(cat(xs), alpha * cat(x)) and np.array_equal(cat(x), [1, 2, 3])
    ya = V.copy(y)
    V.axpby(ya, 2.0 / 3, x, 1.0)                                                   # VI.add!(y, x, 2/3, 1)
    assert np.allclose(cat(ya), cat(y) + 2.0 / 3 * cat(x)) and np.array_equal(cat(x), [1, 2, 3])
    assert V.dot(x, y) == 1 * 4 + 2 * 5 + 3 * 6 and len(x) == 3
    assert len(BV(np.random.default_rng(0).random(10), 1.0)) == 11
    c = V.copy(x)
    V.scale(c, 2.0)
    assert np.array_equal(cat(x), [1, 2, 3])                                       # copies do not alias (neither u nor p)
    V.copyto(c, y)
    V.scale(c, 2.0)
    assert np.array_equal(cat(y), [4, 5, 6])


# ------------------------------------------------------------------------------------------------ test/hopf_codim_2/testHopfMA.jl
def Fbru(x, q):
    """1-D Brusselator, test/hopf_codim_2/testHopfMA.jl:7-29 (q = (alpha, beta, D1, D2, l), Dirichlet values alpha, beta / alpha)"""
    al, be, D1, D2, l = q
    n = len(x) // 2
    h2 = (1.0 / n) ** 2
    c1, c2 = D1 / l**2 / h2, D2 / l**2 / h2
    u, v = x[:n], x[n:]
    up = np.concatenate([[al], u, [al]])
    vp = np.concatenate([[be / al], v, [be / al]])
    f = np.empty_like(x)
    f[:n] = c1 * (up[:-2] - 2 * u + up[2:]) + al - (be + 1) * u + u**2 * v
    f[n:] = c2 * (vp[:-2] - 2 * v + vp[2:]) + be * u - u**2 * v
    return f


def Jbru(x, q):
    """:33-64 (Jbru_sp, dense here)"""
    al, be, D1, D2, l = q
    n = len(x) // 2
    h2 = (1.0 / n) ** 2
    c1, c2 = D1 / l**2 / h2, D2 / l**2 / h2
    u, v = x[:n], x[n:]
    T = lambda c: c * (np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1) - 2 * np.eye(n))
    return np.block([[T(c1) + np.diag(-(be + 1) + 2 * u * v), np.diag(u * u)],
                     [np.diag(be - 2 * u * v), T(c2) + np.diag(-u * u)]])


def test_hopf_ma_linear_solver_on_the_brusselator_pde():
    """testHopfMA.jl:67-141 on its own problem (n = 10, continuation in l from the uniform state): the first Hopf point of the
    branch refined by newton_hopf, then the structural check of the reference test -- the linear solver of the minimally
    augmented system (HopfLinearSolverMinAug, sigma_x / sigma_p / sigma_omega by finite differences) against the
    finite-difference Jacobian of the functional (x, p, omega) -> (F, Re sigma, Im sigma) -- and 3 steps of the Hopf curve in beta
    (:159-162)."""
    bk = g.load_package()
    P, C2 = bk.palc, bk.codim2
    n = 10
    par = [2.0, 5.45, 0.008, 0.004, 0.3]
    x0 = np.concatenate([np.full(n, par[0]), np.full(n, par[1] / par[0])])      # the uniform state solves F = 0 for every l
    assert np.linalg.norm(Fbru(x0, par)) < 1e-12
    E = np.eye(2 * n)
    Jfd = np.column_stack([(Fbru(x0 + 1e-6 * E[:, j], par) - Fbru(x0 - 1e-6 * E[:, j], par)) / 2e-6 for j in range(2 * n)])
    assert np.max(np.abs(Jfd - Jbru(x0, par))) < 1e-6                            # :72-73 (Jbru_sp == Jbru_ana)
    prob = NumpyProblem2(Fbru, Jbru, x0, par, 4)
    cprob = DenseComplexProblem2(prob)
    # first Hopf point along l (the branch of :76-77): leading complex pair crosses the axis
    ls_ = np.linspace(0.3, 1.8, 301)
    lead = [max((ev.real for ev in np.linalg.eigvals(Jbru(x0, prob._par(l))) if abs(ev.imag) > 1e-8), default=np.nan) for l in ls_]
    i = next(k for k in range(len(ls_) - 1) if lead[k] < 0 <= lead[k + 1])
    l0 = ls_[i + 1]
    vals, vecs = np.linalg.eig(Jbru(x0, prob._par(l0)))
    kk = int(np.argmax(np.where(np.abs(vals.imag) > 1e-8, vals.real, -np.inf) + 1e-9 * np.sign(vals.imag)))
    valt, vect = np.linalg.eig(Jbru(x0, prob._par(l0)).T)
    kt = int(np.argmin(np.abs(valt - np.conj(vals[kk]))))
    opts = P.NewtonPar(tol=1e-11, max_iterations=15, linsolver=krylov.DefaultLS())
    hp = C2.newton_hopf(prob, cprob, x0, l0, vals[kk].imag, vecs[:, kk], vect[:, kt], opts, _dense_ls2, _dense_cls)
    assert hp.converged and hp.itnewton <= 6, hp.residuals                        # :80-81, 139-140
    ev = np.linalg.eigvals(Jbru(hp.u, prob._par(hp.p)))
    j = int(np.argmin(np.abs(ev - 1j * abs(hp.omega))))
    assert abs(ev[j].real) < 1e-9 and abs(ev[j].imag - abs(hp.omega)) < 1e-9
    # the MA functional and its linear solver at the Hopf point (:92-137)
    ma = C2.HopfMinAug(prob, cprob, vect[:, kt], vecs[:, kk], _dense_ls2, _dense_cls)

    def H(z):
        F, sr, si = ma.residual(z[:-2].copy(), z[-2], z[-1])
        return np.concatenate([F, [sr, si]])
    z0 = np.concatenate([hp.u, [hp.p, hp.omega]])
    m = len(z0)
    I = np.eye(m)
    eps = 1e-6
    Jma = np.column_stack([(H(z0 + eps * I[:, c]) - H(z0 - eps * I[:, c])) / (2 * eps) for c in range(m)])
    rhs = np.random.default_rng(5).random(m)
    sol_fd = np.linalg.solve(Jma, rhs)
    dX, dp, dom, _ = ma.solve(hp.u.copy(), hp.p, hp.omega, rhs[:-2].copy(), rhs[-2], rhs[-1])
    sol_ma = np.concatenate([dX, [dp, dom]])
    assert np.linalg.norm(sol_ma - sol_fd) < 1e-3 * np.linalg.norm(sol_fd), (sol_ma[-2:], sol_fd[-2:])
    # sigma_p, sigma_omega, sigma_x of the solver against the finite-difference Jacobian (:124-137: rtol 1e-4, 1e-4, 1e-3)
    v, w, dpF, sigma_p, sigma_om = ma.bordered_terms(hp.u, hp.p, hp.omega)
    assert abs(sigma_p - complex(Jma[-2, -2], Jma[-1, -2])) < 1e-4 * abs(sigma_p)
    assert abs(sigma_om - complex(Jma[-2, -1], Jma[-1, -1])) < 1e-4 * abs(sigma_om)
    assert np.max(np.abs(Jma[:-2, -2] - dpF)) < 1e-5 * max(1.0, np.max(np.abs(dpF))) and np.max(np.abs(Jma[:-2, -1])) < 1e-7
    # three steps of the Hopf curve in beta (:159-162)
    cp = P.ContinuationPar(dsmin=0.001, dsmax=0.05, ds=0.01, p_min=0.0, p_max=6.5, max_steps=3, newton_options=opts)
    prob.params[4] = hp.p
    prob.lens = 4
    curve = C2.continuation_hopf(prob, cprob, hp.u, hp.p, hp.omega, 1, vecs[:, kk], vect[:, kt], cp, _dense_ls2, _dense_cls)
    assert len(curve.rows) == 4 and curve.p2[0] == 5.45 and curve.p2[-1] > 5.45
    for l_k, be_k, om_k in zip(curve.p1, curve.p2, curve.omega):
        q = list(par)
        q[4], q[1] = l_k, be_k
        xk = np.concatenate([np.full(n, q[0]), np.full(n, q[1] / q[0])])
        ev = np.linalg.eigvals(Jbru(xk, q))
        assert np.min(np.abs(ev - 1j * abs(om_k))) < 1e-8


def test_fold_ma_linear_solver_against_finite_differences(com_fold):
    """test/fold_codim_2/testJacobianFoldDeflation.jl pattern on the CO model: the linear solver of the minimally augmented Fold
    system (foldMALinearSolver, sigma_x / sigma_p by finite differences, MinAugFold.jl:122-146) against the finite-difference
    Jacobian of (x, p) -> (F, sigma) at the Fold point -- for a non-symmetric Jacobian (J' through jacobian_adjoint)"""
    bk, prob, ls, bls, sol, t = com_fold
    C2 = bk.codim2
    rng = np.random.default_rng(7)
    a, b = t + 0.1 * rng.standard_normal(3), t + 0.1 * rng.standard_normal(3)     # generic border vectors
    ma = C2.FoldMinAug(prob, a, b, bls, symmetric=False)

    def H(z):
        F, sigma = ma.residual(z[:-1].copy(), z[-1])
        return np.concatenate([F, [sigma]])
    z0 = np.concatenate([sol.u, [sol.p]]) + 1e-3 * rng.standard_normal(4)          # near, not on, the Fold: sigma != 0
    I = np.eye(4)
    eps = 1e-6
    Jma = np.column_stack([(H(z0 + eps * I[:, c]) - H(z0 - eps * I[:, c])) / (2 * eps) for c in range(4)])
    rhs = rng.random(4)
    dX, dp, cv = ma.solve(z0[:-1].copy(), z0[-1], rhs[:-1].copy(), rhs[-1])
    ref = np.linalg.solve(Jma, rhs)
    assert cv and np.linalg.norm(np.concatenate([dX, [dp]]) - ref) < 1e-4 * np.linalg.norm(ref)
    v, w, dpF, sigma_p = ma.bordered_terms(z0[:-1].copy(), z0[-1])
    assert abs(sigma_p - Jma[-1, -1]) < 1e-5 * max(1.0, abs(sigma_p)) and np.max(np.abs(dpF - Jma[:-1, -1])) < 1e-6


def test_bordered_vec_zero_is_exact_even_from_nan():
    bk = g.load_package()
    BV, V = bk.codim2.BorderedVec, bk.palc.V
    for p in (np.nan, np.array([np.nan, 1.0])):
        z = V.zeros_like(BV(np.array([np.nan, 2.0]), p))
        assert np.all(z.u == 0.0) and np.all(np.atleast_1d(z.p) == 0.0)


def test_bogdanov_takens_point_located_as_in_the_reference(com_fold):
    """test/fold_codim_2/codim2.jl:75-90 with its own options -- continuation(br, 3, (@optic _.k), ContinuationPar(opts_br, p_max = 1.,
    p_min = 0., max_steps = 50, detect_event = 2), update_minaug_every_step = 1, jacobian_ma = MinAug()); opts_br: ds = 0.002, dsmax = 0.01,
    n_inversion = 4, max_bisection_steps = 25, Newton tol 1e-12 (the NewtonPar default) with max_iterations = 10, normC = norm:
        sn_br.specialpoint[1].type == :bt,   sn_br.specialpoint[1].param ≈ 0.9716038596420551  (rtol 1e-5 there).
    The located parameter is the end point of the event bisection, so it pins the whole chain -- the branch in q2, the Fold
    refinement, 47 adaptive PALC steps on the minimally augmented system and the bisection -- not just the Bogdanov-Takens point."""
    bk, prob, ls, bls, sol, t = com_fold
    P, C2 = bk.palc, bk.codim2
    t2 = t / np.linalg.norm(t)
    s12 = C2.newton_fold(prob, sol.u, sol.p, t2, t2, P.NewtonPar(tol=1e-12, max_iterations=10, linsolver=ls), bls, symmetric=False)
    assert s12.converged
    cpf = P.ContinuationPar(p_min=0.0, p_max=1.0, ds=0.002, dsmax=0.01, dsmin=1e-4, max_steps=50, n_inversion=4, max_bisection_steps=25,
                            newton_options=P.NewtonPar(tol=1e-12, max_iterations=10, linsolver=ls))
    curve = C2.continuation_fold(prob, s12.u, s12.p, 6, t2, t2, cpf, bls, symmetric=False, normC=P.norm2, detect_event=2)
    bts = [sp for sp in curve.specialpoint if sp.type == "bt"]
    assert len(curve.specialpoint) >= 1 and curve.specialpoint[0].type == "bt" and len(bts) == 1
    bt = bts[0]
    assert bt.status == "converged" and bt.interval[0] <= bt.param <= bt.interval[1] and bt.interval[1] - bt.interval[0] < 3e-4
    assert abs(bt.param - BT_K) < 1e-5 * BT_K                        # the reference's own tolerance
    assert abs(bt.param - BT_K) < 1e-9, bt.param - BT_K              # ... and in fact the same number (4e-13 on this host)
    assert prob.params == PAR_COM
    # detect_event = 1: the same crossing recorded without the bisection
    c1 = C2.continuation_fold(prob, s12.u, s12.p, 6, t2, t2, cpf, bls, symmetric=False, normC=P.norm2, detect_event=1)
    assert [sp.type for sp in c1.specialpoint] == ["bt"] and c1.specialpoint[0].status == "guess"
    assert c1.specialpoint[0].interval[0] < 0.9713976 < c1.specialpoint[0].interval[1]


# ------------------------------------------------------------------------------------------------ object vectors (the device code path)
class ObjVec:
    """Host stand-in with the method set of core.DeviceVec: codim2 / palc see "a vector that carries its own algebra" exactly as on the
    device (palc._obj), so the object-vector branches of BorderedVec and of the curve drivers run in the CPU suite too."""

    def __init__(self, a):
        self.a = np.array(a, dtype=float)

    def __len__(self):
        return len(self.a)

    def numpy(self):
        return self.a.copy()

    def copy(self):
        return ObjVec(self.a)

    def copyto(self, src):
        self.a[...] = src.a
        return self

    def zero_(self):
        self.a[...] = 0.0
        return self

    def scale_(self, s):
        self.a *= s
        return self

    def axpby_(self, a, x, b=1.0):
        self.a = a * x.a + b * self.a
        return self

    def dot(self, y):
        return float(self.a @ y.a)

    def norm(self):
        return float(np.linalg.norm(self.a))

    def norminf(self):
        return float(np.max(np.abs(self.a)))

    def diffdot(self, x0, tau):
        return float((self.a - x0.a) @ tau.a)


class ObjProblem2(NumpyProblem2):
    """NumpyProblem2 whose state vectors are ObjVec"""

    def F(self, x, p, out=None):
        r = ObjVec(self.F_(x.a, self._par(p)))
        return r if out is None else out.copyto(r)

    def J(self, x, p):
        return self.J_(x.a, self._par(p))

    def Jt(self, x, p):
        return self.J_(x.a, self._par(p)).T


class ObjBls:
    """bordered solver on ObjVec through the oracle's dense MatrixBLS; a Jacobian applied to an ObjVec returns an ObjVec"""

    def __init__(self):
        self.inner = BlsAdapter(obls.MatrixBLS())

    def __call__(self, J, dR, dzu, dzp, R, n, xiu=1.0, xip=1.0, shift=None, dotscale=1.0):
        u, up, ok, it = self.inner(J, dR.a, dzu.a, dzp, R.a, n, xiu, xip, shift=shift, dotscale=dotscale)
        return ObjVec(u), up, ok, it


def test_fold_curve_on_object_vectors_matches_the_array_run(com_fold, monkeypatch):
    """the CO Fold curve with event location, once on ndarrays and once on DeviceVec-like objects: same points, same special point"""
    bk, prob, ls, bls, sol, t = com_fold
    P, C2 = bk.palc, bk.codim2
    monkeypatch.setattr(C2, "_apply", lambda J, v: ObjVec(J @ v.a) if isinstance(v, ObjVec) else (J(v) if callable(J) else J @ v))
    cpf = P.ContinuationPar(p_min=0.0, p_max=1.0, ds=0.002, dsmax=0.01, dsmin=1e-4, max_steps=50, n_inversion=4, max_bisection_steps=25,
                            newton_options=P.NewtonPar(tol=1e-12, max_iterations=10, linsolver=ls))
    ref = C2.continuation_fold(prob, sol.u, sol.p, 6, t, t, cpf, bls, symmetric=False, normC=P.norm2, detect_event=2)
    oprob = ObjProblem2(COm, COmJ, ObjVec(prob.u0), PAR_COM, 1)
    cur = C2.continuation_fold(oprob, ObjVec(sol.u), sol.p, 6, ObjVec(t), ObjVec(t), cpf, ObjBls(), symmetric=False, normC=P.norm2, detect_event=2)
    assert len(cur.rows) == len(ref.rows) and np.allclose(cur.p1, ref.p1, rtol=0, atol=1e-12) and np.allclose(cur.p2, ref.p2, rtol=0, atol=1e-12)
    assert [s.type for s in cur.specialpoint] == [s.type for s in ref.specialpoint] == ["bt"]
    assert abs(cur.specialpoint[0].param - ref.specialpoint[0].param) < 1e-12 and isinstance(cur.specialpoint[0].x, ObjVec)
    assert isinstance(cur.state.z_u.u, ObjVec) and oprob.params == PAR_COM
