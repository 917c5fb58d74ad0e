"""Numbers the reference's own test-suite prints for the CO-oxidation model, reproduced by the product's HOST logic (palc.py, events.py,
codim2.py) with the oracle's dense solvers as the linear-algebra backend -- test/hopf_codim_2/COModel.jl:19-59.

These values are not bifurcation parameters to full precision: they are the END POINTS of the reference's bisections (n_inversion
halvings of adaptive PALC steps).  They depend on every step length, every Newton iteration count and every reversal along the way,
so matching them to the printed digits -- and the one 16-digit value to 2e-14 -- pins the arithmetic of the whole host chain against
the reference itself, which this image cannot run."""
import numpy as np
import pytest

import __graft_entry__ as g
from oracle import krylov, bls as obls
from tests.test_host_logic_cpu import BlsAdapter
from tests.test_codim2_curves_cpu import NumpyProblem2, COm, COmJ

PAR = [2.5, 1.0, 10.0, 0.0675, 1.0, 0.1, 0.4]                  # par_com, COModel.jl:20
Z0 = np.array([0.001137, 0.891483, 0.062345])                  # :22


def _isapprox(a, b, rtol=np.sqrt(np.finfo(float).eps)):
    """Julia's isapprox default: |a - b| <= sqrt(eps) max(|a|, |b|)"""
    return abs(a - b) <= rtol * max(abs(a), abs(b))


def _dense_eig(J, nev):
    vals, vecs = np.linalg.eig(np.asarray(J))
    o = np.argsort(-vals.real)                                  # decreasing real part (src/EigSolver.jl:16-19)
    return vals[o][:nev], vecs[:, o][:, :nev], True, 1


@pytest.fixture(scope="module")
def co():
    bk = g.load_package()
    return bk, NumpyProblem2(COm, COmJ, Z0.copy(), PAR, 1), krylov.DefaultLS(), BlsAdapter(obls.MatrixBLS())


def test_special_points_of_the_co_branch(co):
    """COModel.jl:27-34: continuation(prob, PALC(), ContinuationPar(p_min = 0.5, p_max = 2.3, ds = 0.002, dsmax = 0.01, n_inversion = 6,
    detect_bifurcation = 3, max_bisection_steps = 25, nev = 3, max_steps = 100); normC = norminf, bothside = true):
    specialpoint[2..5].param ≈ 1.04099606, 1.05220029, 1.04204851, 1.05158367  (Hopf, fold, fold, Hopf; the forward half)."""
    bk, prob, ls, bls = co
    P, E = bk.palc, bk.events
    cp = P.ContinuationPar(p_min=0.5, p_max=2.3, ds=0.002, dsmax=0.01, dsmin=1e-4, max_steps=100, n_inversion=6, max_bisection_steps=25, nev=3,
                           detect_bifurcation=3, newton_options=P.NewtonPar(tol=1e-12, max_iterations=25, linsolver=ls, eigsolver=_dense_eig))
    br = E.continuation(prob, P.PALC(bls=bls), cp, normC=P.norminf)
    sp = [s for s in br.specialpoint if s.type != "endpoint"]
    assert [s.type for s in sp] == ["hopf", "bp", "bp", "hopf"] and all(s.status == "converged" for s in sp)
    for s, gold in zip(sp, (1.04099606, 1.05220029, 1.04204851, 1.05158367)):
        assert _isapprox(s.param, gold), (s.type, s.param, gold)
    # none of these is the bifurcation value itself: the Hopf points of the model sit at 1.04099157 and 1.05155746
    assert abs(sp[0].param - 1.04099157) > 3e-6 and abs(sp[3].param - 1.05155746) > 2e-5


def test_special_points_of_the_fold_curve(co):
    """COModel.jl:36-59: sn_codim2 = continuation(br, 3, (@optic _.k), ContinuationPar(opts_br, p_max = 2.2, p_min = 0., ds = -0.001,
    dsmax = 0.05, n_inversion = 8, max_steps = 50); normC = norminf, detect_codim2_bifurcation = 2, update_minaug_every_step = 1,
    bothside = true) with NewtonPar(max_iterations = 10, tol = 1e-12):
        specialpoint[2] bt   param ≈ 0.97139757 (atol 1e-5),  printsol (k, q2) ≈ (0.971397, 1.417628) rtol 1e-4
        specialpoint[3] cusp param ≈ 0.35665351 (rtol 1e-4)
        specialpoint[4] bt   param ≈ 0.7223392465523879,      printsol (k, q2) ≈ (0.722339, 1.161199) rtol 1e-4"""
    bk, prob, ls, bls = co
    P, C2 = bk.palc, bk.codim2
    # the Fold br.specialpoint[3] (q2 = 1.05220029): second turning point of the branch, refined on the minimally augmented system
    cp = P.ContinuationPar(p_min=0.5, p_max=2.3, ds=0.002, dsmax=0.01, dsmin=1e-4, max_steps=400, newton_options=P.NewtonPar(tol=1e-12, max_iterations=25, linsolver=ls))
    pts = []
    rows, _ = P.continuation(prob, P.PALC(bls=bls), cp, normC=P.norminf, callback=lambda s: pts.append((s.z_u.copy(), s.z_p, s.tau_u.copy())) or True)
    ps = [r["param"] for r in rows]
    i = next(k for k in range(1, len(ps) - 1) if ps[k] > ps[k - 1] and ps[k] > ps[k + 1])
    x0, p0, tau = pts[i]
    t = tau / np.linalg.norm(tau)
    nopt = P.NewtonPar(tol=1e-12, max_iterations=10, linsolver=ls)
    f3 = C2.newton_fold(prob, x0, p0, t, t, nopt, bls, symmetric=False)
    assert f3.converged and _isapprox(f3.p, 1.05220029)
    found = []
    for ds in (-0.001, 0.001):                                   # bothside = true: the two directions from the same Fold point
        cpf = P.ContinuationPar(p_min=0.0, p_max=2.2, ds=ds, dsmax=0.05, dsmin=1e-4, max_steps=50, n_inversion=8, max_bisection_steps=25, newton_options=nopt)
        curve = C2.continuation_fold(prob, f3.u, f3.p, 6, t, t, cpf, bls, symmetric=False, normC=P.norminf, detect_event=2)
        found += curve.specialpoint
    assert sorted(s.type for s in found) == ["bt", "bt", "cusp"]
    bt_hi = next(s for s in found if s.type == "bt" and s.param > 0.9)
    bt_lo = next(s for s in found if s.type == "bt" and s.param < 0.9)
    cusp = next(s for s in found if s.type == "cusp")
    assert abs(bt_hi.param - 0.97139757) < 1e-5 and _isapprox(bt_hi.param, 0.971397, 1e-4) and _isapprox(bt_hi.p1, 1.417628, 1e-4)
    assert _isapprox(cusp.param, 0.35665351, 1e-4)
    assert _isapprox(bt_lo.param, 0.7223392465523879)            # the reference's assertion (default isapprox) ...
    assert abs(bt_lo.param - 0.7223392465523879) < 1e-9, bt_lo.param - 0.7223392465523879    # ... and far inside it (2e-14 on this host)
    assert _isapprox(bt_lo.param, 0.722339, 1e-4) and _isapprox(bt_lo.p1, 1.161199, 1e-4)
    assert bt_lo.status == bt_hi.status == "converged"


def Lor(u, q):
    """test/hopf_codim_2/lorenz84.jl:7-16, q = (alpha, beta, G, delta, gamma, T, F)"""
    al, be, G, de, ga, T, F = q
    X, Y, Z, U = u
    return np.array([-Y**2 - Z**2 - al * X + al * F - ga * U**2, X * Y - be * X * Z - Y + G, be * X * Y + X * Z - Z, -de * U + ga * U * X + T])


def JLor(u, q):
    """:18-27"""
    al, be, G, de, ga, T, F = q
    X, Y, Z, U = u
    return np.array([[-al, -2 * Y, -2 * Z, -2 * ga * U], [Y - be * Z, X - 1, -be * X, 0], [be * Y + Z, be * X, X - 1, 0], [ga * U, 0, 0, -de + ga * X]])


PAR_LOR = [0.25, 1.0, 0.25, 1.04, 0.987, 0.04, 3.0]             # (alpha, beta, G, delta, gamma, T, F), :31
Z0_LOR = np.array([2.9787004394953343, -0.03868302503393752, 0.058232737694740085, -0.02105288273117459])   # :43


def test_special_point_intervals_of_the_lorenz84_branch():
    """test/hopf_codim_2/lorenz84.jl:7-65: continuation(prob, PALC(tangent = Bordered()), ContinuationPar(p_min = -1.5, p_max = 3.0,
    ds = 0.001, dsmax = 0.025, detect_bifurcation = 3, n_inversion = 6, max_bisection_steps = 25, nev = 4, max_steps = 252);
    normC = norminf, bothside = true) -- the bisection INTERVALS of its four special points, 16 digits each:
        (2.859863413561998, 2.859897757930758) (2.467211879219629, 2.467246154619121)
        (1.619657484413436, 1.6196654620692468) (1.546648372620807, 1.5466483727182652)
    (the half of `bothside` that leaves F = 3 = p_max downwards; the other half has nowhere to go).  Bordered tangent, MatrixBLS."""
    bk = g.load_package()
    P, E = bk.palc, bk.events

    ls, bls = krylov.DefaultLS(), BlsAdapter(obls.MatrixBLS())
    prob = NumpyProblem2(Lor, JLor, Z0_LOR.copy(), PAR_LOR, 6)
    cp = P.ContinuationPar(p_min=-1.5, p_max=3.0, ds=-0.001, dsmax=0.025, dsmin=1e-4, max_steps=252, n_inversion=6, max_bisection_steps=25, nev=4,
                           detect_bifurcation=3, newton_options=P.NewtonPar(tol=1e-12, max_iterations=25, linsolver=ls, eigsolver=_dense_eig))
    br = E.continuation(prob, P.PALC(tangent="bordered", bls=bls), cp, normC=P.norminf)
    sp = [s for s in br.specialpoint if s.type != "endpoint"]
    gold = [(1.546648372620807, 1.5466483727182652), (1.619657484413436, 1.6196654620692468),
            (2.467211879219629, 2.467246154619121), (2.859863413561998, 2.859897757930758)]
    assert [s.type for s in sp] == ["bp", "hopf", "hopf", "hopf"]   # the fold of the branch (one real eigenvalue), then three Hopf points
    for s, (lo, hi) in zip(sp, gold):
        assert _isapprox(s.interval[0], lo) and _isapprox(s.interval[1], hi), (s.type, s.interval, (lo, hi))   # lorenz84.jl:62-65
        assert s.status == "converged" and s.interval[0] <= s.param <= s.interval[1]
    assert abs(sp[0].interval[0] - gold[0][0]) < 1e-9 and abs(sp[0].interval[1] - gold[0][1]) < 1e-9   # 1e-13 on this host
    assert br.specialpoint[-1].type == "endpoint" and br.specialpoint[-1].param == 3.0


def test_special_points_of_the_lorenz84_fold_curve():
    """lorenz84.jl:70-84: the Fold br.specialpoint[5] (F = 1.5466) continued in T with ContinuationPar(opts_br, p_max = 3.2, p_min = -0.1,
    detect_bifurcation = 1, dsmin = 1e-5, ds = -0.001, dsmax = 0.005, max_steps = 60), detect_codim2_bifurcation = 1 (events recorded at
    the point after the change, no bisection):
        specialpoint[1..4].param ≈ +0.02058724 (rtol 1e-5), +0.00004983 (atol 1e-8), -0.00045281 (rtol 1e-5), -0.02135893 (rtol 1e-5)
    -- Bogdanov-Takens, Zero-Hopf, Zero-Hopf, Bogdanov-Takens: step points of the adaptive PALC run on the minimally augmented system."""
    bk = g.load_package()
    P, C2 = bk.palc, bk.codim2
    ls, bls = krylov.DefaultLS(), BlsAdapter(obls.MatrixBLS())
    prob = NumpyProblem2(Lor, JLor, Z0_LOR.copy(), PAR_LOR, 6)
    nopt = P.NewtonPar(tol=1e-12, max_iterations=25, linsolver=ls)
    cp = P.ContinuationPar(p_min=-1.5, p_max=3.0, ds=-0.001, dsmax=0.025, dsmin=1e-4, max_steps=252, newton_options=nopt)
    pts = []
    rows, _ = P.continuation(prob, P.PALC(tangent="bordered", bls=bls), cp, normC=P.norminf,
                             callback=lambda s: pts.append((s.z_u.copy(), s.z_p, s.tau_u.copy())) or True)
    i = int(np.argmin([r["param"] for r in rows]))               # the turning point of the branch in F
    x0, p0, tau = pts[i]
    t = tau / np.linalg.norm(tau)
    f = C2.newton_fold(prob, x0, p0, t, t, nopt, bls, symmetric=False)
    assert f.converged and 1.546648372620807 <= f.p + 1e-9 and f.p - 1e-9 <= 1.5466483727182652   # inside the reference's interval for it
    cpf = P.ContinuationPar(p_min=-0.1, p_max=3.2, ds=-0.001, dsmax=0.005, dsmin=1e-5, max_steps=60, n_inversion=8, max_bisection_steps=25, nev=4,
                            newton_options=nopt)
    curve = C2.continuation_fold(prob, f.u, f.p, 5, t, t, cpf, bls, symmetric=False, normC=P.norminf, detect_event=1, eigsolver=_dense_eig)
    sp = curve.specialpoint
    assert [s.type for s in sp] == ["bt", "zh", "zh", "bt"], [(s.type, s.param) for s in sp]
    assert _isapprox(sp[0].param, 0.02058724, 1e-5) and abs(sp[1].param - 0.00004983) < 1e-8
    assert _isapprox(sp[2].param, -0.00045281, 1e-5) and _isapprox(sp[3].param, -0.02135893, 1e-5)


def _complex_bordered_dense(Jc, a, b, shift):
    """MatrixBLS on the complex bordered system [Jc + shift, a; b^H, 0] [v; sigma] = [0; 1] (bdlinsolver = MatrixBLS(), lorenz84.jl:95)"""
    n = len(a)
    M = np.zeros((n + 1, n + 1), dtype=complex)
    M[:n, :n] = Jc.M + shift * np.eye(n)
    M[:n, n] = a
    M[n, :n] = np.conj(b)
    r = np.zeros(n + 1, dtype=complex)
    r[n] = 1
    sol = np.linalg.solve(M, r)
    return sol[:n], sol[n], True, 1


def test_special_points_of_the_lorenz84_hopf_curve():
    """lorenz84.jl:88-99: hp_codim2_test = continuation(br, 2, (@optic _.T), ContinuationPar(opts_br, ds = -0.001, dsmax = 0.02, dsmin = 1e-4,
    n_inversion = 6, max_steps = 100); normC = norminf, detect_codim2_bifurcation = 2, update_minaug_every_step = 1, start_with_eigen =
    true, bothside = true, bdlinsolver = MatrixBLS()) on the branch computed with PALC(tangent = Bordered()):
        specialpoint[2].param ≈ +0.02627393 (rtol 1e-5),   specialpoint[3].param ≈ -0.02627430 (atol 1e-8)
    -- two Hopf-Hopf points (a second complex pair crosses the axis) located by the event bisection on the number of unstable
    eigenvalues; the whole Hopf chain (complex bordered systems, HopfLinearSolverMinAug, BorderingBLS over it, Bordered tangent)."""
    from tests.test_host_logic_cpu import _dense_cls, _dense_ls2
    from tests.test_codim2_curves_cpu import DenseComplexProblem2
    bk = g.load_package()
    P, E, C2 = bk.palc, bk.events, bk.codim2
    ls, bls = krylov.DefaultLS(), BlsAdapter(obls.MatrixBLS())
    prob = NumpyProblem2(Lor, JLor, Z0_LOR.copy(), PAR_LOR, 6)
    nopt = P.NewtonPar(tol=1e-12, max_iterations=25, linsolver=ls)
    cp = P.ContinuationPar(p_min=-1.5, p_max=3.0, ds=-0.001, dsmax=0.025, dsmin=1e-4, max_steps=252, n_inversion=6, max_bisection_steps=25, nev=4,
                           detect_bifurcation=3, newton_options=P.NewtonPar(tol=1e-12, max_iterations=25, linsolver=ls, eigsolver=_dense_eig))
    br = E.continuation(prob, P.PALC(tangent="bordered", bls=bls), cp, normC=P.norminf)
    h2 = [s for s in br.specialpoint if s.type == "hopf"][-1]    # br.specialpoint[2]: the Hopf point at F = 2.8599
    assert _isapprox(h2.interval[1], 2.859897757930758)
    Jh = JLor(h2.x, prob._par(h2.param))
    vals, vecs = np.linalg.eig(Jh)
    kk = int(np.argmax(vals.imag))
    valt, vect = np.linalg.eig(Jh.T)
    kt = int(np.argmin(valt.imag))
    cprob = DenseComplexProblem2(prob)
    hp = C2.newton_hopf(prob, cprob, h2.x, h2.param, vals[kk].imag, vecs[:, kk], vect[:, kt], nopt, _dense_ls2, _dense_cls, cbls=_complex_bordered_dense)
    assert hp.converged and h2.interval[0] - 1e-6 < hp.p < h2.interval[1] + 1e-6
    cph = P.ContinuationPar(p_min=-1.5, p_max=3.0, ds=-0.001, dsmax=0.02, dsmin=1e-4, max_steps=100, n_inversion=6, max_bisection_steps=25, nev=4,
                            newton_options=nopt)
    curve = C2.continuation_hopf(prob, cprob, hp.u, hp.p, hp.omega, 5, vecs[:, kk], vect[:, kt], cph, _dense_ls2, _dense_cls,
                                 alg=P.PALC(tangent="bordered"), normC=P.norminf, cbls=_complex_bordered_dense, detect_event=2, eigsolver=_dense_eig)
    sp = curve.specialpoint
    assert [s.type for s in sp] == ["hh", "hh"] and all(s.status == "converged" for s in sp), [(s.type, s.param, s.status) for s in sp]
    assert _isapprox(sp[0].param, 0.02627393, 1e-5) and abs(sp[0].param - 0.02627393) < 1e-8
    assert abs(sp[1].param - (-0.02627430)) < 1e-8
    assert prob.params == PAR_LOR and not curve.stopped_at_bt
    # at the located points J has two pairs on the imaginary axis (Hopf-Hopf): the Hopf pair +- i omega and a second one within the interval
    for s in sp:
        q = list(PAR_LOR)
        q[6], q[5] = s.p1, s.param
        ev = np.linalg.eigvals(JLor(s.x, q))
        assert np.sum(np.abs(ev.real) < 1e-5) == 4, ev
