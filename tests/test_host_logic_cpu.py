"""CPU tests of the product's HOST logic (bifurcationkit.jl_b200/palc.py, segments.py): the continuation loop is
driven with NumPy vectors and NumPy solvers (duck-typed problem; the oracle is only the checker), and the multi-GPU
row exchange runs on world_size-2 gloo."""
import os
import sys

import numpy as np
import pytest

import __graft_entry__ as g
from oracle import krylov, bls as obls, palc as opalc, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyProblem:
    """Duck-typed stand-in for BifurcationProblemB200 on host arrays."""

    def __init__(self, F, J, u0, p0, record=None):
        self.F_, self.J_, self.u0, self.p0 = F, J, u0, p0
        self.delta = float(np.sqrt(np.finfo(float).eps))
        self.record = record or (lambda x: float(np.linalg.norm(x)))

    def F(self, x, p, out=None):
        r = self.F_(x, p)
        if out is not None:
            out[...] = r
            return out
        return r

    def J(self, x, p):
        return self.J_(x, p)


class BlsAdapter:
    """product calling convention (dotscale=) -> oracle solver (dotp=)"""

    def __init__(self, inner):
        self.inner = inner

    def __call__(self, J, dR, dzu, dzp, R, n, xiu=1.0, xip=1.0, shift=None, dotscale=1.0):
        N = len(R)
        return self.inner(J, dR, dzu, dzp, R, n, xiu, xip, shift=shift, dotp=lambda a, b: dotscale * float(np.dot(a, b)),
                          apply_xiu=lambda row: row * dotscale)


def _fold_problem():
    F = lambda x, r: r + x - x**3
    J = lambda x, r: np.diag(1 - 3 * x**2)
    return F, J


@pytest.mark.parametrize("tangent", ["secant", "bordered"])
def test_product_host_loop_matches_oracle(tangent):
    """Same branch, row by row, as the oracle on test-cont-non-vector.jl:22-45 (param[end] == -1)."""
    bk = g.load_package()
    P = bk.palc
    F, J = _fold_problem()
    ls = krylov.DefaultLS()
    kw = dict(dsmin=0.001, dsmax=0.07, ds=-0.02, p_max=4.1, p_min=-1.0, max_steps=150)
    orows, _ = opalc.continuation(opalc.Problem(F=F, J=J, u0=np.array([0.8]), p0=1.0, record=lambda x: x[0]),
                                  opalc.PALC(tangent=tangent, bls=obls.MatrixBLS()),
                                  opalc.ContinuationPar(newton_options=opalc.NewtonPar(tol=1e-8, linsolver=ls), **kw))
    rows, st = P.continuation(NumpyProblem(F, J, np.array([0.8]), 1.0, record=lambda x: x[0]),
                              P.PALC(tangent=tangent, bls=BlsAdapter(obls.MatrixBLS())),
                              P.ContinuationPar(newton_options=P.NewtonPar(tol=1e-8, linsolver=ls), **kw))
    assert rows[-1]["param"] == -1.0
    assert len(rows) == len(orows)
    for r, o in zip(rows, orows):
        assert abs(r["param"] - o["param"]) < 1e-12 and abs(r["x"] - o["x"]) < 1e-12 and r["itnewton"] == o["itnewton"]
    assert st.nfail == 0 and st.work_newton == sum(r["itnewton"] for r in rows)


def test_step_size_control_and_rejection():
    bk = g.load_package()
    P = bk.palc
    cp = P.ContinuationPar(dsmin=1e-3, dsmax=0.1, a=0.5, newton_options=P.NewtonPar(max_iterations=10))
    assert np.isclose(P.step_size_control(0.01, True, 2, cp)[0], 0.01 * (1 + 0.5 * 0.8**2))
    assert P.step_size_control(-0.01, False, 10, cp) == (-0.005, False)
    assert P.step_size_control(1e-3, False, 10, cp)[1] is True
    # an initial guess that cannot converge raises like the reference (src/Continuation.jl:375-379)
    F = lambda x, p: np.exp(x) + 0 * p  # no root: Newton walks x <- x - 1 forever
    J = lambda x, p: np.diag(np.exp(x))
    prob = NumpyProblem(F, J, np.array([1.0]), 0.0)
    with pytest.raises(RuntimeError):
        P.continuation(prob, P.PALC(bls=BlsAdapter(obls.MatrixBLS())),
                       P.ContinuationPar(newton_options=P.NewtonPar(tol=1e-10, max_iterations=5, linsolver=krylov.DefaultLS())))


def test_two_point_start_continues_the_same_curve():
    """Segment seeding: a run started from two consecutive scout points traces the same curve."""
    bk = g.load_package()
    P = bk.palc
    n = 31
    beta = 0.01
    F = lambda x, a: problems.chan_F(x, a, beta)
    def Jd(x, a):
        E = np.eye(n)
        return np.column_stack([problems.chan_dF(x, E[:, k], a, beta) for k in range(n)])
    ls = krylov.DefaultLS()
    mk = lambda ms: P.ContinuationPar(dsmin=0.01, dsmax=0.2, ds=0.01, p_max=4.2, p_min=-1.0, max_steps=ms,
                                      newton_options=P.NewtonPar(tol=1e-10, max_iterations=10, linsolver=ls))
    alg = P.PALC(bls=BlsAdapter(obls.MatrixBLS()))
    stride = 6
    grab = bk.segments.SeedGrabber(1, stride, lambda v: v.copy())
    full, _ = P.continuation(NumpyProblem(F, Jd, problems.chan_sol0(n), 3.3), alg, mk(2 * stride))
    P.continuation(NumpyProblem(F, Jd, problems.chan_sol0(n), 3.3), alg, mk(stride + 1), callback=grab)
    u0, p0, u1, p1 = grab.pair()
    assert abs(p0 - full[stride]["param"]) < 1e-14
    resid = []
    seg, _ = P.continuation(NumpyProblem(F, Jd, u0, p0), alg, mk(stride), u1=u1, p1=p1,
                            callback=lambda st: resid.append(np.linalg.norm(F(st.z_u, st.z_p))) or True)
    assert seg[0]["param"] == p0 and len(seg) == stride + 1
    assert max(resid) < 1e-9                              # every segment point solves F(x, alpha) = 0
    fx = np.array([r["x"] for r in full]); fp = np.array([r["param"] for r in full])
    assert np.all(np.diff(fx) > 0) and np.all(np.diff([r["x"] for r in seg]) > 0)   # same direction along the branch
    for r in seg:                                          # and lies on the scout curve ||x||(alpha) (coarse interpolation)
        if fx[0] <= r["x"] <= fx[-1]:
            assert abs(np.interp(r["x"], fx, fp) - r["param"]) < 2e-2


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as gg
    bk = gg.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = [dict(param=-0.1 - 0.01 * (rank * 3 + i), x=1.0 + rank + 0.1 * i, itnewton=2, itlinear=40 + i) for i in range(4 - rank)]
    gathered = bk.segments.all_gather_rows(rows, 5, dist, torch, "cpu")
    merged = bk.segments.merge_branch(gathered)
    q.put((rank, gathered.shape, merged.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_rows_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == (2, 5, 4) and res[0][2] == res[1][2]      # every rank assembles the same branch
    merged = np.array(res[0][2])
    assert merged.shape == (4 + 3 - 1, 4)                           # rank-1 segment starts at rank-0's last point (-0.13): dropped once
    assert np.all(np.diff(merged[:, 0]) < 0)


# ------------------------------------------------------------------------------------------------ Floquet
class NumpyVF:
    """Duck-typed vector-field context (what floquet.py needs of a BK_CGL2D Context) on the NumPy cGL operator."""

    def __init__(self, gl):
        self.gl, self.u = gl, None

    def jacobian(self, u):
        self.u = np.array(u)
        return self

    def jvp(self, v, a0=0.0, a1=1.0):
        return a0 * np.asarray(v) + a1 * self.gl.dF(self.u, np.asarray(v))


def _dense_shifted_solver(gl):
    def ls(J, rhs, a0=0.0, a1=1.0):
        A = a0 * np.eye(gl.N) + a1 * np.column_stack([gl.dF(J.u, e) for e in np.eye(gl.N)])
        return np.linalg.solve(A, rhs), True, 1
    return ls


def test_floquet_host_logic_matches_oracle():
    """FloquetQaDB200 / ArnoldiLMB200 driven with NumPy vectors against oracle.floquet (Floquet.jl:285-316, 59-85)."""
    from oracle import floquet as ofl
    bk = g.load_package()
    nx, ny, M = 6, 5, 9
    gl = problems.GinzburgLandau2D(nx, ny, np.pi, np.pi / 2, r=1.3)
    N = gl.N
    ph = gl.phi11()
    x = np.concatenate([np.concatenate([0.5 * ph * np.cos(2 * np.pi * k / M), 0.5 * ph * np.sin(2 * np.pi * k / M)]) for k in range(M)] + [np.array([6.4])])
    jac = lambda u: np.column_stack([gl.dF(u, e) for e in np.eye(N)])
    mono = ofl.monodromy_dense(jac, x, M, N)
    fl = bk.floquet.FloquetQaDB200(NumpyVF(gl), _dense_shifted_solver(gl), M, eigsolver=bk.floquet.ArnoldiLMB200(krylovdim=40, tol=1e-10))
    v = np.random.default_rng(5).standard_normal(N)
    assert np.allclose(fl.monodromy(x, v), mono @ v, rtol=1e-10, atol=1e-12)
    assert fl.solves == M - 1
    sig, vecs, cv, info = fl(x, 4)
    ref, _ = ofl.floquet_exponents(np.linalg.eigvals(mono))
    ref4 = ref[np.argsort(-np.abs(np.exp(ref)))][:4]      # the 4 multipliers of largest modulus ...
    ref4 = ref4[np.argsort(-ref4.real, kind="stable")]      # ... as exponents by decreasing real part
    assert cv and np.allclose(np.sort(sig.real)[::-1], np.sort(ref4.real)[::-1], rtol=1e-7, atol=1e-9)
    assert np.all(np.diff(sig.real) <= 1e-12)
    mu = info["multipliers"]
    for k in range(4):                                       # Ritz pairs really are eigenpairs of the monodromy
        z = vecs[k][0] + 1j * vecs[k][1]
        assert np.linalg.norm(mono @ z - mu[k] * z) < 1e-7 * abs(mu[k]) * np.linalg.norm(z)
    sl = fl.extract_eigenvector(x, vecs[0][0])
    assert len(sl) == M and np.allclose(sl[M - 2], mono @ vecs[0][0], rtol=1e-9, atol=1e-12)
