"""Generates tests/golden/*.npz.  The reference (pure Julia) cannot be executed in this image, so these fixtures freeze
OUTPUTS OF THE ORACLE (oracle/, itself pinned against the reference's known-answer tests) on seeded inputs; they guard both
the oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) against drift.  The one vector that comes
from the reference itself is the 5x5 spectrum of test/linear_solvers/test_linear.jl:595-614 (stored verbatim).
Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import problems, krylov, bls, palc, precond, potrap  # noqa: E402


def main():
    rng = np.random.default_rng(20260923)
    LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)
    out = {}
    # P2: SH2d 40 x 24
    dims = (40, 24)
    sh = problems.SwiftHohenberg(dims, (LX, LY), l=-0.1, nu=1.3)
    u = problems.sh2d_sol0(*dims, LX, LY)
    v = rng.standard_normal(sh.N)
    out.update(sh2d_u=u, sh2d_v=v, sh2d_F=sh.F(u), sh2d_Jv=sh.dF(u, v), sh2d_Pinv_v=precond.dct_precond(dims, (LX, LY), 1.0)(v))
    x, ok, it = krylov.GMRESIterativeSolvers(reltol=1e-10, restart=80, maxiter=80)(lambda w: sh.dF(u, w), v, a0=30.0, a1=-1.0)
    out.update(sh2d_gmres_x=x, sh2d_gmres_iters=np.array([it]))
    # P3: SH3d 12 x 10 x 8
    d3, L3 = (12, 10, 8), (2 * np.pi, 2 * np.pi, 1.5 * np.pi)
    sh3 = problems.SwiftHohenberg(d3, L3, l=0.1, nu=1.2)
    u3 = problems.sh3d_sol0(*d3, *L3)
    v3 = rng.standard_normal(sh3.N)
    out.update(sh3d_u=u3, sh3d_v=v3, sh3d_F=sh3.F(u3), sh3d_Jv=sh3.dF(u3, v3))
    # P1: chan n = 101 + the plumbing branch (config 1 wiring, dense Jacobian solve)
    n = 101
    xc = problems.chan_sol0(n)
    dxc = rng.standard_normal(n)
    out.update(chan_x=xc, chan_dx=dxc, chan_F=problems.chan_F(xc, 3.3, 0.01), chan_Jdx=problems.chan_dF(xc, dxc, 3.3, 0.01))
    # P4 / P5: cGL 10 x 6, M = 5
    gl = problems.GinzburgLandau2D(10, 6, np.pi, np.pi / 2, r=1.2)
    ug, dug = 0.3 * rng.standard_normal(gl.N), rng.standard_normal(gl.N)
    out.update(cgl_u=ug, cgl_du=dug, cgl_F=gl.F(ug), cgl_Jdu=gl.dF(ug, dug))
    M = 5
    xpo = np.concatenate([0.3 * rng.standard_normal(gl.N * M), [6.2]])
    dxpo = np.concatenate([rng.standard_normal(gl.N * M), [0.4]])
    phi, xpi = rng.standard_normal(gl.N * M), rng.standard_normal(gl.N * M)
    tr = potrap.Trapeze(gl.F, gl.dF, phi, xpi, M, gl.N)
    out.update(po_x=xpo, po_dx=dxpo, po_phi=phi, po_xpi=xpi, po_res=tr.residual(xpo), po_jvp=tr.jvp(xpo, dxpo))
    # PALC branch: fold problem of test-cont-non-vector.jl:22-45
    prob = palc.Problem(F=lambda x, r: r + x - x**3, J=lambda x, r: np.diag(1 - 3 * x**2), u0=np.array([0.8]), p0=1.0,
                        record=lambda x: x[0])
    cp = palc.ContinuationPar(dsmin=0.001, dsmax=0.07, ds=-0.02, p_max=4.1, p_min=-1.0, max_steps=150,
                              newton_options=palc.NewtonPar(tol=1e-8, linsolver=krylov.DefaultLS()))
    rows, _ = palc.continuation(prob, palc.PALC(bls=bls.MatrixBLS()), cp)
    out["fold_branch"] = np.array([[r["param"], r["x"], r["itnewton"]] for r in rows])
    # reference-owned golden vector
    out["ref_5x5_eigvals_re"] = np.array([2.750124876460063, 0.2338099902832191, 0.2338099902832191, -0.42584697851325004, -0.42584697851325004])
    out["ref_5x5_eigvals_im"] = np.array([0.0, -0.3203002738693372, 0.3203002738693372, -0.17961985097997188, 0.17961985097997188])
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
