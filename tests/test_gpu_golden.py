"""CUDA path (through the C ABI) vs the committed golden fixtures -- no oracle code on this path."""
import os

import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_kernels_vs_golden():
    bk = g.load_package()
    c = bk.Context(bk.BK_SH2D, (40, 24), (LX, LY), krylov_m=80, params=(-0.1, 1.3))
    assert _rel(c.residual(G["sh2d_u"]), G["sh2d_F"]) < 1e-12
    J = c.jacobian(G["sh2d_u"])
    assert _rel(J(G["sh2d_v"]), G["sh2d_Jv"]) < 1e-12
    c.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    assert _rel(c.precond_apply(G["sh2d_v"]), G["sh2d_Pinv_v"]) < 1e-11
    x, ok, it = bk.GMRESB200(reltol=1e-10, restart=80, maxiter=80)(J, G["sh2d_v"], a0=30.0, a1=-1.0)
    assert ok and abs(it - int(G["sh2d_gmres_iters"][0])) <= 2 and _rel(x, G["sh2d_gmres_x"]) < 1e-8
    c3 = bk.Context(bk.BK_SH3D, (12, 10, 8), (2 * np.pi, 2 * np.pi, 1.5 * np.pi), krylov_m=4, params=(0.1, 1.2))
    assert _rel(c3.residual(G["sh3d_u"]), G["sh3d_F"]) < 1e-12 and _rel(c3.jacobian(G["sh3d_u"])(G["sh3d_v"]), G["sh3d_Jv"]) < 1e-12
    cc = bk.Context(bk.BK_CHAN, (101,), (1.0,), krylov_m=4, params=(3.3, 0.01))
    assert _rel(cc.residual(G["chan_x"]), G["chan_F"]) < 1e-13 and _rel(cc.jacobian(G["chan_x"])(G["chan_dx"]), G["chan_Jdx"]) < 1e-13
    cg = bk.Context(bk.BK_CGL2D, (10, 6), (np.pi, np.pi / 2), krylov_m=4, params=(1.2, 0.1, 1.0, -1.0, 1.0))
    assert _rel(cg.residual(G["cgl_u"]), G["cgl_F"]) < 1e-12 and _rel(cg.jacobian(G["cgl_u"])(G["cgl_du"]), G["cgl_Jdu"]) < 1e-12
    cp = bk.Context(bk.BK_POTRAP_CGL2D, (10, 6, 5), (np.pi, np.pi / 2), krylov_m=4, params=(1.2, 0.1, 1.0, -1.0, 1.0))
    cp.potrap_set_section(G["po_phi"], G["po_xpi"])
    assert _rel(cp.residual(G["po_x"]), G["po_res"]) < 1e-12 and _rel(cp.jacobian(G["po_x"])(G["po_dx"]), G["po_jvp"]) < 1e-12
