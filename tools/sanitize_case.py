"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel family runs once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
bk = g.load_package(); P = bk.palc
LX, LY = 8 * np.pi, 4 * np.pi / np.sqrt(3)
n = 256
ctx = bk.Context(bk.BK_SH2D, (n, 64), (LX, LY), krylov_m=30, params=(-0.1, 1.3))
ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
rng = np.random.default_rng(0)
u = ctx.to_device(rng.standard_normal(ctx.N) * 0.1); rhs = ctx.to_device(rng.standard_normal(ctx.N))
J = ctx.jacobian(u)
for kw in (dict(Pr=True), dict(Pl=True), dict(Pr=True, orth="cgs2"), dict()):
    x, ok, it = bk.GMRESB200(reltol=1e-6, restart=30, maxiter=30, **kw)(J, rhs)
a, b = ctx.to_device(rng.standard_normal(ctx.N)), ctx.to_device(rng.standard_normal(ctx.N))
ls = bk.GMRESB200(reltol=1e-6, restart=30, maxiter=30, Pr=True)
bk.MatrixFreeBLSB200(ls)(J, a, b, 0.9, rhs, 0.1, 0.5, 0.5, dotscale=1.0 / ctx.N)
bk.BorderingBLSB200(ls, check_precision=True, k=1)(J, a, b, 0.9, rhs, 0.1, 0.5, 0.5, dotscale=1.0 / ctx.N)
bk.ShiftInvertB200(0.1, ls, krylovdim=12, tol=1e-4, maxrestart=2)(J, 3)
c3 = bk.Context(bk.BK_SH3D, (32, 16, 16), (np.pi, np.pi, np.pi), krylov_m=10, params=(0.1, 1.2))
c3.precond_setup(bk.BK_PC_SH_DCT, 1.0)
u3 = c3.to_device(rng.standard_normal(c3.N) * 0.1)
bk.GMRESB200(reltol=1e-4, restart=10, maxiter=10, Pr=True)(c3.jacobian(u3), c3.to_device(rng.standard_normal(c3.N)))
print("sanitize case done", ok, it)
