import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
bk = g.load_package()
nx = int(os.environ.get("NX", "512")); tag = os.environ.get("TAG", "")
ny = nx; M = 30; L = (np.pi, np.pi / 2)
hx, hy = 2 * L[0] / nx, 2 * L[1] / ny
lam1 = -(2 - 2 * np.cos(np.pi / (nx + 1))) / hx**2 - (2 - 2 * np.cos(np.pi / (ny + 1))) / hy**2
pars = (-lam1 - 0.01, 0.1, 1.0, -1.0, 1.0)
ctx = bk.Context(bk.BK_POTRAP_CGL2D, (nx, ny, M), L, krylov_m=60, params=pars)
i = np.arange(1, nx + 1); j = np.arange(1, ny + 1)
phi11 = (np.sin(np.pi * i / (nx + 1))[None, :] * np.sin(np.pi * j / (ny + 1))[:, None]).reshape(-1)
xs = np.concatenate([np.concatenate([phi11 * np.cos(2 * np.pi * k / M), phi11 * np.sin(2 * np.pi * k / M)]) for k in range(M)] + [np.array([7.1])])
N = ctx.N
rng = np.random.default_rng(0)
phi = rng.standard_normal(N - 1); phi /= np.linalg.norm(phi)
ctx.potrap_set_section(phi, np.zeros(N - 1))
x = ctx.to_device(xs)
J = ctx.jacobian(x)
rhs = ctx.residual(x); bn = np.linalg.norm(rhs.numpy())
def report(name, ls, a0=0.0, a1=1.0):
    sol, ok, it = ls(J, rhs, a0=a0, a1=a1)
    tr = ctx.jvp(sol, a0=a0, a1=a1).numpy() - rhs.numpy()
    print(tag, nx, name, "ok", ok, "its", it, "est %.3e true %.3e" % (ls.last_resnorm / bn, np.linalg.norm(tr) / bn), flush=True)
    return sol
# (C) no preconditioner, strongly shifted operator: tests the Krylov kernels alone at this N
report("shifted-noprec-cgs", bk.GMRESB200(reltol=1e-8, restart=50, maxiter=50, orth="cgs"), a0=3e5)
ctx.precond_setup(bk.BK_PC_POTRAP_CIRC, 7.1)
report("right-cgs2", bk.GMRESB200(reltol=1e-6, restart=50, maxiter=50, Pr=True, orth="cgs2"))
report("right-cgs", bk.GMRESB200(reltol=1e-6, restart=50, maxiter=50, Pr=True, orth="cgs"))
# left preconditioning: the estimate is the preconditioned residual
ls = bk.GMRESB200(reltol=1e-6, restart=50, maxiter=50, Pl=True, orth="cgs2")
sol, ok, it = ls(J, rhs)
tr = ctx.precond_apply(ctx.to_device(rhs.numpy() - J(sol).numpy())).numpy()
pb = np.linalg.norm(ctx.precond_apply(rhs).numpy())
print(tag, nx, "left-cgs2 ok", ok, "its", it, "est(prec) %.3e true(prec) %.3e" % (ls.last_resnorm / pb, np.linalg.norm(tr) / pb), flush=True)
