import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
bk = g.load_package(); P = bk.palc
for nx in (256, 512):
    ny = nx; M = 30; L = (np.pi, np.pi / 2); n = nx * ny
    hx, hy = 2 * L[0] / nx, 2 * L[1] / ny
    lam1 = -(2 - 2 * np.cos(np.pi / (nx + 1))) / hx**2 - (2 - 2 * np.cos(np.pi / (ny + 1))) / hy**2
    r = -lam1 - 0.01; pars = (r, 0.1, 1.0, -1.0, 1.0)
    ctx = bk.Context(bk.BK_POTRAP_CGL2D, (nx, ny, M), L, krylov_m=60, params=pars)
    i = np.arange(1, nx + 1); j = np.arange(1, ny + 1)
    phi11 = (np.sin(np.pi * i / (nx + 1))[None, :] * np.sin(np.pi * j / (ny + 1))[:, None]).reshape(-1)
    xs = np.concatenate([np.concatenate([phi11 * np.cos(2 * np.pi * k / M), phi11 * np.sin(2 * np.pi * k / M)]) for k in range(M)] + [np.array([7.1])])
    N = ctx.N
    rng = np.random.default_rng(0)
    phi = rng.standard_normal(N - 1); phi /= np.linalg.norm(phi)
    ctx.potrap_set_section(phi, np.zeros(N - 1))
    v = np.concatenate([np.concatenate([rng.standard_normal() * phi11 * np.cos(2 * np.pi * k / M + 0.3), phi11 * np.sin(4 * np.pi * k / M)]) for k in range(M)] + [np.array([0.37])])
    x = ctx.to_device(xs); vd = ctx.to_device(v)
    J = ctx.jacobian(x)
    Jv = J(vd).numpy()
    for eps in (1e-4, 1e-6):
        fp = ctx.residual(ctx.to_device(xs + eps * v)).numpy(); fm = ctx.residual(ctx.to_device(xs - eps * v)).numpy()
        fd = (fp - fm) / (2 * eps)
        print(nx, "eps", eps, "rel FD-vs-JVP", np.linalg.norm(fd - Jv) / np.linalg.norm(Jv), "max abs diff", np.abs(fd - Jv).max(), "at", int(np.argmax(np.abs(fd - Jv))), "N", N, flush=True)
    # GMRES true residual with the preconditioner
    ctx.precond_setup(bk.BK_PC_POTRAP_CIRC, 7.1)
    rhs = ctx.residual(x)
    for orth in ("cgs", "cgs2"):
        ls = bk.GMRESB200(reltol=1e-6, restart=50, maxiter=50, Pr=True, orth=orth)
        sol, ok, it = ls(J, rhs)
        tr = (J(sol).numpy() - rhs.numpy())
        print(nx, orth, "ok", ok, "its", it, "estimated", ls.last_resnorm / np.linalg.norm(rhs.numpy()), "true rel res", np.linalg.norm(tr) / np.linalg.norm(rhs.numpy()), flush=True)
    del ctx
