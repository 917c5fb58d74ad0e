"""Newton on the Trapeze periodic-orbit functional of cGL2d (config 4) with the matrix-free solver + circulant preconditioner."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
bk = g.load_package(); P = bk.palc
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = 30
L = (np.pi, np.pi / 2)
n = nx * ny
hx, hy = 2 * L[0] / nx, 2 * L[1] / ny
lam1 = -(2 - 2 * np.cos(np.pi / (nx + 1))) / hx**2 - (2 - 2 * np.cos(np.pi / (ny + 1))) / hy**2
r_hopf = -lam1
i = np.arange(1, nx + 1); j = np.arange(1, ny + 1)
phi11 = (np.sin(np.pi * i / (nx + 1))[None, :] * np.sin(np.pi * j / (ny + 1))[:, None]).reshape(-1)
for dr, amp in ((-0.01, 0.6), (-0.01, 1.0), (-0.01, 1.4), (0.05, 0.3), (-0.1, 1.0)):
    r = r_hopf + dr
    ctx = bk.Context(bk.BK_POTRAP_CGL2D, (nx, ny, M), L, krylov_m=60, params=(r, 0.1, 1.0, -1.0, 1.0))
    xs = np.concatenate([np.concatenate([amp * phi11 * np.cos(2 * np.pi * k / M), amp * phi11 * np.sin(2 * np.pi * k / M)]) for k in range(M)] + [np.array([2 * np.pi])])
    N = ctx.N
    x = ctx.to_device(xs)
    f1 = ctx.residual(x).numpy()[: 2 * n]
    phi = np.zeros(N - 1); phi[: 2 * n] = f1 / np.linalg.norm(f1)
    ctx.potrap_set_section(phi, np.zeros(N - 1))
    ctx.precond_setup(bk.BK_PC_POTRAP_CIRC, 2 * np.pi)
    ls = bk.GMRESB200(reltol=1e-4, restart=60, maxiter=60, Pr=True)
    prob = P.BifurcationProblemB200(ctx, x, (r, 0.1, 1.0, -1.0, 1.0), lens=0)
    ctx.sync(); t0 = time.perf_counter()
    sol = P.newton(prob, x, r, P.NewtonPar(tol=1e-8, max_iterations=15, linsolver=ls), P.norminf)
    ctx.sync(); dt = time.perf_counter() - t0
    u = sol.u.numpy()
    print(json.dumps({"grid": nx, "dr": dr, "amp0": amp, "converged": sol.converged, "newton_its": sol.itnewton, "linear_its": sol.itlineartot,
                      "seconds": dt, "T": float(u[-1]), "max_abs_u": float(np.max(np.abs(u[:-1]))), "residuals": [float(f"{q:.3e}") for q in sol.residuals[-4:]]}), flush=True)
    del ctx
