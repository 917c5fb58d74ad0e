"""GMRES estimate-vs-true residual on a 1-D problem of arbitrary length (diagnostic for the TMA-ring kernels)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
bk = g.load_package()
N = int(os.environ["NN"]); tag = os.environ.get("TAG", "")
ctx = bk.Context(bk.BK_CHAN, (N,), (1.0,), krylov_m=40, params=(3.3, 0.01))
rng = np.random.default_rng(1)
u = 0.1 * rng.standard_normal(N); b = rng.standard_normal(N)
J = ctx.jacobian(ctx.to_device(u)); rhs = ctx.to_device(b); bn = np.linalg.norm(b)
# scale: J ~ (N-1)^2 * tridiag; a0 makes A = a0 I + J moderately conditioned (GMRES needs ~25-40 its)
a0 = -40.0 * float(N - 1) ** 2 / 100.0
for orth in ("cgs", "cgs2"):
    ls = bk.GMRESB200(reltol=1e-9, restart=40, maxiter=40, orth=orth)
    sol, ok, it = ls(J, rhs, a0=a0)
    tr = ctx.jvp(sol, a0=a0).numpy() - b
    print(tag, N, os.environ.get("BK2_E", "auto"), orth, "ok", ok, "its", it, "est %.6e true %.6e" % (ls.last_resnorm / bn, np.linalg.norm(tr) / bn), flush=True)
