import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
bk = g.load_package()
N = int(os.environ.get("NN", "3932161"))
ctx = bk.Context(bk.BK_CHAN, (N,), (1.0,), krylov_m=6, params=(3.3, 0.01))
rng = np.random.default_rng(1)
u = 0.1 * rng.standard_normal(N); b = rng.standard_normal(N)
J = ctx.jacobian(ctx.to_device(u)); rhs = ctx.to_device(b); bn = np.linalg.norm(b)
a0 = -40.0 * float(N - 1) ** 2 / 100.0
ls = bk.GMRESB200(reltol=1e-9, restart=6, maxiter=6, orth="cgs")
sol, ok, it = ls(J, rhs, a0=a0)
tr = ctx.jvp(sol, a0=a0).numpy() - b
print("SAN", N, "its", it, "est %.6e true %.6e" % (ls.last_resnorm / bn, np.linalg.norm(tr) / bn), flush=True)
