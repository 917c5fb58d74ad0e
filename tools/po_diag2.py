import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import precond as oprecond
bk = g.load_package()
for nx in (256, 512):
    ny = nx; M = 30; L = (np.pi, np.pi / 2)
    pars = (1.2, 0.1, 1.0, -1.0, 1.0)
    ctx = bk.Context(bk.BK_POTRAP_CGL2D, (nx, ny, M), L, krylov_m=8, params=pars)
    N = ctx.N
    ctx.precond_setup(bk.BK_PC_POTRAP_CIRC, 7.1)
    v = np.random.default_rng(0).standard_normal(N)
    vd = ctx.to_device(v)
    a = ctx.precond_apply(vd).numpy(); b = ctx.precond_apply(vd).numpy()
    Po = oprecond.potrap_circulant_precond(nx, ny, *L, M, 7.1, 1.2, 1.0, workers=32)
    ref = Po(v)
    print(nx, "deterministic", np.array_equal(a, b), "rel err vs oracle", np.linalg.norm(a - ref) / np.linalg.norm(ref),
          "max abs", np.abs(a - ref).max(), "argmax", int(np.argmax(np.abs(a - ref))), "N", N, flush=True)
    # per-slice error
    Ns = 2 * nx * ny
    e = np.array([np.linalg.norm(a[s * Ns:(s + 1) * Ns] - ref[s * Ns:(s + 1) * Ns]) / np.linalg.norm(ref[s * Ns:(s + 1) * Ns]) for s in range(M)])
    print("per-slice rel err:", np.array2string(e, precision=1), flush=True)
    del ctx
