"""one-off diagnostics for the Hopf Newton on device (GPU call 14)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import problems
bk = g.load_package()
P = bk.palc
restart = int(sys.argv[1]) if len(sys.argv) > 1 else 300
orth = sys.argv[2] if len(sys.argv) > 2 else "cgs2"
Nx, Ny = 24, 12
gl = problems.GinzburgLandau2D(Nx, Ny, np.pi, np.pi / 2)
n = gl.N
rH, nu = gl.r_hopf(), gl.nu
par = (rH + 0.3, gl.mu, gl.nu, gl.c3, gl.c5)
rctx = bk.Context(bk.BK_CGL2D, (Nx, Ny), (np.pi, np.pi / 2), krylov_m=200, params=par)
cctx = bk.Context(bk.BK_CGL2D, (Nx, Ny), (np.pi, np.pi / 2), krylov_m=300, params=par, complex=True)
rctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)
cctx.precond_setup(bk.BK_PC_CGL_DST, -1.0, 1.0)
ls = bk.GMRESB200(reltol=1e-11, restart=200, maxiter=600, Pr=True, orth="cgs2")
base = bk.ComplexGMRESB200(reltol=1e-11, restart=restart, maxiter=900, Pr=True, orth=orth)
def cls(J, rhs, a0=0.0, a1=1.0):
    try:
        x, cv, it = base(J, rhs, a0=a0, a1=a1)
    except Exception as e:
        print("  complex solve FAILED a0", a0, "transpose", J.transpose, "|rhs|", np.linalg.norm(rhs), str(e)[:120], flush=True)
        raise
    print(f"  complex solve a0={a0} T={J.transpose} cv={cv} it={it} |x|={np.linalg.norm(x):.3e} res={base.last_resnorm:.2e}", flush=True)
    return x, cv, it
rng = np.random.default_rng(14)
phi = gl.phi11()
zeta = np.concatenate([phi, -1j * phi]) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
u0 = 1e-3 * rng.standard_normal(n)
prob = P.BifurcationProblemB200(rctx, u0, par, lens=0)
cprob = bk.codim2.ComplexProblemB200(cctx, par, lens=0)
ma = bk.codim2.HopfMinAug(prob, cprob, zeta, zeta, ls, cls)
x, p, om = u0.copy(), rH + 0.3, nu + 0.2
for it in range(8):
    F, sr, si = ma.residual(x, p, om)
    print("iter", it, "p-rH", p - rH, "om-nu", om - nu, "|F|", np.linalg.norm(F), "sigma", sr, si, flush=True)
    dX, dp, dom, _ = ma.solve(x, p, om, F, sr, si)
    x = x - dX
    p -= dp
    om -= dom
print("final", p - rH, om - nu, np.linalg.norm(x))
