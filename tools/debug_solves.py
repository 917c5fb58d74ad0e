"""Per-solve accounting: iterations used vs Arnoldi steps launched, wall vs device time."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench

bk = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx, ls, u_front = bench.gpu_setup(bk, n, 0)
P = bk.palc
orig = bk.GMRESB200.__call__
log = []

def wrapped(self, J, rhs, rhs2=None, a0=0.0, a1=1.0):
    if rhs2 is not None:
        return orig(self, J, rhs, rhs2, a0, a1)
    ctx.sync()
    s0 = ctx.stats(); t0 = time.perf_counter()
    out = orig(self, J, rhs, None, a0, a1)
    ctx.sync()
    t1 = time.perf_counter(); s1 = ctx.stats()
    log.append((out[2], (s1["total_fused_launches"] - s0["total_fused_launches"]) // 2, s1["kernel_launches"] - s0["kernel_launches"], (t1 - t0) * 1e3, s1["last_fused_ms"]))
    return out

bk.GMRESB200.__call__ = wrapped
ctx.set_timing(True)
import ctypes as C
# direct C-level bordering solve also goes through bk_gmres_dev (not the Python wrapper), so time the BLS call instead
orig_bls = bk.BorderingBLSB200.__call__
def wrapped_bls(self, *a, **k):
    ctx.sync(); s0 = ctx.stats(); t0 = time.perf_counter()
    out = orig_bls(self, *a, **k)
    ctx.sync(); t1 = time.perf_counter(); s1 = ctx.stats()
    log.append(("bls", out[3], (s1["total_fused_launches"] - s0["total_fused_launches"]) // 2, s1["kernel_launches"] - s0["kernel_launches"], round((t1 - t0) * 1e3, 3)))
    return out
bk.BorderingBLSB200.__call__ = wrapped_bls
cp = P.ContinuationPar(max_steps=3, newton_options=P.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ls), **bench.CONT)
prob = P.BifurcationProblemB200(ctx, u_front, bench.PAR, lens=0)
t0 = time.perf_counter()
rows, st = P.continuation(prob, P.PALC(bls=bk.BorderingBLSB200(ls, check_precision=False)), cp, normC=P.norminf)
ctx.sync()
print("total wall ms", (time.perf_counter() - t0) * 1e3)
for l in log: print(l)
for r in rows: print(r)
