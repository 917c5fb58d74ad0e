import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
bk = g.load_package()
n = int(sys.argv[1])
ctx = bk.Context(bk.BK_SH2D, (n, n), bench.domain(n), krylov_m=100, params=bench.PAR)
u = ctx.to_device(bench.sol0(n)); rhs = ctx.to_device(np.random.default_rng(1234).standard_normal(n * n))
J = ctx.jacobian(u); ls = bk.GMRESB200(reltol=1e-30, restart=100, maxiter=100)
ctx.set_timing(True)
for _ in range(3): ls(J, rhs)
s = ctx.stats()
print(f"n={n} E={os.environ.get('BK2_E')} fused_ms={s['last_fused_ms']:.3f} GBps={s['last_fused_bytes'] / 1e6 / s['last_fused_ms']:.0f}")
