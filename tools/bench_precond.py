"""Times the preconditioner application and a short GMRES solve (CUDA events via torch on the library stream)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
import bench
bk = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = bk.Context(bk.BK_SH2D, (n, n), bench.domain(n), krylov_m=100, params=bench.PAR)
ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
stream = torch.cuda.ExternalStream(ctx.lib.bk_stream(ctx.handle))
x = ctx.to_device(np.random.default_rng(0).standard_normal(n * n)); y = ctx.zeros()
for _ in range(5): ctx.precond_apply(x, y)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx.sync(); e0.record(stream)
R = 200
for _ in range(R): ctx.precond_apply(x, y)
e1.record(stream); ctx.sync()
print(f"precond apply n={n}: {e0.elapsed_time(e1) / R * 1e3:.1f} us  (W={os.environ.get('BK_DCT_W')}, T={os.environ.get('BK_DCT_THREADS')})")
