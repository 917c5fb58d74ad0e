"""Times the preconditioner application (CUDA events via torch on the library stream).
   python tools/bench_precond.py [n ...]        2-D n x n (default 512 1024 2048) and 128^3"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
import bench
bk = g.load_package()


def time_ctx(ctx, label, R=200):
    stream = torch.cuda.ExternalStream(ctx.lib.bk_stream(ctx.handle))
    x = ctx.to_device(np.random.default_rng(0).standard_normal(ctx.N)); y = ctx.zeros()
    for _ in range(5): ctx.precond_apply(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync(); e0.record(stream)
    for _ in range(R): ctx.precond_apply(x, y)
    e1.record(stream); ctx.sync()
    us = e0.elapsed_time(e1) / R * 1e3
    passes = 3 if ctx.kind == bk.BK_SH2D else 5
    print(f"precond apply {label}: {us:.1f} us  ({passes} kernels x 16 N bytes = {passes * 16 * ctx.N / 1e6:.1f} MB -> {passes * 16 * ctx.N / us / 1e3:.0f} GB/s)", flush=True)
    return us


sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048]
for n in sizes:
    ctx = bk.Context(bk.BK_SH2D, (n, n), bench.domain(n), krylov_m=2, params=bench.PAR)
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    time_ctx(ctx, f"SH2d {n}x{n}")
    del ctx
if not sys.argv[1:]:
    ctx = bk.Context(bk.BK_SH3D, (128, 128, 128), (np.pi,) * 3, krylov_m=2, params=(0.1, 1.2))
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    time_ctx(ctx, "SH3d 128^3", R=50)
    del ctx
    ctx = bk.Context(bk.BK_SH2D, (151 * 4, 100 * 4), bench.domain(256), krylov_m=2, params=bench.PAR)
    ctx.precond_setup(bk.BK_PC_SH_DCT, 1.0)
    time_ctx(ctx, "SH2d 604x400 (general kernel)", R=50)
    del ctx
    ctx = bk.Context(bk.BK_CGL2D, (512, 512), (np.pi, np.pi), krylov_m=2, params=(1.2, 0.1, 1.0, -1.0, 1.0))
    ctx.precond_setup(bk.BK_PC_CGL_DST, 1.0, -0.05)
    time_ctx(ctx, "cGL 512x512 DST-I (general kernel, L = 1026)", R=50)
