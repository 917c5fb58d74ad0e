"""Kernel micro-benchmark: GMRES(m) Arnoldi cycle on SH2d (no preconditioner, fixed m iterations),
CUDA-event timed fused kernels -> achieved algorithmic GB/s  (bytes = 8N(2j+4) per step)."""
import json
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g


def run(n, m, fused=True, orth="cgs", reps=3):
    bk = g.load_package()
    lx, ly = 8 * np.pi, 4 * np.pi / np.sqrt(3)
    ctx = bk.Context(bk.BK_SH2D, (n, n), (lx, ly), krylov_m=m, params=(-0.1, 1.3))
    rng = np.random.default_rng(1234)
    X = -lx + 2 * lx / n * np.arange(n)
    Y = -ly + 2 * ly / n * np.arange(n)
    u = (np.cos(X)[None, :] + np.cos(X / 2)[None, :] * np.cos(np.sqrt(3) * Y / 2)[:, None]).reshape(-1)
    ud = ctx.to_device(u)
    rhs = ctx.to_device(rng.standard_normal(n * n))
    J = ctx.jacobian(ud)
    ls = bk.GMRESB200(reltol=1e-30, restart=m, maxiter=m, fused=fused, orth=orth)
    ctx.set_timing(True)
    out = []
    for r in range(reps):
        t0 = time.perf_counter()
        x, ok, it = ls(J, rhs)
        ctx.sync()
        wall = time.perf_counter() - t0
        s = ctx.stats()
        out.append(dict(n=n, m=m, iters=it, fused=fused, orth=orth, wall_ms=wall * 1e3, fused_ms=s["last_fused_ms"],
                        fused_GB=s["last_fused_bytes"] / 1e9,
                        fused_GBps=s["last_fused_bytes"] / 1e9 / (s["last_fused_ms"] * 1e-3) if s["last_fused_ms"] else None,
                        wall_GBps=s["last_fused_bytes"] / 1e9 / wall))
    return out


if __name__ == "__main__":
    for n in (512, 1024):
        for fused in (True, False):
            for row in run(n, 100, fused=fused)[1:]:
                print(json.dumps(row))
    for row in run(1024, 100, fused=True, orth="cgs2")[1:]:
        print(json.dumps(row))
