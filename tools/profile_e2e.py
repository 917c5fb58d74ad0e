import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
import bench
bk = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx, ls, u_front = bench.gpu_setup(bk, n, 0)
ctx.pin_host = True
bk.palc.V.host_alloc = ctx.pinned_empty
uh = ctx.pinned_array(u_front.numpy())
pr = cProfile.Profile()
pr.enable()
rows, ms, d, st = bench.gpu_run(bk, ctx, ls, uh, bench.PAR[0], 5, 1, torch, timing=False, flush=None)
pr.disable()
print("ms per step", ms)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
