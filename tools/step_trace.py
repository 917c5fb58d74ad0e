"""Per-step trace of the bench workload: wall ms, itnewton, itlinear, rejected attempts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
bk = g.load_package(); P = bk.palc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
bench.BRANCH['kind'] = sys.argv[4] if len(sys.argv) > 4 else 'front'
ctx, ls, u_front = bench.gpu_setup(bk, n, 0)
cp = P.ContinuationPar(max_steps=steps, newton_options=P.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ls), **bench.CONT)
prob = P.BifurcationProblemB200(ctx, u_front, bench.PAR, lens=0)
t = [time.perf_counter()]; last = {"f": 0, "wl": 0, "wn": 0}
def cb(st):
    ctx.sync(); now = time.perf_counter()
    print(f"step {st.step:3d} p={st.z_p:+.5f} ds={st.ds:+.2e} itn={st.itnewton} itl={st.itlinear:4d} work_lin={st.work_linear - last['wl']:4d} work_newton={st.work_newton - last['wn']:2d} fails={st.nfail - last['f']} ms={(now - t[0]) * 1e3:7.2f}", flush=True)
    last.update(f=st.nfail, wl=st.work_linear, wn=st.work_newton); t[0] = time.perf_counter()
    return True
kind = sys.argv[3] if len(sys.argv) > 3 else "bordering"
blsobj = bk.MatrixFreeBLSB200(ls) if kind == "matrixfree" else bk.BorderingBLSB200(ls, check_precision=False)
print("bls:", kind)
P.continuation(prob, P.PALC(bls=blsobj), cp, normC=P.norminf, callback=cb)
