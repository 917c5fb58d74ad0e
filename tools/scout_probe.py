"""1-GPU experiment behind bench.py --gpus N: how cheap can the replicated scout be, and what N-GPU time does the
cost partition predict?  Runs the window once at full accuracy (per-step arclength and device time), then several scout settings
(time, points, rejected steps), and for each the predicted max-over-ranks time for N = 2, 4, 8 from the measured per-step costs.
   python tools/scout_probe.py [S_total=1.0] [grid=1024]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench

bk = g.load_package(); P = bk.palc; S = bk.segments
s_total = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx, ls, u_front = bench.gpu_setup(bk, n, 0)
mkprob = lambda u, p: P.BifurcationProblemB200(ctx, u, [p, bench.PAR[1]], lens=0)
alg = P.PALC(bls=bk.MatrixFreeBLSB200(ls))
cp = P.ContinuationPar(max_steps=100000, newton_options=P.NewtonPar(tol=1e-9, max_iterations=15, linsolver=ls), **bench.CONT)

# ---- the window at full accuracy, one GPU
stamps = []
def tick(st, s):
    ctx.sync(); stamps.append((s, time.perf_counter())); return True
trk = S.ArcTracker(P.V, alg.theta, 0.0, s_total, tick)
ctx.sync(); t0 = time.perf_counter()
rows, st = P.continuation(mkprob(u_front, bench.PAR[0]), alg, cp, normC=P.norminf, callback=trk)
ctx.sync(); T1 = time.perf_counter() - stamps[0][1]
sig = np.array([a for a, _ in stamps]); tt = np.array([b for _, b in stamps]) - stamps[0][1]
print(json.dumps({"fine": {"steps": len(rows) - 1, "seconds": T1, "rejected": st.nfail, "lambda_end": rows[-1]["param"],
                           "newton_its": st.work_newton, "linear_its": st.work_linear}}), flush=True)
ref = np.array([[r["param"], r["x"]] for r in rows])

def predict(sc, world):
    b = S.partition_by_cost(sc.cost, world)
    out = []
    for r in range(len(b) - 1):
        a0, a1 = sc.sigma[b[r]], min(sc.sigma[b[r + 1]], s_total)
        out.append(float(np.interp(a1, sig, tt) - np.interp(a0, sig, tt)))
    return b, out

for name, kw in [("ds x3, tol 1e-5, gmres 1e-2", dict(f=3, tol=1e-5, rt=1e-2, mi=8)),
                 ("ds x4, tol 1e-4, gmres 1e-2", dict(f=4, tol=1e-4, rt=1e-2, mi=8)),
                 ("ds x6, tol 1e-4, gmres 3e-2", dict(f=6, tol=1e-4, rt=3e-2, mi=8)),
                 ("ds x3, tol 1e-6, gmres 1e-3", dict(f=3, tol=1e-6, rt=1e-3, mi=10))]:
    ls_s = bk.GMRESB200(N=n * n, Pr=True, reltol=kw["rt"], restart=100, maxiter=100)
    cps = P.ContinuationPar(max_steps=100000, newton_options=P.NewtonPar(tol=kw["tol"], max_iterations=kw["mi"], linsolver=ls_s),
                            dsmin=bench.CONT["dsmin"], dsmax=kw["f"] * bench.CONT["dsmax"], ds=kw["f"] * bench.CONT["ds"],
                            p_min=bench.CONT["p_min"], p_max=bench.CONT["p_max"])
    ctx.sync(); t0 = time.perf_counter()
    try:
        sc = S.run_scout(P, mkprob(u_front, bench.PAR[0]), P.PALC(bls=bk.MatrixFreeBLSB200(ls_s)), cps, P.norminf, s_total,
                         lambda v: v.copy(), margin=2 * kw["f"] * bench.CONT["dsmax"])
    except Exception as e:
        print(json.dumps({"scout": name, "error": str(e)[:200]}), flush=True); continue
    ctx.sync(); Ts = time.perf_counter() - t0
    scr = np.array([[r["param"], r["x"]] for r in sc.rows])
    res = {"scout": name, "seconds": Ts, "points": len(sc.points), "frac_of_T1": Ts / T1, "off_curve": S.curve_distance(scr, ref),
           "sigma_end": sc.sigma[-1]}
    for world in (2, 4, 8):
        b, ch = predict(sc, world)
        tN = Ts + max(ch) + 0.03  # + ~2 steps of seed correction
        res[f"N{world}"] = {"chunks_s": [round(c, 3) for c in ch], "pred_T": round(tN, 3), "pred_eff": round(T1 / (world * tN), 3)}
    print(json.dumps(res), flush=True)
    # one real chunk (the last of 4): does the full-accuracy run from loose seeds land on the same curve?
    b = S.partition_by_cost(sc.cost, 4)
    ctx.sync(); t0 = time.perf_counter()
    crow, cst, ctrk = S.run_chunk(P, mkprob, alg, P.ContinuationPar(max_steps=100000, newton_options=cp.newton_options, dsmin=cp.dsmin, dsmax=cp.dsmax,
                                  ds=np.sign(cp.ds) * cp.dsmax, p_min=cp.p_min, p_max=cp.p_max), P.norminf, sc, b[2], b[3], s_total, 2, False)
    ctx.sync(); tc = time.perf_counter() - t0
    cr = np.array([[r["param"], r["x"]] for r in crow])
    print(json.dumps({"chunk 3 of 4": {"steps": len(crow), "seconds": tc, "rejected": cst.nfail, "first_itnewton": crow[0]["itnewton"] if crow else None,
                                      "distance_to_1gpu_curve": S.curve_distance(cr, ref) if len(cr) else None}}), flush=True)
