"""SASS evidence per kernel of libbk200.so (cuobjdump -sass): instruction count and the mnemonics that prove the Blackwell-native
paths (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, FENCE.VIEW.ASYNC = fence.proxy.async, DFMA/DADD/DMUL = fp64 pipe,
LDS/STS = shared memory, LDG/STG = global, ACQBULK / griddepcontrol = PDL).   python tools/sass_summary.py > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bifurcationkit.jl_b200", "libbk200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
cur, cnt = None, collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        op = m.group(2)
        cnt[cur][op.split(".")[0]] += 1
        if op.startswith("FENCE.VIEW.ASYNC"):
            cnt[cur]["FENCE.VIEW.ASYNC"] += 1
        if "ACQBULK" in op or "PREEXIT" in op or op.startswith("ACQ"):
            cnt[cur]["PDL"] += 1
KEYS = ["UBLKCP", "UBLKPF", "SYNCS", "FENCE.VIEW.ASYNC", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "LDL", "STL"]
print(f"# {os.path.relpath(so, ROOT)}: SASS mnemonic counts per kernel (sm_100a), from `cuobjdump -sass`")
print(f"{'instr':>7} " + " ".join(f"{k[:9]:>9}" for k in KEYS) + "  kernel")
tot = collections.Counter()
for k, c in sorted(cnt.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(v for kk, v in c.items() if kk not in ("FENCE.VIEW.ASYNC", "PDL"))
    name = re.sub(r"\(.*", "", dem(k))
    name = re.sub(r"^void ", "", name)
    print(f"{n:7d} " + " ".join(f"{c.get(kk, 0):9d}" for kk in KEYS) + f"  {name[:100]}")
    tot.update(c)
print("# totals: " + ", ".join(f"{k} {tot.get(k, 0)}" for k in KEYS) + f"; kernels {len(cnt)}")
