"""Shared-memory bank-conflict model for the DCT kernels (bk_dct2.cuh): counts wavefronts per phase for one CTA.

Model: 32 banks x 4 B.  A warp access of 8 B per lane is served per half-warp, 16 B per lane per quarter-warp; the number of
wavefronts of a group is max over banks of the number of distinct 4-byte words requested in that bank.  ideal = one wavefront
per group.  Usage: python bank_model.py [logM=9] [logW=2] [threads=512]"""
import sys
import numpy as np

LOGM = int(sys.argv[1]) if len(sys.argv) > 1 else 9
LOGW = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 512
M, W = 1 << LOGM, 1 << LOGW
n = 2 * M
MP = 0  # set below


PADKIND = sys.argv[4] if len(sys.argv) > 4 else "A"
TWK = sys.argv[5] if len(sys.argv) > 5 else "strided"
REMAP = len(sys.argv) > 6 and sys.argv[6] == "remap"


def pad(i):
    if PADKIND == "A":
        return i + (i >> 3)
    if PADKIND == "B":
        return i + (i >> 3) + (i >> 6)
    if PADKIND == "C":
        return i + (i >> 3) + (i >> 6) + (i >> 9)
    if PADKIND == "D":   # XOR swizzle of the low 3 bits, no extra memory
        return i ^ ((i >> 3) & 7) ^ ((i >> 6) & 7)
    if PADKIND == "E":
        return i ^ ((i >> 3) & 7) ^ ((i >> 6) & 7) ^ ((i >> 9) & 7)
    raise SystemExit("pad kind?")


def _mp():
    return max(pad(i) for i in range(M)) + 1


def brev(v, bits):
    return int(format(v, f"0{bits}b")[::-1], 2) if bits else 0


def slot(line, pi, strided):
    return (pi << LOGW) + line if strided else line * MP + pi


def wavefronts(addrs, size):
    """addrs: byte addresses of the 32 lanes (None = inactive); size 8 or 16."""
    group = 16 if size == 8 else 8
    tot = ideal = 0
    for g0 in range(0, 32, group):
        lanes = [a for a in addrs[g0:g0 + group] if a is not None]
        if not lanes:
            continue
        banks = {}
        for a in lanes:
            for w in range(size // 4):
                word = a // 4 + w
                banks.setdefault(word % 32, set()).add(word)
        tot += max(len(v) for v in banks.values())
        ideal += 1
    return tot, ideal


def phase(name, nitems, accesses):
    """accesses(q) -> list of (byte address, size) for item q (same length for all items)."""
    tot = ideal = 0
    for q0 in range(0, nitems, T):            # one pass of the CTA over T items
        for w0 in range(q0, min(q0 + T, nitems), 32):
            per_lane = [accesses(q) if q < nitems else None for q in range(w0, w0 + 32)]
            nacc = len(next(p for p in per_lane if p is not None))
            for k in range(nacc):
                addrs = [p[k][0] if p is not None else None for p in per_lane]
                size = next(p for p in per_lane if p is not None)[k][1]
                t, i = wavefronts(addrs, size)
                tot += t
                ideal += i
    print(f"  {name:34s} wavefronts {tot:7d}  ideal {ideal:7d}  x{tot / max(ideal, 1):.2f}")
    return tot, ideal


def run(strided, layout=None):
    lay = strided if layout is None else layout
    print(f"M={M} W={W} threads={T} mapping={'line fastest' if strided else 'line-major'} layout={'line fastest' if lay else 'line-major'}")
    tt = ti = 0

    def load(q):
        line = (q & (W - 1)) if strided else (q >> LOGM)
        j = (q >> LOGW) if strided else (q & (M - 1))
        p0 = brev(j >> 1, LOGM)
        p1 = (M - 1) - p0
        return [(16 * slot(line, pad(p0), lay) + 8 * (j & 1), 8), (16 * slot(line, pad(p1), lay) + 8 * (1 - (j & 1)), 8)]

    a, b = phase("load (2 x STS.64)", M * W, load); tt += a; ti += b
    st = 0
    if LOGM & 1:
        def r2(q):
            line = (q & (W - 1)) if strided else (q >> (LOGM - 1))
            bf = (q >> LOGW) if strided else (q & ((M >> 1) - 1))
            p = pad(2 * bf)
            return [(16 * slot(line, p, lay), 16), (16 * slot(line, p + 1, lay), 16)] * 2
        a, b = phase("radix-2 pass (2 LDS.128 + 2 STS.128)", (M >> 1) * W, r2); tt += a; ti += b
        st = 1
    while st < LOGM:
        half = 1 << st

        def r4(q, st=st, half=half):
            line = (q & (W - 1)) if strided else (q >> (LOGM - 2))
            gi = (q >> LOGW) if strided else (q & ((M >> 2) - 1))
            if REMAP and not strided and 0 < st < 3:
                # within each block of 8 * half consecutive butterflies let the group index run fastest:
                # lanes 0..7 -> 8 consecutive groups at pos 0, next 8 lanes pos 1, ...
                blk, r = gi // (8 * half), gi % (8 * half)
                gi = blk * 8 * half + (r % 8) * half + r // 8
            pos = gi & (half - 1)
            i = ((gi >> st) << (st + 2)) + pos
            acc = [(16 * slot(line, pad(i + c * half), lay), 16) for c in range(4)]
            if TWK == "compact":   # per-pass tables w1[pos], w2[pos], contiguous in pos
                off = MP * W + 2 * (half - 1)   # any per-pass offset; tables of `half` entries each
                tw = [(16 * (off + pos), 16), (16 * (off + half + pos), 16)]
            else:
                tw = [(16 * (MP * W + (pos << (LOGM - st - 1))), 16), (16 * (MP * W + (pos << (LOGM - st - 2))), 16)]
            return tw + acc + acc
        a, b = phase(f"radix-4 pass st={st} (2 tw + 4 LDS + 4 STS)", (M >> 2) * W, r4); tt += a; ti += b
        st += 2

    def post(q):
        if q < (M // 2) * W:
            line = (q & (W - 1)) if strided else (q >> (LOGM - 1))
            k = (q >> LOGW) if strided else (q & (M // 2 - 1))
        else:
            line, k = q - (M // 2) * W, M // 2
        kc = (M - k) & (M - 1)
        return [(16 * slot(line, pad(k), lay), 16), (16 * slot(line, pad(kc), lay), 16)]
    a, b = phase("post (2 LDS.128)", (M // 2 + 1) * W, post); tt += a; ti += b

    def pre(q):
        if q < (M // 2) * W:
            line = (q & (W - 1)) if strided else (q >> (LOGM - 1))
            k = (q >> LOGW) if strided else (q & (M // 2 - 1))
        else:
            line, k = q - (M // 2) * W, M // 2
        p, p2 = brev(k, LOGM), brev((M - k) & (M - 1), LOGM)
        return [(16 * slot(line, pad(p), lay), 16), (16 * slot(line, pad(p2), lay), 16)]
    a, b = phase("inverse pre (2 STS.128, bit-reversed)", (M // 2 + 1) * W, pre); tt += a; ti += b

    def store(q):
        line = (q & (W - 1)) if strided else (q >> LOGM)
        j = (q >> LOGW) if strided else (q & (M - 1))
        a0, a1 = j >> 1, (M - 1) - (j >> 1)
        return [(16 * slot(line, pad(a0), lay) + 8 * (j & 1), 8), (16 * slot(line, pad(a1), lay) + 8 * (1 - (j & 1)), 8)]
    a, b = phase("inverse store (2 LDS.64)", M * W, store); tt += a; ti += b
    print(f"  total x{tt / ti:.2f}")


MP = _mp()
run(False)
run(True)
