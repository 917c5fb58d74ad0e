"""Padding search used for csrc/bk_fft_fast.cuh::Cfg::pad (element-major layout, pairs interleaved); see model.bank_report_em."""
import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from model import *
def report(pl, slot, PP, nthreads):
    n,E,T=pl.n,pl.E,pl.T
    lanes=[(tid//PP, tid%PP) for tid in range(nthreads)]
    res=[]
    for p,r in enumerate(pl.rad):
        tot=ideal=0
        for w0 in range(0,len(lanes),32):
            warp=lanes[w0:w0+32]
            for u in range(E//r):
                for a in range(r):
                    sl=[slot(pr, pl.positions(p,tau,u)[0][a]) for tau,pr in warp]
                    tot+=wavefronts(sl); ideal+=(len(sl)+7)//8
        res.append(tot/ideal)
    # partner read
    tot=ideal=0
    rl=pl.rad[-1]
    for w0 in range(0,len(lanes),32):
        warp=lanes[w0:w0+32]
        for u in range(E//rl):
            for j in range(rl):
                sl=[]
                for tau,pr in warp:
                    p=pl.positions(len(pl.rad)-1,tau,u)[0][brev(j,rl)]
                    k=pl.k_of_pos[p]
                    sl.append(slot(pr, pl.pos_of_k[(n-k)%n]))
                tot+=wavefronts(sl); ideal+=(len(sl)+7)//8
    res.append(tot/ideal)
    return res
for E in (32,16):
  for n in (64,128,256,512,1024,2048):
    if E>n: continue
    pl=Plan(n,E); T=n//E
    PP=max(2,64//T); nthr=T*PP
    out=[]
    for a in range(2,7):
        for b in [0]+list(range(a+1,9)):
            pad=(lambda i,a=a,b=b: i+(i>>a)+((i>>b) if b else 0))
            # element-major
            r=report(pl, lambda pr,i: pad(i)*PP+pr, PP, nthr)
            out.append((sum(r),'EM',a,b,0,r))
            NP0=pad(n-1)+1
            for off in range(8):
                NP=NP0+off
                r=report(pl, lambda pr,i: pr*NP+pad(i), PP, nthr)
                out.append((sum(r),'PM',a,b,NP%8,r))
    out.sort(key=lambda t:t[0])
    print(n,E,pl.rad,'PP',PP)
    for x in out[:4]: print('   ',x[1:5],['%.2f'%v for v in x[5]])
print("---- EM candidates")
for E in (32,16):
  for n in (64,128,256,512,1024,2048):
    if E>n: continue
    pl=Plan(n,E); T=n//E
    PP=max(2,64//T); nthr=T*PP
    for (a,b) in ((2,4),(2,5),(2,0),(3,5)):
        pad=(lambda i,a=a,b=b: i+(i>>a)+((i>>b) if b else 0))
        r=report(pl, lambda pr,i: pad(i)*PP+pr, PP, nthr)
        print(n,E,(a,b),['%.2f'%v for v in r])
