"""Offline statistics for a digestion order (bra, D-block, C, D) on (H2O)16/cc-pVDZ (TEST/ANALYSIS TOOL, uses the CPU oracle
for the Schwarz bounds): how full would 32-lane groups "one shell C x 32 consecutive shells D" be, and what would that
order cost the thread-per-quartet ERI kernels in unequal primitive trip counts (max / mean over 32-task chunks) if they
walked the list without a permutation.  Output quoted in profiles/r02/digest_history.md.   python tools/digest_order_stats.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, ctypes as C, collections, random
import numpy as np, ctypes as C, collections
import oracle, quiqbox_b200 as qb
from molecules import water_cluster
nuc,xyz = water_cluster(16)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
n=ob.nbf; M=n*(n+1)//2
Q=np.zeros(M); oracle.lib().orc_schwarz(C.byref(ob.s), oracle._p(Q))
Qf=np.zeros((n,n))
iu=np.triu_indices(n)  # i<=j ; orc_tri2: n -> (i<=j) with n = i + j(j+1)/2
for j in range(n):
    for i in range(j+1):
        Qf[i,j]=Qf[j,i]=Q[i+j*(j+1)//2]
# shells: group functions by (center, exps, l)
shells=[]; key2shell={}
for f,g in enumerate(bs):
    k=(tuple(np.round(g.center,10)), tuple(g.xpns), sum(g.ang), tuple(np.round(np.array(g.cons)/g.cons[0],10)) if hasattr(g,'cons') else 0)
    if k not in key2shell:
        key2shell[k]=len(shells); shells.append(dict(l=sum(g.ang), cen=np.array(g.center), xp=np.array(g.xpns), co=np.array(g.cons), fn=[]))
    shells[key2shell[k]]['fn'].append(f)
print(len(shells),"shells")
order=sorted(range(len(shells)), key=lambda s:(shells[s]['l'], -len(shells[s]['xp'])))
shells=[shells[s] for s in order]
of_l={l:[i for i,s in enumerate(shells) if s['l']==l] for l in (0,1,2)}
pref=np.sqrt(2)*np.pi**1.25
def nprim(a,b):
    A,B=shells[a],shells[b]; r2=((A['cen']-B['cen'])**2).sum()
    x=A['xp'][:,None]; y=B['xp'][None,:]; z=x+y
    K=pref*A['co'][:,None]*B['co'][None,:]*np.exp(-x*y/z*r2)/z
    return int((np.abs(K)>=1e-24).sum())
def qpair(a,b):
    return max(Qf[i,j] for i in shells[a]['fn'] for j in shells[b]['fn'])
pos_in_l={}
for l in (0,1,2):
    for k,s in enumerate(of_l[l]): pos_in_l[s]=k
def pairsA(la,lb):
    SA,SB=of_l[la],of_l[lb]; out=[]
    if la==lb:
        for ia_,a in enumerate(SA):
            for b in SA[:ia_+1]: out.append((a,b))
    else:
        for a in SA:
            for b in SB: out.append((a,b))
    return out
cache={}
def buildA(la,lb):
    if (la,lb) in cache: return cache[(la,lb)]
    P=pairsA(la,lb); cnt=[nprim(a,b) for a,b in P]
    idx=sorted(range(len(P)), key=lambda i:-cnt[i])
    P=[P[i] for i in idx]; q=np.array([qpair(a,b) for a,b in P]); c=np.array([cnt[i] for i in idx])
    cache[(la,lb)]=(P,q,c); return cache[(la,lb)]
cls=[(0,0),(1,0),(1,1),(2,0),(2,1),(2,2)]
print("class            rows  quartets/row  groups/row  fill   ERIwaste_old ERIwaste_new")
for bi,bl in enumerate(cls):
    for kl in cls[:bi+1]:
        PB,qb_,cb=buildA(*bl); PK,qk,ck=buildA(*kl); same=bl==kl
        rows=random.Random(1).sample(range(len(PB)), min(40,len(PB)))
        nq=0; ng=0; w_old=[0,0]; w_new=[0,0]
        for i in rows:
            jmax=i+1 if same else len(PK)
            surv=[j for j in range(jmax) if qb_[i]*qk[j]>=1e-12]
            nq+=len(surv)
            key=lambda j:(pos_in_l[PK[j][1]]//32, PK[j][0], PK[j][1])
            groups=collections.Counter((pos_in_l[PK[j][1]]//32, PK[j][0]) for j in surv)
            ng+=len(groups)
            for order,w in ((surv,w_old),(sorted(surv,key=key),w_new)):
                for t0 in range(0,len(order),32):
                    c=ck[order[t0:t0+32]]; w[0]+=c.max()*len(c); w[1]+=c.sum()
        print("(%d%d|%d%d) %6d %10.1f %10.1f   %.2f      %.3f      %.3f"%(bl+kl+(len(rows),nq/len(rows),ng/len(rows),nq/max(1,ng*32),w_old[0]/max(1,w_old[1]),w_new[0]/max(1,w_new[1]))))
