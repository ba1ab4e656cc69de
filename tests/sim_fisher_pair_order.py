#!/usr/bin/env python
"""Lane-occupancy simulation behind DESIGN.md's K3 ordering numbers (CPU only, not a test).

For 24 junctions x 2,016 sample pairs of the configs[2] distribution it computes every table's
true near- and far-tail length under the kernel's stopping rule (term below 2^-48 of the running
sum, tested every four terms) and replays how warps of 32 consecutive pairs would spend their
tail iterations under different pair orders: pair order, the kernel's 64 cost buckets, finer
buckets, exact / oracle sorts, and two-key (near, far) orders on predicted and on true lengths.

    python tests/sim_fisher_pair_order.py

Lives under tests/ because it uses the CPU oracle for the exclusion sums."""
import os
import sys

import numpy as np
from scipy.special import gammaln

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np  # noqa: E402
from splicedice_b200 import synth  # noqa: E402

J,S=3000,64
c,s,st,en,_,_=synth.junction_arrays(J,20261025)
csr=oracle_np.cluster_csr(c,s,st,en)
inc=(synth.counts_host(8,0,J,S)+synth.counts_host(9,0,J,S)).astype(np.int64)
exc=oracle_np.exclusion_sums(inc.astype(np.int32),csr['row_ptr'],csr['col_idx']).astype(np.int64)
pa,pb=oracle_np.all_pairs(S)
rows=np.random.default_rng(0).choice(J,24,replace=False)
CUT=2.0**-48
def tail_iters(t):
    # t: term ratios relative to first (t[0]=1), beyond support -> 0. returns number of 4-step iterations
    # loop: do {4 steps; } while(!done) ; step k adds term k (k>=1)
    n=len(t); S=1.0; k=0; it=0
    while True:
        for i in range(4):
            k+=1; S+= t[k] if k<n else 0.0
        it+=1
        P=t[k] if k<n else 0.0
        if P < CUT*S: return it
def analyse(a,b,c,d):
    n1,n2,n=a+b,c+d,a+c; N=n1+n2
    if n1==0 or n2==0 or n==0 or b+d==0: return None
    mode=int((n+1)*(n1+1)/(N+2))
    if a==mode: return None
    if a>mode:
        a,b=b,a; c,d=d,c; n=N-n; mode=n1-mode
    lo=max(0,n-n2); hi=min(n1,n)
    x=np.arange(lo,hi+1)
    lp=-(gammaln(x+1)+gammaln(n1-x+1)+gammaln(n-x+1)+gammaln(n2-n+x+1))
    la=lp[a-lo]
    if abs(lp[mode-lo]-la)<=1e-14: return None
    near=np.exp(lp[:a-lo+1][::-1]-la)   # a, a-1, ...
    up=np.nonzero(lp[mode-lo+1:]<=la+1e-14)[0]
    ni=tail_iters(near)
    fi=0
    if len(up):
        g=mode+1+up[0]
        far=np.exp(lp[g-lo:]-lp[g-lo]); fi=tail_iters(far)
    # prediction as in cost_bucket
    fa,fb,fc,fd=map(float,(a,b,c,d)); n1f=fa+fb;n2f=fc+fd;nf=fa+fc;Nf=n1f+n2f
    var=max((nf/Nf)*(n1f/Nf)*(n2f*(Nf-nf))/max(Nf-1,1),1e-6); sig=var**.5
    z=abs(fa-np.floor((nf+1)*(n1f+1)/(Nf+2)))/sig
    terms=sig*(np.sqrt(z*z+66.5)-z)
    return ni,fi,terms
res=[]
for j in rows:
    near=np.zeros(len(pa),int); far=np.zeros(len(pa),int); pred=np.zeros(len(pa))
    for k,(sa,sb) in enumerate(zip(pa,pb)):
        r=analyse(int(inc[j,sa]),int(inc[j,sb]),int(exc[j,sa]),int(exc[j,sb]))
        if r: near[k],far[k],pred[k]=r
    res.append((near,far,pred))
def occ(order_fn,label):
    tot=0;used=0
    for near,far,pred in res:
        o=order_fn(near,far,pred)
        for g in range(0,len(o),32):
            idx=o[g:g+32]
            w=near[idx].max()+far[idx].max()
            tot+=32*w; used+=(near[idx]+far[idx]).sum()
    print(f"{label:40s} lane occupancy in tails {used/tot:.3f}  warp-iterations {tot/32/len(res):.0f} per junction")
occ(lambda n,f,p: np.arange(len(n)),"pair order")
def bucket(p,scale,nb): return np.where(p>0,np.minimum(nb-1,1+(scale*np.log2(1+p)).astype(int)),0)
occ(lambda n,f,p: np.argsort(-bucket(p,6,64),kind='stable'),"64 buckets (current)")
occ(lambda n,f,p: np.argsort(-bucket(p,12,128),kind='stable'),"128 buckets")
occ(lambda n,f,p: np.argsort(-bucket(p,24,256),kind='stable'),"256 buckets")
occ(lambda n,f,p: np.argsort(-p,kind='stable'),"exact sort by predicted")
occ(lambda n,f,p: np.argsort(-(n+f),kind='stable'),"oracle sort by true total")
occ(lambda n,f,p: np.lexsort((-f,-n)),"oracle sort by near then far")
occ(lambda n,f,p: np.argsort(-np.maximum(n,f),kind='stable'),"oracle sort by max(near,far)")
tot=sum((n+f).sum() for n,f,p in res); print("mean 4-step iterations per table", tot/len(res)/len(pa), "near share", sum(n.sum() for n,f,p in res)/tot)

print('--- orders on predicted near / far lengths (support-truncated) ---')
def preds(j):
    a=inc[j,pa].astype(float); b=inc[j,pb].astype(float); c=exc[j,pa].astype(float); d=exc[j,pb].astype(float)
    n1=a+b;n2=c+d;n=a+c;N=n1+n2
    ok=(n1>0)&(n2>0)&(n>0)&(b+d>0)
    N_=np.where(ok,N,1)
    mode=np.floor((n+1)*(n1+1)/(N_+2))
    swap=a>mode
    a2=np.where(swap,b,a); n_=np.where(swap,N-n,n); mode2=np.where(swap,n1-mode,mode)
    lo=np.maximum(0,n_-n2); hi=np.minimum(n1,n_)
    var=np.maximum((n_/N_)*(n1/N_)*(n2*(N-n_))/np.maximum(N-1,1),1e-6); sig=np.sqrt(var)
    z=np.abs(a2-mode2)/sig
    terms=sig*(np.sqrt(z*z+66.5)-z)
    near=np.minimum(terms,a2-lo+1)
    g=np.minimum(2*mode2-a2,hi)
    far=np.minimum(terms,hi-g+1)
    far=np.where(g>hi,0,far)
    triv=(~ok)|(a2==mode2)
    return np.where(triv,0,terms),np.where(triv,0,near),np.where(triv,0,far)
P=[preds(j) for j in rows]
def occ2(order_fn,label):
    tot=0;used=0
    for (near,far,pred),(pt,pn,pf) in zip(res,P):
        near=near.astype(int);far=far.astype(int)
        o=order_fn(pt,pn,pf)
        for g in range(0,len(o),32):
            idx=o[g:g+32]
            w=near[idx].max()+far[idx].max()
            tot+=32*w; used+=(near[idx]+far[idx]).sum()
    print(f"{label:60s} occupancy {used/tot:.3f}  warp-iterations {tot/32/len(res):.0f}")
def bucket(p,scale,nb): return np.where(p>0,np.minimum(nb-1,1+(scale*np.log2(1+p)).astype(int)),0)
occ2(lambda t,n,f: np.argsort(-bucket(t,6,64),kind='stable'),"current: 64 buckets of predicted terms")
occ2(lambda t,n,f: np.argsort(-bucket(n+f,6,64),kind='stable'),"64 buckets of near+far (support-truncated)")
occ2(lambda t,n,f: np.argsort(-bucket(np.maximum(n,f),6,64),kind='stable'),"64 buckets of max(near,far)")
occ2(lambda t,n,f: np.lexsort((-f,-n)),"exact lexsort near,far (predicted)")
for nb in (4,8,16):
  for sc in (1,1.5,2,3):
    occ2(lambda t,n,f: np.lexsort((-bucket(f,6,64),-bucket(n,sc,nb))),f"near {nb} buckets scale {sc} major, far 64 minor")
for sc in (1.0,1.5,2.0):
    occ2(lambda t,n,f: np.argsort(-(bucket(n,sc,8)*8+bucket(f,sc,8)),kind='stable'),f"single key 8x8 scale {sc}")
for sc in (1.5,2.0,2.5):
    occ2(lambda t,n,f: np.argsort(-(bucket(n,sc,16)*16+bucket(f,sc,16)),kind='stable'),f"single key 16x16 scale {sc}")
# rank-based: near rank quantiles (adaptive) major 8 slices, far minor
def adaptive(t,n,f,k):
    o=np.argsort(-n,kind='stable'); out=[]
    for s in range(0,len(o),len(o)//k):
        sl=o[s:s+len(o)//k]; out.append(sl[np.argsort(-f[sl],kind='stable')])
    return np.concatenate(out)
for k in (4,7,9,14):
    occ2(lambda t,n,f: adaptive(t,n,f,k),f"adaptive: {k} near-slices, far sorted within")
