"""Bank-conflict simulation of the block-max pooling hot loop on the c2 descriptor stream (CPU only).

    python tools/bank_sim.py

Replays the geometry of csrc/pool_pyr.cuh in numpy (classes, first-block cells, distances to the last block),
walks the lane slots of one image the way pyr_run does (quarter-warps of 8 consecutive bins, one LDS.128 per
block position) and counts shared-memory wavefronts per quarter-warp load: plain layout against two column
rotations of the plane.  Result on c2: 1.69 wavefronts per ideal wavefront (ncu measures 1.66), 1.69 with
col + (col >> 3): the rotation only moves the conflicts from even bin steps to odd ones."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import synth
w = synth.workload('c2')
rois = w['rois'].numpy()[:8000]
H,W = w['features'].shape[2:]
scale = np.float32(1/8)
def roundf(v): return np.sign(v)*np.floor(np.abs(v)+np.float32(0.5))
def axis(lo,hi,L):
    rs = np.clip(roundf((lo*scale).astype(np.float32)),-1e6,1e6).astype(np.int64)
    re = np.clip(roundf((hi*scale).astype(np.float32)),-1e6,1e6).astype(np.int64)
    n = np.maximum(re-rs+1,1)
    b = (n.astype(np.float32)/np.float32(7)).astype(np.float32)
    S=[];E=[]
    for p in range(7):
        s = np.floor((np.float32(p)*b).astype(np.float32)).astype(np.int64)+rs
        e = np.ceil((np.float32(p+1)*b).astype(np.float32)).astype(np.int64)+rs
        S.append(np.clip(s,0,L)); E.append(np.clip(e,0,L))
    S=np.stack(S,1); E=np.stack(E,1); nn=E-S; ne=nn>0
    border=(S==0)|(E==L)
    ok4=np.all(~ne|(nn>=4)|border,1); ok2=np.all(~ne|(nn>=2)|border,1)
    k=np.where(ok4,4,np.where(ok2,2,1))
    nmax=np.max(np.where(ne,nn,0),1)
    c=np.maximum((nmax+k-1)//k,1)
    # blocks
    kk=k[:,None]
    p0=np.where(nn>=kk,S,np.where(S==0,E-kk,S)); last=np.where(nn>=kk,nn-kk,0)
    return k,c,p0,last,ne
kh,ch,p0h,lasth,neh=axis(rois[:,2],rois[:,4],H)
kw,cw,p0w,lastw,new=axis(rois[:,1],rois[:,3],W)
fb=(ch>4)|(cw>4)
CH=np.where(ch<=2,2,4); CW=np.where(cw<=2,2,4)
RS=W+3
def groups(row,col,mode):
    if mode==0: return (row*RS+col)%8
    if mode==1: return (row*RS+col+(col>>3))%8
    if mode==2: return (row*RS+col+(col>>3)+(col>>6))%8
tot={0:0,1:0,2:0}; ideal=0; nl=0
R=len(rois)
for r in range(R):
    if fb[r]: continue
    # 64 slots: bins 0..48 then idle; quarter-warps of 8 consecutive slots
    bins=np.arange(49); ph=bins//7; pw=bins%7
    empty=~(neh[r][ph]&new[r][pw])
    row0=p0h[r][ph]+3; col0=p0w[r][pw]+3; lh=lasth[r][ph]; lw=lastw[r][pw]
    for i in range(CH[r]):
        ni=(i==0)|((i-1)*kh[r]<lh)
        ro=np.minimum(i*kh[r],lh)
        for j in range(CW[r]):
            nj=(j==0)|((j-1)*kw[r]<lw)
            co=np.minimum(j*kw[r],lw)
            act=ni&nj
            if i==0 and j==0: act=np.ones(49,bool)
            rr=np.where(empty,H+5,row0+ro); cc=np.where(empty,0,col0+co)   # empty -> zero cell (all same)
            for q in range(0,49,8):
                a=act[q:q+8]
                if not a.any(): continue
                nl+=1
                cells=set(zip(rr[q:q+8][a].tolist(),cc[q:q+8][a].tolist()))
                ideal+=1
                for mode in (0,1,2):
                    cnt={}
                    for (y,x) in cells:
                        g=int(groups(y,x,mode)); cnt[g]=cnt.get(g,0)+1
                    tot[mode]+=max(cnt.values())
print('quarter-warp loads',nl,'wavefronts: plain',tot[0],'swz',tot[1],'swz2',tot[2],'factor plain',tot[0]/nl,'swz',tot[1]/nl, tot[2]/nl)
# breakdown: per proposal bin width (cells) vs conflict factor, first row of bins only, load (0,0)
import collections
acc=collections.defaultdict(lambda:[0,0,0])
for r in range(R):
    if fb[r]: continue
    cols=p0w[r]+3
    rows=np.full(7,p0h[r][0]+3)
    bw=int(np.median(np.diff(cols))) if len(cols)>1 else 0
    cells=set(zip(rows.tolist(),cols.tolist()))
    for mode in (0,1):
        cnt={}
        for (y,x) in cells:
            g=int(groups(y,x,mode)); cnt[g]=cnt.get(g,0)+1
        acc[bw][mode]+=max(cnt.values())
    acc[bw][2]+=1
for bw in sorted(acc): print('bin step',bw,'n',acc[bw][2],'plain',acc[bw][0]/acc[bw][2],'swz',acc[bw][1]/acc[bw][2])
