"""Model of the table's line occupancy (measurement tool, CPU only): how many keys find their 13-slot home line full at a given
density, for the k-mers of a random genome homed by minimizer (m = 21, w = 11: runs of 1..11 consecutive k-mers, 6 on average, land
in one line together).  One-choice = the layout of csrc/kcf_db.cu; the two-choice column (a run goes whole to the emptier of two
candidate lines) is the alternative that was evaluated and not built: it moves the same share of keys out of the first line read.
Used to size the overflow region (kcf_db.cu) and quoted in DESIGN.md §3.

    python tools/line_occupancy_model.py
"""
import numpy as np, sys
rng=np.random.default_rng(1)
N=4_000_000; k=31; m=21; w=k-m+1
g=rng.integers(0,4,N,dtype=np.uint8)
# m-mer values (forward) as python ints via uint64
def mmers(codes,m):
    v=np.zeros(len(codes)-m+1,dtype=np.uint64)
    for j in range(m):
        v=(v<<np.uint64(2))|codes[j:len(codes)-m+1+j].astype(np.uint64)
    return v
f=mmers(g,m)
rc=mmers((3-g)[::-1],m)[::-1]
can=np.minimum(f,rc)
def mix(x):
    x=x.astype(np.uint64)
    x^=x>>np.uint64(33); x*=np.uint64(0xff51afd7ed558ccd); x^=x>>np.uint64(33); x*=np.uint64(0xc4ceb9fe1a85ec53); x^=x>>np.uint64(33)
    return x
h=mix(can)>>np.uint64(32)
# sliding min over w
from numpy.lib.stride_tricks import sliding_window_view
mu=sliding_window_view(h,w).min(axis=1)   # per k-mer
nk=len(mu)
# clumps: runs of equal mu (consecutive)
chg=np.flatnonzero(np.diff(mu)!=0)+1
starts=np.concatenate([[0],chg]); sizes=np.diff(np.concatenate([starts,[nk]]))
print("kmers",nk,"clumps",len(sizes),"mean",sizes.mean(),"max",sizes.max())
print("size hist",np.bincount(sizes)[:16]/len(sizes))
S=13
for dens in (0.15,0.3,0.5,0.7,0.8,0.9):
    nl=int(nk/(S*dens))
    home=(mix(mu[starts]^np.uint64(0x9E3779B9))%np.uint64(nl)).astype(np.int64)
    load=np.bincount(home,weights=sizes,minlength=nl)
    over=np.maximum(load-S,0).sum()/nk
    occ=(load>0).mean()
    # two-choice at clump level, greedy in random order
    h2=(mix(mu[starts]^np.uint64(0x85EBCA6B))%np.uint64(nl)).astype(np.int64)
    ld=np.zeros(nl,dtype=np.int64); spilled=0; second=0
    order=rng.permutation(len(sizes))
    for i in order:
        a,b,s=home[i],h2[i],sizes[i]
        if ld[a]+s<=S: ld[a]+=s
        elif ld[b]+s<=S: ld[b]+=s; second+=s
        else:
            # split: fill a then b then spill
            fa=max(S-ld[a],0); ld[a]+=min(fa,s); r=s-min(fa,s)
            fb=max(S-ld[b],0); t=min(fb,r); ld[b]+=t; second+=t; r-=t
            spilled+=r
    print(f"dens {dens}: lines {nl} occupied {occ:.3f} overflow-frac(one-choice) {over:.4f} | two-choice: second {second/nk:.4f} spilled {spilled/nk:.4f}")
