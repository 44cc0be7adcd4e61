"""CPU: how many colours does the CUT set of a region partition need? Pyramid lattice of N rows (each box touches 6 others),
P regions by Hilbert chunks / slabs / k-d tiles / lattice-aligned k-d tiles; prints the share of cut constraints, the histogram
of cut constraints per boundary body and the colours several greedy edge colourings need. (The design decision it backs:
DESIGN.md 3.1, region-local schedule.)  Usage: python tools/region_cut_experiment.py [N=447] [P=296]"""
import numpy as np, sys, random
N=int(sys.argv[1]) if len(sys.argv)>1 else 447
P=int(sys.argv[2]) if len(sys.argv)>2 else 296
# pyramid lattice: row i has N-i boxes; box (i,j) j in [0,N-i): x = j + 0.5*i, y = i
idx={}
pos=[]
for i in range(N):
    for j in range(N-i):
        idx[(i,j)]=len(pos); pos.append((j+0.5*i, float(i)))
pos=np.array(pos); nb=len(pos)
edges=[]
for (i,j),b in idx.items():
    if (i,j+1) in idx: edges.append((b,idx[(i,j+1)]))
    if (i+1,j) in idx: edges.append((b,idx[(i+1,j)]))
    if (i+1,j-1) in idx: edges.append((b,idx[(i+1,j-1)]))
edges=np.array(edges)
print("bodies",nb,"edges",len(edges), "(+ground contacts, interior)")
def hilbert(x,y):
    d=0; s=32768
    while s>0:
        rx=1 if (x&s) else 0; ry=1 if (y&s) else 0
        d+=s*s*((3*rx)^ry)
        if ry==0:
            if rx==1: x=65535-x; y=65535-y
            x,y=y,x
        s>>=1
    return d
def regions_hilbert(P):
    mn=pos.min(0); ext=(pos.max(0)-mn).max(); sc=65535/ext
    q=((pos-mn)*sc).astype(np.int64).clip(0,65535)
    keys=np.array([hilbert(int(a),int(b)) for a,b in q])
    order=np.argsort(keys,kind='stable')
    chunk=(nb+P-1)//P
    reg=np.empty(nb,int); reg[order]=np.arange(nb)//chunk
    return reg
def regions_slab(P):
    order=np.lexsort((pos[:,0],pos[:,1]))
    chunk=(nb+P-1)//P
    reg=np.empty(nb,int); reg[order]=np.arange(nb)//chunk
    return reg
def analyse(reg,name):
    cut=reg[edges[:,0]]!=reg[edges[:,1]]
    ce=edges[cut]
    print(name,"cut edges",len(ce),"= %.1f%%"%(100*len(ce)/len(edges)))
    # adjacency of cut edges by body
    from collections import defaultdict
    inc=defaultdict(list)
    for k,(a,b) in enumerate(ce):
        inc[a].append(k); inc[b].append(k)
    deg=np.array([len(v) for v in inc.values()])
    print("  body cut-degree hist",np.bincount(deg))
    def greedy(order):
        col=-np.ones(len(ce),int)
        for k in order:
            a,b=ce[k]
            used=set(col[j] for j in inc[a])|set(col[j] for j in inc[b])
            c=0
            while c in used: c+=1
            col[k]=c
        return col.max()+1
    print("  greedy index order:",greedy(range(len(ce))))
    rnd=list(range(len(ce))); random.seed(1); random.shuffle(rnd)
    print("  greedy random order:",greedy(rnd))
    # primary colouring: first-fit in index order over ALL edges
    incall=defaultdict(list)
    for k,(a,b) in enumerate(edges):
        incall[a].append(k); incall[b].append(k)
    pc=-np.ones(len(edges),int)
    for k,(a,b) in enumerate(edges):
        used=set(pc[j] for j in incall[a])|set(pc[j] for j in incall[b])
        c=0
        while c in used: c+=1
        pc[k]=c
    print("  primary colours",pc.max()+1)
    cutidx=np.nonzero(cut)[0]
    order=sorted(range(len(ce)), key=lambda k:(pc[cutidx[k]],k))
    print("  iterated greedy by primary class:",greedy(order))
    # degree-ordered (largest conflict degree first)
    cdeg=[len(inc[a])+len(inc[b]) for a,b in ce]
    order=sorted(range(len(ce)), key=lambda k:-cdeg[k])
    print("  greedy largest-degree-first:",greedy(order))
analyse(regions_hilbert(P),"hilbert P=%d"%P)
analyse(regions_slab(P),"slab P=%d"%P)

def regions_kd(P, coords):
    reg=np.zeros(nb,int)
    def rec(ids, lo, hi):
        # assign regions lo..hi-1 to ids
        n=hi-lo
        if n==1:
            reg[ids]=lo; return
        c=coords[ids]
        ext=c.max(0)-c.min(0)
        ax=0 if ext[0]>=ext[1] else 1
        order=ids[np.argsort(c[:,ax],kind='stable')]
        nl=n//2
        split=len(ids)*nl//n
        rec(order[:split], lo, lo+nl); rec(order[split:], lo+nl, hi)
    rec(np.arange(nb),0,P)
    return reg
analyse(regions_kd(P,pos),"kd axis-aligned P=%d"%P)
sh=pos.copy(); sh[:,0]=pos[:,0]-0.5*pos[:,1]
analyse(regions_kd(P,sh),"kd sheared (lattice aligned) P=%d"%P)
