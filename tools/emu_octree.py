import sys, numpy as np
sys.path.insert(0,'.')
from tests import oracle_lib
from wut_cuda_orb_slam3_b200 import synth
from wut_cuda_orb_slam3_b200.capi import lib, ptr
o = oracle_lib.load()

def emu(xs, ys, sc, W, H, N, verbose=False):
    n = len(xs)
    nIni = int(np.floor(np.float32(W)/np.float32(H) + 0.5))
    hX = np.float32(W)/np.float32(nIni)
    def clog2(v):
        b=0
        while (1<<b) < v: b+=1
        return b
    maxRootW = max(int(hX*np.float32(i+1)) - int(hX*np.float32(i)) for i in range(nIni))
    D = max(clog2(maxRootW)+1, clog2(H)) + 1
    codes = []
    for x,y in zip(xs,ys):
        r = int(np.float32(x)/hX)
        ulx, urx = int(hX*np.float32(r)), int(hX*np.float32(r+1))
        uly, bry = 0, H
        c = r
        for d in range(D):
            mx = ulx + ((urx-ulx+1)>>1); my = uly + ((bry-uly+1)>>1)
            qx = int(x>=mx); qy=int(y>=my)
            c = (c<<2)|qx|(qy<<1)
            if qx: ulx=mx
            else: urx=mx
            if qy: uly=my
            else: bry=my
        codes.append(c)
    codes = np.array(codes, np.int64)
    order = np.argsort(codes, kind='stable')
    C = codes[order]; K = order
    # nodes: dict(beg,cnt,ulx,urx,dep)
    L = []
    for r in range(nIni):
        lo = np.searchsorted(C>>(2*D), r, 'left'); hi = np.searchsorted(C>>(2*D), r+1, 'left')
        if hi>lo: L.append(dict(beg=lo,cnt=hi-lo,ulx=int(hX*np.float32(r)),urx=int(hX*np.float32(r+1)),dep=0))
    def split(nd):
        shift = 2*(D-1-nd['dep'])
        seg = (C[nd['beg']:nd['beg']+nd['cnt']]>>shift)&3
        cn = [int((seg==k).sum()) for k in range(4)]
        mx = nd['ulx'] + ((nd['urx']-nd['ulx']+1)>>1)
        b = nd['beg']; ch=[]
        for k in range(4):
            ch.append(dict(beg=b,cnt=cn[k],ulx=(nd['ulx'] if k in (0,2) else mx), urx=(mx if k in (0,2) else nd['urx']),dep=nd['dep']+1)); b+=cn[k]
        return ch
    def apply(L, proc):  # proc = list positions in processing order
        groups=[]; vec=[]
        for pos in proc:
            ch = split(L[pos])
            groups.append([c for c in reversed(ch) if c['cnt']>0])
            vec += [c for c in ch if c['cnt']>1]
        newL=[]
        for g in reversed(groups): newL += g
        ps=set(proc)
        newL += [nd for i,nd in enumerate(L) if i not in ps]
        return newL, vec
    finish=False
    while not finish:
        prev=len(L)
        proc=[i for i,nd in enumerate(L) if nd['cnt']>1]
        L, vec = apply(L, proc)
        if len(L)>=N or len(L)==prev: finish=True
        elif len(L)+3*len(vec) > N:
            while not finish:
                prev2=len(L)
                items = np.array([(((v['cnt']<<13)|v['ulx'])<<24)|i for i,v in enumerate(vec)], np.uint64)
                srt = o.std_sort_hi40(items)
                sorted_vec = [vec[int(it)&0xffffff] for it in srt]
                proc=[]; size=prev2
                idpos = {id(nd):i for i,nd in enumerate(L)}
                for v in reversed(sorted_vec):
                    ch = split(v); ne = sum(1 for c in ch if c['cnt']>0)
                    proc.append(idpos[id(v)]); size += ne-1
                    if size>=N: break
                L, vec = apply(L, proc)
                if len(L)>=N or len(L)==prev2: finish=True
    out=[]
    for nd in L:
        seg = sorted(K[nd['beg']:nd['beg']+nd['cnt']])
        best = seg[0]
        for k in seg[1:]:
            if sc[k] > sc[best]: best=k
        out.append(int(best))
    return np.array(out)

if __name__=='__main__':
    img = synth.image(105,160,120)
    oex = o.extractor(300,1.2,8,20,7); oex.extract(img,(0,0))
    xs,ys,sc = oex.candidates(0)
    W,H = 160-32, 120-32
    ref = o.octree(xs,ys,sc,16,16+W,16,16+H,67)
    got = emu(xs,ys,sc,W,H,67)
    print(len(ref),len(got), (ref!=got).sum() if len(ref)==len(got) else 'len differs')
    rng=np.random.default_rng(0)
    bad=0
    for t in range(300):
        W=int(rng.integers(60,800)); H=int(rng.integers(40,500))
        if round(W/H)<1: continue
        n=int(rng.integers(1,3000)); N=int(rng.integers(1,400))
        pts=np.unique(np.stack([rng.integers(0,H,n),rng.integers(0,W,n)],1),axis=0)
        ys2,xs2=pts[:,0].astype(np.int32),pts[:,1].astype(np.int32); sc2=rng.integers(7,40,len(xs2)).astype(np.int32)
        ref=o.octree(xs2,ys2,sc2,16,16+W,16,16+H,N); got=emu(xs2,ys2,sc2,W,H,N)
        if len(ref)!=len(got) or (ref!=got).any(): bad+=1; print('BAD',W,H,len(xs2),N,len(ref),len(got))
    print('bad',bad)
    print('--- small search')
    rng=np.random.default_rng(1)
    for t in range(20000):
        W=int(rng.integers(40,120)); H=int(rng.integers(40,120))
        if round(W/H)<1: continue
        n=int(rng.integers(2,14)); N=int(rng.integers(1,10))
        pts=np.unique(np.stack([rng.integers(0,H,n),rng.integers(0,W,n)],1),axis=0)
        ys2,xs2=pts[:,0].astype(np.int32),pts[:,1].astype(np.int32); sc2=rng.integers(7,40,len(xs2)).astype(np.int32)
        ref=o.octree(xs2,ys2,sc2,16,16+W,16,16+H,N); got=emu(xs2,ys2,sc2,W,H,N)
        if len(ref)!=len(got) or (ref!=got).any():
            print('BAD',W,H,N,'pts',list(zip(xs2.tolist(),ys2.tolist(),sc2.tolist())),'ref',ref,'got',got); break
