"""Reference swap-with-last removal vs the ballot-based compaction of l_reduce_region_radius (line.cu)."""
import random
def ref(reg, keepf):
    reg=list(reg); i=0
    while i < len(reg):
        if not keepf(reg[i]):
            reg[i], reg[-1] = reg[-1], reg[i]; reg.pop(); continue
        i+=1
    return reg
def popc(x): return bin(x).count('1')
def mine(reg, keepf):
    reg=list(reg); n=len(reg)
    m2=sum(1 for v in reg if keepf(v))
    lo=0; hi=n; nfill=0; fpos=0; s_fill=[None]*32
    while lo < m2:
        lanes=[]
        hm=0
        for lane in range(32):
            i=lo+lane
            if i<m2 and not keepf(reg[i]): hm|=1<<lane
        while hm:
            if fpos==nfill:
                nfill=0; fpos=0
                while nfill==0 and hi>m2:
                    fm=0
                    for lane in range(32):
                        j=hi-1-lane
                        if j>=m2 and keepf(reg[j]): fm|=1<<lane
                    for lane in range(32):
                        if (fm>>lane)&1: s_fill[popc(fm&((1<<lane)-1))]=reg[hi-1-lane]
                    nfill=popc(fm); hi-=32
                assert nfill>0
            t=min(popc(hm), nfill-fpos)
            served=0
            for lane in range(32):
                r=popc(hm&((1<<lane)-1))
                if (hm>>lane)&1 and r<t:
                    reg[lo+lane]=s_fill[fpos+r]; served|=1<<lane
            fpos+=t; hm&=~served
        lo+=32
    return reg[:m2]
def run(iters=3000, seed=0):
    random.seed(seed)
    for it in range(iters):
        n=random.randint(1,300); pr=random.random()
        reg=list(range(n)); keep={v:(random.random()<pr) for v in reg}
        a=ref(reg, lambda v:keep[v]); b=mine(reg, lambda v:keep[v])
        assert a==b,(n,a,b)
    return True

if __name__ == '__main__':
    run(); print('ok')
