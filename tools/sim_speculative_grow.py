"""Sequential model vs the speculative multi-accept rounds of l_region_grow (line.cu): python tools/sim_speculative_grow.py"""
import numpy as np, random, math
f32=np.float32
def atan2deg(y,x):
    a=f32(math.degrees(math.atan2(float(y),float(x))))
    if a<0: a=f32(a+f32(360))
    return a
DEG=math.pi/180
def aligned(ad,theta,prec):
    n=abs(theta-ad)
    if n>1.5*math.pi: n=abs(n-2*math.pi)
    return n<=prec
def seq(lanes,sx,sy,ang,prec):
    acc=[];used=set()
    for j,(p,q,a,cx,cy) in enumerate(lanes):
        if not p or q in used: continue
        if aligned(float(a)*DEG,ang,prec):
            acc.append(j);used.add(q); sx=f32(sx+cx); sy=f32(sy+cy); ang=float(atan2deg(sy,sx))*DEG
    return acc,sx,sy,ang
def popc(x): return bin(x).count('1')
def spec(lanes,sx,sy,ang,prec):
    pending=sum(1<<j for j,l in enumerate(lanes) if l[0])
    accepted=[]
    rounds=0
    while pending:
        rounds+=1
        mep=[(pending>>j)&1 for j in range(32)]
        S0=sum(1<<j for j in range(32) if mep[j] and aligned(float(lanes[j][2])*DEG,ang,prec))
        if not S0: break
        if S0&(S0-1)==0:
            k0=S0.bit_length()-1
            accepted.append(k0); sx=f32(sx+lanes[k0][3]); sy=f32(sy+lanes[k0][4]); ang=float(atan2deg(sy,sx))*DEG
            pending&=~((2<<k0)-1)
            pending&=~sum(1<<j for j in range(32) if lanes[j][1]==lanes[k0][1])
            continue
        peers=[sum(1<<k for k in range(32) if mep[k] and mep[j] and lanes[k][1]==lanes[j][1]) if mep[j] else (1<<j) for j in range(32)]
        S=S0&~sum(1<<j for j in range(32) if (S0>>j)&1 and (peers[j]&S0&((1<<j)-1)))
        bx=[sx]*32;by=[sy]*32
        T=S
        while T:
            m=(T&-T).bit_length()-1; T&=T-1
            for j in range(m+1,32): bx[j]=f32(bx[j]+lanes[m][3]); by[j]=f32(by[j]+lanes[m][4])
        ax=[f32(bx[j]+lanes[j][3]) for j in range(32)]; ay=[f32(by[j]+lanes[j][4]) for j in range(32)]
        aft=[float(atan2deg(ay[j],ax[j]))*DEG for j in range(32)]
        bef=[]
        for j in range(32):
            prevm=S&((1<<j)-1)
            bef.append(aft[prevm.bit_length()-1] if prevm else ang)
        actual=[bool(mep[j] and not (peers[j]&S&((1<<j)-1)) and aligned(float(lanes[j][2])*DEG,bef[j],prec)) for j in range(32)]
        mism=sum(1<<j for j in range(32) if mep[j] and actual[j]!=bool((S>>j)&1))
        if not mism: A=S; src=S.bit_length()-1
        else:
            src=(mism&-mism).bit_length()-1
            A=(S&((1<<src)-1))|((1<<src) if actual[src] else 0)
        assert A
        accepted+= [j for j in range(32) if (A>>j)&1]
        if actual[src]: sx,sy,ang=ax[src],ay[src],aft[src]
        else: sx,sy,ang=bx[src],by[src],bef[src]
        if not mism: break
        pending&=~((2<<src)-1)
        pending&=~sum(1<<j for j in range(32) if mep[j] and (peers[j]&A))
    return accepted,sx,sy,ang,rounds
def run(iters=20000, seed=1):
    random.seed(seed); tot_r=0; tot_a=0
    for it in range(iters):
        base=random.uniform(0,360); spread=random.choice([5,15,25,40])
        npix=random.randint(3,14)
        lanes=[]
        for j in range(32):
            p=random.random()<0.4
            q=random.randint(0,npix)
            a=f32((base+random.uniform(-spread,spread))%360)
            lanes.append((p,q,a,f32(math.cos(float(a)*DEG)),f32(math.sin(float(a)*DEG))))
        # same q => same attrs
        attrs={}
        lanes=[(p,q)+attrs.setdefault(q,(a,cx,cy)) for (p,q,a,cx,cy) in lanes]
        n0=random.randint(1,30)
        sx=f32(n0*math.cos(base*DEG)); sy=f32(n0*math.sin(base*DEG)); ang=float(atan2deg(sy,sx))*DEG
        r1=seq(lanes,sx,sy,ang,math.pi/8); r2=spec(lanes,sx,sy,ang,math.pi/8)
        assert r1[0]==r2[0] and r1[1]==r2[1] and r1[2]==r2[2] and r1[3]==r2[3],(it,r1,r2)
        tot_r+=r2[4]; tot_a+=len(r1[0])
    return tot_a/iters, tot_r/iters

if __name__ == '__main__':
    a, r = run()
    print('ok', a, 'accepts/iter', r, 'rounds/iter')
