"""Model of csrc/scalar.cuh lattice3_reduce (three short scalars for the var-generator equation) in Python floats and integers:
the same basis, the same uniform iteration (order by length; Gauss step of the second against the shortest; the longest against
the plane of the other two), the same caps and the laggard step, so that iteration counts and the size of the result can be
studied without a GPU.   usage: tools/lattice3_model.py [trials] [seed] [cap bits] [inner iterations per round]"""
import random, math, sys, collections
r = 0x0e7db4ea6533afa906673b0101343b00a6682093ccc81082d0970e5ed6f72cb7
CAPBITS = int(sys.argv[3]) if len(sys.argv)>3 else 22
K_INNER = int(sys.argv[4]) if len(sys.argv)>4 else 10
CAP = float(1<<CAPBITS)
BIG=float(1<<20)
def half(c):
    a,b,ta,tb,neg = r,c,0,1,False
    while b >= 1<<126:
        q = a//b
        a,b = b,a-q*b
        ta,tb = tb,ta+q*tb
        neg = not neg
    s2 = -1 if neg else 1
    return (a, -s2*ta), (b, s2*tb)
SC=2.0**-130
def clamp(x):
    return max(-(CAP-1),min(CAP-1,x))
def reduce3(u,c,max_rounds=20):
    (t1,r1),(t2,r2) = half(c)
    sg=lambda v:(1 if v>=0 else -1)
    E = [[sg(r1)*((abs(r1)*u)%r), t1, r1],[sg(r2)*((abs(r2)*u)%r), t2, r2],[r,0,0]]
    stats=dict(rounds=0,useful=0,maxq=0,big=0)
    conv=False
    for rd in range(max_rounds):
        if conv: break
        stats['rounds']+=1
        D=[[float(x)*SC for x in v] for v in E]
        n=[sum(x*x for x in v) for v in D]
        # laggards: a row that needs a quotient beyond 2^20 against the shortest gets it exactly, 31 bits at a time
        for tgt in (1,2):
            ratio=sum(a*b for a,b in zip(D[tgt],D[0]))/n[0]
            if abs(ratio)>=BIG:
                m=abs(ratio); sh=0; sc=1.0
                while m>=2147483648.0 and sh<8: m*=2.0**-32; sc*=2.0**-32; sh+=1
                k=int(round(ratio*sc))
                E[tgt]=[x-((k*y)<<(32*sh)) for x,y in zip(E[tgt],E[0])]
                D[tgt]=[float(x)*SC for x in E[tgt]]; n[tgt]=sum(x*x for x in D[tgt]); stats['big']+=1
        T=[[1.0,0,0],[0,1.0,0],[0,0,1.0]]
        for it in range(K_INNER):
            o=sorted(range(3),key=lambda k:n[k])
            D=[D[k] for k in o]; T=[T[k] for k in o]; n=[n[k] for k in o]
            dot=lambda i,j: sum(a*b for a,b in zip(D[i],D[j]))
            did=False
            # step 1: Gauss on (A,B)
            ab=dot(0,1); ra=ab/n[0]
            if abs(ra)>0.500001:
                k=clamp(float(round(ra)))
                newT=[x-k*y for x,y in zip(T[1],T[0])]
                if max(abs(x) for x in newT)<=CAP:
                    T[1]=newT; D[1]=[x-k*y for x,y in zip(D[1],D[0])]; n[1]=sum(x*x for x in D[1]); did=True
                    stats['maxq']=max(stats['maxq'],abs(k))
                ab=dot(0,1)
            # step 2: C against A,B: Babai if (A,B) well conditioned else pairwise with the shorter... 
            ac=dot(0,2); bc=dot(1,2)
            if abs(ab) <= 0.55*min(n[0],n[1]):
                det=n[0]*n[1]-ab*ab
                a=(ac*n[1]-bc*ab)/det; b=(bc*n[0]-ac*ab)/det
            else:
                a=ac/n[0]; b=0.0
            ka=clamp(float(round(a))); kb=clamp(float(round(b)))
            if ka!=0 or kb!=0:
                newT=[x-ka*y-kb*z for x,y,z in zip(T[2],T[0],T[1])]
                if max(abs(x) for x in newT)<=CAP:
                    T[2]=newT; D[2]=[x-ka*y-kb*z for x,y,z in zip(D[2],D[0],D[1])]; n[2]=sum(x*x for x in D[2]); did=True
                    stats['maxq']=max(stats['maxq'],abs(ka),abs(kb))
            elif abs(ra)<=0.500001:
                conv=True
            if did: stats['useful']+=1
        o=sorted(range(3),key=lambda k:n[k]); T=[T[k] for k in o]
        Ti=[[int(x) for x in row] for row in T]
        E=[[sum(Ti[i][j]*E[j][k] for j in range(3)) for k in range(3)] for i in range(3)]
    return E,stats
if __name__=="__main__":
    random.seed(int(sys.argv[2]) if len(sys.argv)>2 else 1)
    N=int(sys.argv[1]) if len(sys.argv)>1 else 500
    mx=[];S=collections.defaultdict(list);oddbits=[]
    for t in range(N):
        u=random.randrange(r); c=random.randrange(1<<250)
        B,st=reduce3(u,c)
        for v in B:
            assert (v[0]-v[2]*u)%r==0 and (v[1]-v[2]*c)%r==0
        bits=[max(abs(x).bit_length() for x in v) for v in B]
        mx.append(min(bits))
        for k,v in st.items(): S[k].append(v)
        ob=[b for b,v in zip(bits,B) if v[2]&1]
        oddbits.append(min(ob) if ob else 999)
    for k,v in S.items(): print(k,"avg",sum(v)/len(v),"max",max(v), "hist", sorted(collections.Counter(v).items())[-6:] if k=='rounds' else "")
    print("min bits hist",sorted(collections.Counter(mx).items()))
    print("odd basis-only hist",sorted(collections.Counter(oddbits).items()))
