"""Which sliced operand limits the gradient of the int8 GLM path at the full configs[1] size (CPU, long double): X, beta and the\nadjoint are quantised one at a time (2^-54 of their power-of-two scale, truncated or rounded to nearest) and X'(Y - X B) is\ncompared with the unquantised long-double value, relative to max |G| per chain.  python tools/ozaki_error_parts.py"""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, bench
import test_ozaki_numerics as T
N,M,C=1_000_000,64,4
X,Y,b0,_=bench.make_data(); X=np.ascontiguousarray(X)
mode=bench.posterior_mode(X,Y)
B=(mode[None,:]+1e-3*np.random.default_rng(100).standard_normal((4096,M)))[:C].T.copy()
def q(v,scale,rnd): return rnd(v*(2.0**54/scale))*(scale/2.0**54)
def G_of(Xa,Ba,dq=None):
    Gt=np.zeros((M,C),np.longdouble); Bl=Ba.astype(np.longdouble)
    for r0 in range(0,N,100_000):
        Xl=Xa[r0:r0+100_000].astype(np.longdouble); d=Y[r0:r0+100_000].astype(np.longdouble)[:,None]-Xl@Bl
        if dq is not None:
            d64=d.astype(np.float64)
            for t0 in range(0,d64.shape[0],128):
                blk=d64[t0:t0+128]; sd=T._scale(np.abs(blk).max(axis=0)); d64[t0:t0+128]=q(blk,sd[None,:],dq)
            d=d64.astype(np.longdouble)
        Gt+=Xl.T@d
    return Gt
G0=G_of(X,B)
rel=lambda a: float(np.max(np.abs(a-G0)/np.max(np.abs(G0),axis=0)))
sx=float(T._scale(np.abs(X).max())); sb=T._scale(np.abs(B).max(axis=0))
for name,rnd in (("trunc",np.trunc),("rint",np.rint)):
    print(name,"X only",rel(G_of(q(X,sx,rnd),B)),"B only",rel(G_of(X,q(B,sb[None,:],rnd))),"d only",rel(G_of(X,B,dq=rnd)),flush=True)
print("eta computed in float64 then exact:", rel(G_of(X,B,dq=lambda z:z)))
