import numpy as np
rng=np.random.default_rng(0)
D=32;K=8;N=4096
nu=D+3.0
# random SPD sigma_mf, Rs = sqrt(nu/2) L^-1
x=(rng.standard_normal((N,D))*1.0+rng.standard_normal(D)*0.3).astype(np.float32).astype(np.float64)
x[5]*=1e3; x[7]*=1e-4
Rs=[];gk=[]
for k in range(K):
    A=rng.standard_normal((D,D)); S=A@A.T/D*2+np.eye(D)*0.5
    L=np.linalg.cholesky(S); R=np.sqrt(nu/2)*np.linalg.inv(L)
    mu=rng.standard_normal(D)*0.5
    Rs.append(np.tril(R)); gk.append(R@mu)
Rs=np.array(Rs);gk=np.array(gk)
ref=np.einsum('kij,nj->nki',Rs,x)-gk[None]
ref_q=(ref**2).sum(-1)
def digits(v,scale,ns=4):
    v=v/scale; out=[]; r=v*256
    for s in range(ns):
        d=np.rint(r); out.append(d); r=(r-d)*512
    return out
def pow2ceil(m):
    e=np.ceil(np.log2(np.maximum(m,1e-300)))
    e=np.where(2.0**e<=m,e+1,e)
    return 2.0**e
rsx=pow2ceil(np.abs(x).max(1))[:,None]
dx=digits(x,rsx)
Rm=Rs.reshape(K*D,D)
rsR=pow2ceil(np.abs(Rm).max(1))[:,None]
dR=digits(Rm,rsR)
for d in dx+dR: assert np.abs(d).max()<=256
Ls=[]
for l in range(4):
    acc=np.zeros((N,K*D))
    for i in range(l+1):
        acc+=dx[i]@dR[l-i].T
    assert np.abs(acc).max()<2**24
    Ls.append(acc.astype(np.float32))
f=np.float32
t=(Ls[3]*f(2**-9)+Ls[2]).astype(np.float32)   # fmaf
Lw=(t*f(2**-9)+Ls[1]).astype(np.float32)
a0=(Ls[0]*rsx.astype(np.float32)).astype(np.float32)
lw=(Lw*(rsx*2**-9).astype(np.float32)).astype(np.float32)
vd=a0.astype(np.float64)+lw.astype(np.float64)
cs=(rsR[:,0]*2.0**-16)
y=vd*cs[None]-gk.reshape(-1)[None]
q=(y.reshape(N,K,D)**2).sum(-1)
err=np.abs(q-ref_q)
print("quad max",ref_q.max(),"abs err max",err.max(),"rel",(err/ref_q).max())
print("err rows 5,7",err[5].max(),ref_q[5].max(),err[7].max())
# fp32 emulation for contrast
y32=(np.einsum('kij,nj->nki',Rs.astype(np.float32),x.astype(np.float32))-gk.astype(np.float32)[None])
print("fp32 err",np.abs((y32.astype(np.float64)**2).sum(-1)-ref_q).max())
mask=np.ones(N,bool);mask[5]=False
print("normal rows: quad max",ref_q[mask].max(),"abs err max",err[mask].max(), "median", np.median(err[mask]))
e32=np.abs((y32.astype(np.float64)**2).sum(-1)-ref_q)
print("fp32 normal rows err max",e32[mask].max())
