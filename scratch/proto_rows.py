# numpy prototype of the per-slice algorithm of the rows kernel:
# (m,s) from inf-norm bound with safe limit 2 ln 2, Pade m<=9 scaled by 1/b0, unpivoted Gauss-Jordan
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import c3_oracle as orc
import scipy.linalg
LIM = 1.386
def slice_expm(A):
    D=A.shape[0]
    nb = np.abs(A).sum(axis=1).max()
    s=0
    while nb/2**s >= LIM: s+=1
    ns = nb/2**s
    m = 3 if ns<orc.THETA[0] else 5 if ns<orc.THETA[1] else 7 if ns<orc.THETA[2] else 9
    b = np.array(orc.PADE_B[m]); c=b/b[0]
    A = A/2**s
    I=np.eye(D)
    A2=A@A
    W=c[1]*I+c[3]*A2; V=c[0]*I+c[2]*A2
    X=A2
    for i in range((m-3)//2):
        X = X@A2
        W += c[2*i+5]*X; V += c[2*i+4]*X
    U = W@A
    Q=V-U; R=V+U
    growth=0
    for k in range(D):
        pv_q=Q[k].copy(); pv_r=R[k].copy()
        inv=1/pv_q[k]
        f = Q[:,k]*inv
        f[k] = 1-inv
        Q[:,k+1:] -= np.outer(f, pv_q[k+1:])
        R -= np.outer(f, pv_r)
        growth=max(growth,np.abs(Q).max())
    for i in range(s): R=R@R
    return R, m, s, growth
rng=np.random.default_rng(0)
worst=0
for d in [2,3,4,9,12,16]:
  for scale in [1e-3,0.01,0.2,0.5,0.9,1.2,1.38,1.39,2.0,3.0,5.0,11.,40.]:
    for herm in [True,False]:
      for t in range(5):
        h=rng.normal(size=(d,d))+1j*rng.normal(size=(d,d))
        if herm: h=h+h.conj().T; a=-1j*h
        else: a=h
        a*=scale/np.abs(a).sum(axis=1).max()
        R,m,s,gr=slice_expm(a)
        ref=scipy.linalg.expm(a)
        err=np.linalg.norm(R-ref)/np.linalg.norm(ref)
        worst=max(worst,err)
        if err>1e-13: print(d,scale,herm,m,s,err,gr)
print("worst",worst)
