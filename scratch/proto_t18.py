# numpy prototype: Bader-Blanes-Casas degree-18 Taylor scheme (5 products) + scaling & squaring
import numpy as np, scipy.linalg, sys
sys.path.insert(0,'/root/repo')
a11=-0.10036558103014462001; a21=-0.00802924648241156960; a31=-0.00089213849804572995
b11=0.39784974949964507614; b21=1.36783778460411719922; b31=0.49828962252538267755; b61=-0.00063789819459472330
b02=-10.9676396052962062593; b12=1.68015813878906197182; b22=0.05717798464788655127; b32=-0.00698210122488052084; b62=0.00003349750170860705
b03=-0.09043168323908105619; b13=-0.06764045190713819075; b23=0.06759613017704596460; b33=0.02955525704293155274; b63=-0.00001391802575160607
b24=-0.09233646193671185927; b34=-0.01693649390020817171; b64=-0.00001400867981820361
def t18(A):
    I=np.eye(A.shape[0])
    A2=A@A; A3=A2@A; A6=A3@A3
    B1=a11*A+a21*A2+a31*A3
    B2=b11*A+b21*A2+b31*A3+b61*A6
    B3=b02*I+b12*A+b22*A2+b32*A3+b62*A6
    B4=b03*I+b13*A+b23*A2+b33*A3+b63*A6
    B5=b24*A2+b34*A3+b64*A6
    A9=B1@B5+B4
    return B2+(B3+A9)@A9
def expm_t18(A, theta=1.09):
    n=np.abs(A).sum(axis=0).max()
    s=0
    while n/2**s>theta: s+=1
    X=t18(A/2**s)
    for _ in range(s): X=X@X
    return X,s
rng=np.random.default_rng(0)
for d in [3,9,27,81]:
  for scale in [0.01,0.3,0.8,1.0,1.09,1.3,2.7,6.0,25.0]:
    worst=0
    for herm in [True,False]:
      for t in range(3):
        h=rng.normal(size=(d,d))+1j*rng.normal(size=(d,d))
        if herm: h=h+h.conj().T; a=-1j*h
        else: a=h
        a*=scale/np.abs(a).sum(axis=0).max()
        X,s=expm_t18(a)
        ref=scipy.linalg.expm(a)
        worst=max(worst,np.linalg.norm(X-ref)/np.linalg.norm(ref))
    print(d,scale,"s",s,"worst rel err %.2e"%worst)
# no-scaling accuracy vs norm (theta check)
for nrm in [0.8,1.0,1.09,1.2,1.4,1.6,2.0]:
    h=rng.normal(size=(9,9))+1j*rng.normal(size=(9,9)); h=h+h.conj().T; a=-1j*h; a*=nrm/np.abs(a).sum(axis=0).max()
    print("unscaled norm",nrm,"err %.2e"%(np.linalg.norm(t18(a)-scipy.linalg.expm(a))/3))
