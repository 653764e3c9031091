"""How the coefficients of the four-product degree-15+ Taylor scheme (C3B_T15_* in c3_b200/csrc/c3b_common.cuh) were found.

    A2 = A A
    P0 = A2 (a1 A2 + a2 A)
    P1 = (P0 + b1 A2 + b2 A)(P0 + b3 A2 + b4 I) + b5 P0
    T  = (P1 + c1 A2 + c2 A)(P1 + c3 P0 + c4 A) + c9 P1 + c5 P0 + c6 A2 + c7 A + c8 I

T has degree 16 in A; 16 free coefficients are matched to 1/k!, k = 0..15 (a polynomial system with several isolated
real solutions; the x^16 coefficient of a solution is whatever it is -- 0.5457/16! for the one shipped).  The family of
evaluation formulas is the one of Sastre, Ibanez, Defez (2019) "Boosting the computation of the matrix exponential";
the coefficients here were solved from scratch:

    python scratch/proto_t15.py search SEED     Levenberg-Marquardt from random starts, prints distinct real solutions
                                                with their error at |A| = 0.7 / 0.8 and the largest coefficient
    python scratch/proto_t15.py polish          Newton polish of the candidates in 50-digit arithmetic (mpmath) and an
                                                accuracy table against scipy's expm

The shipped solution is the candidate with the smallest coefficients (max |c| = 11.6: no cancellation in fp64) and
relative error < 2.3e-16-ish at ||A||_inf <= 0.8 (C3B_THETA15); tests/test_taylor_schemes.py re-checks the 16 order
conditions with exact rationals and the accuracy inside theta on every run.
"""
import sys


def search(argv):
    import numpy as np, math, sys, json
    from numpy.polynomial import polynomial as P
    from scipy.optimize import least_squares
    import scipy.linalg
    seed=int(argv[0]) if argv else 0
    def poly_p2(v):
        a1,a2,b1,b2,b3,b4,b5,c1,c2,c3,c4,c5,c6,c7,c8,c9 = v
        x = np.array([0.0,1.0]); x2 = P.polymul(x,x)
        p0 = P.polymul(x2, P.polyadd(a1*x2, a2*x))
        p1 = P.polyadd(P.polymul(P.polyadd(p0, P.polyadd(b1*x2, b2*x)), P.polyadd(p0, P.polyadd(b3*x2, [b4]))), b5*p0)
        p2 = P.polymul(P.polyadd(p1, P.polyadd(c1*x2, c2*x)), P.polyadd(p1, P.polyadd(c3*p0, c4*x)))
        for t in (c9*p1, c5*p0, c6*x2, c7*x, [c8]): p2 = P.polyadd(p2, t)
        out = np.zeros(17); out[:len(p2)] = p2
        return out
    target = np.array([1.0/math.factorial(k) for k in range(16)]); scale = np.array([math.factorial(k) for k in range(16)], dtype=float)
    def resid(v): return (poly_p2(v)[:16] - target) * scale
    def scheme(A, v):
        a1,a2,b1,b2,b3,b4,b5,c1,c2,c3,c4,c5,c6,c7,c8,c9 = v
        I=np.eye(A.shape[0]); A2=A@A
        p0=A2@(a1*A2+a2*A)
        p1=(p0+b1*A2+b2*A)@(p0+b3*A2+b4*I)+b5*p0
        return (p1+c1*A2+c2*A)@(p1+c3*p0+c4*A)+c9*p1+c5*p0+c6*A2+c7*A+c8*I
    rng=np.random.default_rng(seed)
    mats=[]
    r2=np.random.default_rng(123)
    for rho in (0.7,0.8):
        for t in range(6):
            A=1j*np.diag(r2.uniform(-rho,rho,9)); A[0,0]=1j*rho; A[1,1]=-1j*rho
            B=0.03*(r2.normal(size=(9,9))+1j*r2.normal(size=(9,9))); A=A+(B-B.conj().T)/2
            mats.append((rho,A,scipy.linalg.expm(A)))
    found=[]
    for trial in range(3000):
        v0 = rng.normal(size=16) * rng.choice([0.01, 0.1, 1.0, 3.0], size=16)
        try: r = least_squares(resid, v0, method='lm', xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=3000)
        except Exception: continue
        if np.abs(r.fun).max() < 1e-10 and not any(np.allclose(r.x, f[0], rtol=1e-5, atol=1e-7) for f in found):
            errs={0.7:0.0,0.8:0.0}
            for rho,A,E in mats: errs[rho]=max(errs[rho], np.linalg.norm(scheme(A,r.x)-E)/np.linalg.norm(E))
            c16=poly_p2(r.x)[16]*math.factorial(16)
            found.append((r.x, errs, c16))
            print(json.dumps({"seed":seed,"e07":errs[0.7],"e08":errs[0.8],"c16":c16,"maxc":float(np.abs(r.x).max()),"v":r.x.tolist()}), flush=True)


def polish():
    import numpy as np, math, mpmath as mp, scipy.linalg
    mp.mp.dps = 50
    cands = [
     [ 4.01876161e-04,  2.94553144e-03,  8.71216757e-02,  4.01756844e-01, -6.35231134e-02,  3.00146658e+00,  1.00460296e+01, -2.38107037e-01, -1.24716250e+00,  5.79236171e+00,  1.01834943e+00, -3.03012340e+00, -2.12975559e+00, -1.15506091e+01,  1.00000000e+00,  1.04080174e+01],
     [ 4.01876161e-04,  2.94553144e-03, -8.70906658e-03,  4.01756844e-01,  3.23076289e-02, -2.68522007e+00,  8.45420858e+00,  2.38107037e-01,  3.30301471e+00, -5.79236171e+00,  1.03750278e+00, -6.33171246e+01,  3.48466586e-01,  1.22282268e+01,  1.00000000e+00,  1.04080174e+01],
     [-6.40007200e-04, -7.29265753e-04,  4.54646091e-02, -2.38132119e-01, -6.48496635e-02,  6.11623950e-01, -1.07858512e+01,  3.30337023e-01,  4.04806611e+00,  5.53864011e+00, -8.22843235e-01,  1.64754961e+02,  3.64071244e+00,  4.34557560e+00,  1.00000000e+00,  2.29703910e+01],
    ]
    def polys(v):
        a1,a2,b1,b2,b3,b4,b5,c1,c2,c3,c4,c5,c6,c7,c8,c9 = v
        def mul(p,q):
            r=[mp.mpf(0)]*(len(p)+len(q)-1)
            for i,a in enumerate(p):
                for j,b in enumerate(q): r[i+j]+=a*b
            return r
        def add(*ps):
            n=max(len(p) for p in ps); r=[mp.mpf(0)]*n
            for p in ps:
                for i,a in enumerate(p): r[i]+=a
            return r
        def sc(c,p): return [c*a for a in p]
        x=[mp.mpf(0),mp.mpf(1)]; x2=mul(x,x)
        p0=mul(x2, add(sc(a1,x2), sc(a2,x)))
        p1=add(mul(add(p0, sc(b1,x2), sc(b2,x)), add(p0, sc(b3,x2), [b4])), sc(b5,p0))
        p2=add(mul(add(p1, sc(c1,x2), sc(c2,x)), add(p1, sc(c3,p0), sc(c4,x))), sc(c9,p1), sc(c5,p0), sc(c6,x2), sc(c7,x), [c8])
        return p2
    def F(*v):
        p2=polys(v)
        return [ (p2[k] - 1/mp.factorial(k))*mp.factorial(k) for k in range(16)]
    def scheme(A, v):
        a1,a2,b1,b2,b3,b4,b5,c1,c2,c3,c4,c5,c6,c7,c8,c9 = [float(t) for t in v]
        I=np.eye(A.shape[0]); A2=A@A
        p0=A2@(a1*A2+a2*A)
        p1=(p0+b1*A2+b2*A)@(p0+b3*A2+b4*I)+b5*p0
        return (p1+c1*A2+c2*A)@(p1+c3*p0+c4*A)+c9*p1+c5*p0+c6*A2+c7*A+c8*I
    rng=np.random.default_rng(0)
    for ci,c in enumerate(cands):
        sol=mp.findroot(F, [mp.mpf(t) for t in c], tol=1e-40, maxsteps=200)
        v=[sol[i] for i in range(16)]
        p2=polys(v)
        print("cand",ci,"max|c|",float(max(abs(t) for t in v)),"x^16 coeff * 16! =", float(p2[16]*mp.factorial(16)))
        print("  ", [mp.nstr(t,20) for t in v])
        for nrm in (0.3,0.75,1.0,1.5):
            errs=[]
            for t in range(20):
                d=9; H=rng.normal(size=(d,d))+1j*rng.normal(size=(d,d)); H=H+H.conj().T
                A=-1j*H; A*=nrm/np.abs(A).sum(axis=1).max()
                E=scipy.linalg.expm(A); errs.append(np.linalg.norm(scheme(A,v)-E)/np.linalg.norm(E))
            print(f"   inf-norm {nrm}: max rel err {max(errs):.2e}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "search":
        search(sys.argv[2:])
    else:
        polish()
