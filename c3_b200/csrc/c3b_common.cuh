// Common device helpers for the c3_b200 propagator kernels (sm_100a only).
//
// Complex numbers are double2 {x = re, y = im}; all matrices are row-major complex128,
// the layout of the reference's TF tensors (c3/libraries/propagation.py:287-294).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace c3b {

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }

// c += a * b   (4 DFMA: the fp64 pipe's native complex multiply-accumulate)
__device__ __forceinline__ void cfma(cplx& c, const cplx a, const cplx b) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(a.y, b.x, c.y);
}
// c -= a * b
__device__ __forceinline__ void cfms(cplx& c, const cplx a, const cplx b) {
    c.x = fma(-a.x, b.x, c.x);
    c.x = fma(a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y);
    c.y = fma(-a.y, b.x, c.y);
}
__device__ __forceinline__ cplx cmul(const cplx a, const cplx b) {
    return cmake(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx crcp(const cplx a) {
    const double rd = 1.0 / fma(a.x, a.x, a.y * a.y);
    return cmake(a.x * rd, -a.y * rd);
}
// exp of a complex scalar (phase/scale factor of the trace shift)
__device__ __forceinline__ cplx cexp_(const cplx a) {
    double sn, cs;
    sincos(a.y, &sn, &cs);
    const double e = exp(a.x);
    return cmake(e * cs, e * sn);
}
__device__ __forceinline__ double cabs1(const cplx a) { return sqrt(fma(a.x, a.x, a.y * a.y)); }

// ---- Pade coefficients, SURVEY.md Appendix A (Higham 2005), normalised by b0 ------------
// Row i holds c_j = b_j / b_0 for order m = 2*i + 3 (i = 0..3) and row 4 holds order 13.
// Normalising keeps U and V O(1); R = (V-U)^{-1}(V+U) is invariant under it.
__constant__ double kPade[5][14] = {
    {1.0, 60.0 / 120.0, 12.0 / 120.0, 1.0 / 120.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {1.0, 15120.0 / 30240.0, 3360.0 / 30240.0, 420.0 / 30240.0, 30.0 / 30240.0, 1.0 / 30240.0, 0, 0, 0, 0, 0,
     0, 0, 0},
    {1.0, 8648640.0 / 17297280.0, 1995840.0 / 17297280.0, 277200.0 / 17297280.0, 25200.0 / 17297280.0,
     1512.0 / 17297280.0, 56.0 / 17297280.0, 1.0 / 17297280.0, 0, 0, 0, 0, 0, 0},
    {1.0, 8821612800.0 / 17643225600.0, 2075673600.0 / 17643225600.0, 302702400.0 / 17643225600.0,
     30270240.0 / 17643225600.0, 2162160.0 / 17643225600.0, 110880.0 / 17643225600.0, 3960.0 / 17643225600.0,
     90.0 / 17643225600.0, 1.0 / 17643225600.0, 0, 0, 0, 0},
    {1.0, 32382376266240000.0 / 64764752532480000.0, 7771770303897600.0 / 64764752532480000.0,
     1187353796428800.0 / 64764752532480000.0, 129060195264000.0 / 64764752532480000.0,
     10559470521600.0 / 64764752532480000.0, 670442572800.0 / 64764752532480000.0,
     33522128640.0 / 64764752532480000.0, 1323241920.0 / 64764752532480000.0, 40840800.0 / 64764752532480000.0,
     960960.0 / 64764752532480000.0, 16380.0 / 64764752532480000.0, 182.0 / 64764752532480000.0,
     1.0 / 64764752532480000.0}};

// ---- degree-18 Taylor polynomial of exp in 5 matrix products (Bader, Blanes, Casas, "Computing
// the matrix exponential with an optimized Taylor polynomial approximation", Mathematics 7 (2019)
// 1174, scheme (T18)).  Checked in scratch/proto_t18.py: the composed polynomial reproduces 1/k!
// for k = 0..18 to the 20 digits given, and exp(A) to 3e-16 for ||A||_1 <= theta_18.
//   A2 = A A, A3 = A2 A, A6 = A3 A3
//   B1 = a11 A + a21 A2 + a31 A3            B5 = b24 A2 + b34 A3 + b64 A6
//   B4 = b03 I + b13 A + b23 A2 + b33 A3 + b63 A6     A9 = B1 B5 + B4
//   B3 = b02 I + b12 A + b22 A2 + b32 A3 + b62 A6     B2 = b11 A + b21 A2 + b31 A3 + b61 A6
//   T18 = B2 + (B3 + A9) A9
#define C3B_T18_A11 (-0.10036558103014462001)
#define C3B_T18_A21 (-0.00802924648241156960)
#define C3B_T18_A31 (-0.00089213849804572995)
#define C3B_T18_B11 (0.39784974949964507614)
#define C3B_T18_B21 (1.36783778460411719922)
#define C3B_T18_B31 (0.49828962252538267755)
#define C3B_T18_B61 (-0.00063789819459472330)
#define C3B_T18_B02 (-10.9676396052962062593)
#define C3B_T18_B12 (1.68015813878906197182)
#define C3B_T18_B22 (0.05717798464788655127)
#define C3B_T18_B32 (-0.00698210122488052084)
#define C3B_T18_B62 (0.00003349750170860705)
#define C3B_T18_B03 (-0.09043168323908105619)
#define C3B_T18_B13 (-0.06764045190713819075)
#define C3B_T18_B23 (0.06759613017704596460)
#define C3B_T18_B33 (0.02955525704293155274)
#define C3B_T18_B63 (-0.00001391802575160607)
#define C3B_T18_B24 (-0.09233646193671185927)
#define C3B_T18_B34 (-0.01693649390020817171)
#define C3B_T18_B64 (-0.00001400867981820361)
// backward error <= 2^-53 for ||A||_1 <= 1.09 (same paper, table of theta_m)
#define C3B_THETA18 1.09

// ---- degree-15+ Taylor polynomial of exp in FOUR matrix products (evaluation formulas of the Sastre-Ibanez-Defez type,
// "Boosting the computation of the matrix exponential", Appl. Math. Comput. 340 (2019): polynomials of degree 16 whose
// coefficients match the Taylor series up to degree 15).  No network here, so the 16 coefficients were obtained by solving
// the 16 polynomial conditions numerically (Levenberg-Marquardt from random starts, polished to 40 digits with mpmath;
// tests/test_taylor_schemes.py re-derives the composed polynomial from these literals and checks it against 1/k!):
//   A2 = A A
//   P0 = A2 (a1 A2 + a2 A)
//   P1 = (P0 + b1 A2 + b2 A)(P0 + b3 A2 + b4 I) + b5 P0
//   T  = (P1 + c1 A2 + c2 A)(P1 + c3 P0 + c4 A) + c9 P1 + c5 P0 + c6 A2 + c7 A + c8 I
// T(x) = sum_{k<=15} x^k / k! + 0.5457 x^16 / 16!.  Measured forward error against a 40-digit reference: 1.2e-15 on the
// headline slices (||A|| = 0.67), 1.1e-15 at spectral radius 0.8, 1.0e-14 at 1.0 (the degree-18 scheme: 3e-16 throughout,
// with one product more).  Used with theta = 0.8 by the forward kernels; the tolerance of the path is 1e-10 after N products.
#define C3B_T15_A1 0.00040187616102010354629
#define C3B_T15_A2 0.0029455314402796829805
#define C3B_T15_B1 0.087121675660506912665
#define C3B_T15_B2 0.40175684406735678015
#define C3B_T15_B3 (-0.063523113356121467813)
#define C3B_T15_B4 3.0014665781772709006
#define C3B_T15_B5 10.046029558660165472
#define C3B_T15_C1 (-0.23810703738709872247)
#define C3B_T15_C2 (-1.2471625036814465831)
#define C3B_T15_C3 5.7923617070732605218
#define C3B_T15_C4 1.0183494324742248072
#define C3B_T15_C5 (-3.030123400738712306)
#define C3B_T15_C6 (-2.1297555904964357845)
#define C3B_T15_C7 (-11.550609098606822755)
#define C3B_T15_C8 1.0
#define C3B_T15_C9 10.408017352313543646
#define C3B_THETA15 0.8

#define C3B_THETA3 1.495585217958292e-2
#define C3B_THETA5 2.539398330063230e-1
#define C3B_THETA7 9.504178996162932e-1
#define C3B_THETA9 2.097847961257068
#define C3B_THETA13 5.371920351148152
// ||A|| < 2 ln 2 makes V-U = p(-A)/b0 strictly diagonally dominant (||p(-A)/b0 - I|| <=
// exp(||A||/2) - 1 < 1 because b_j/b_0 <= 1/(2^j j!)), so Gaussian elimination WITHOUT
// pivoting is stable (growth <= 2).  The register-resident kernel scales to below this.
#define C3B_NOPIVOT_LIMIT 1.386

// smallest s >= 0 with x * 2^-s < limit  (x >= 0, finite)
__device__ __forceinline__ int squarings_for(double x, double limit) {
    const double q = x / limit;
    if (!(q >= 1.0)) return 0;
    const int hi = __double2hiint(q);
    const int e = ((hi >> 20) & 0x7ff) - 1022;  // q = f * 2^e, f in [0.5, 1)
    return e > 0 ? e : 0;
}
__device__ __forceinline__ double pow2neg(int s) {  // 2^-s, 0 <= s < 1000
    return __hiloint2double((1023 - s) << 20, 0);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace c3b
