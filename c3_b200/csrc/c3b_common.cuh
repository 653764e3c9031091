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

}  // namespace c3b
