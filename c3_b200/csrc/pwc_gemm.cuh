// CTA-cooperative PWC propagator kernel for larger dimensions (closed d > 12, Lindblad D = d^2 up to
// 81 and beyond), built on fp64 tensor-core tiles.
//
// Per slice: A_n (trace-shifted generators, exact 1-norm -> s) -> degree-18 Taylor scheme in 5 complex
// products (c3b_common.cuh) -> s squarings -> running ordered product.  Every O(D^3) step is ONE
// routine, cta_zgemm: C = A B on zero-padded DP x DP complex matrices (DP = D rounded up to 8), each
// warp owning 2x2 (or 1x2) m8n8 tiles and issuing mma.sync.m8n8k4.f64 (DMMA): a complex tile step is
// 4 real DMMAs (Cr += Ar Br - Ai Bi, Ci += Ar Bi + Ai Br) on fragments loaded straight from the
// interleaved complex storage (one 16-byte load = re and im of a fragment element).  One operand
// load feeds 8 complex MACs per lane (vs 1-1.5 in the register kernels), so the D^2 x D^2 Lindblad
// superoperator is a genuinely dense contraction on the fp64 pipe, as BASELINE.json's north_star
// asks.  Matrices live in shared memory (DP <= 32) or in a per-CTA global workspace that stays in
// L1/L2 (D = 81: 10 x 124 KB).  Replaces c3/libraries/propagation.py:426-440,551-585 and
// c3/utils/tf_utils.py:120-193 for these shapes; no linear solve, no pivoting.
#pragma once
#include "c3b_common.cuh"
#include "pwc_cta.cuh"   // CtaParams, kCtaThreads

namespace c3b {

constexpr int kGemmSlots = 10;   // S0..S8 scratch + P

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// C = A * B for zero-padded DP x DP complex matrices (row-major, leading dimension DP, DP % 8 == 0).
// Warp w owns macro tiles of TM x TN m8n8 blocks, assigned round-robin.  No __restrict__: operands
// may be global-workspace buffers written earlier by this CTA.
// LD is the leading dimension (>= DP); KP = D rounded up to 4 bounds the k loop (columns/rows beyond D
// are zero).  In shared memory LD = DP + 4 (= 4 mod 8 in 16-byte units) makes the A-fragment loads
// bank-conflict free and the B-fragment loads 2-way (with LD = DP = 32 they were 8-way / 4-way).
template <int TM, int TN>
__device__ __forceinline__ void cta_zgemm(cplx* C, const cplx* A, const cplx* B, const int DP, const int LD, const int KP) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kCtaThreads / 32;
    const int nb = DP >> 3;                       // m8n8 blocks per dimension
    const int mt_r = (nb + TM - 1) / TM, mt_c = (nb + TN - 1) / TN;
    const int fr = lane >> 2, fc = lane & 3;      // fragment coordinates
    for (int mt = warp; mt < mt_r * mt_c; mt += NW) {
        const int bi0 = (mt / mt_c) * TM, bj0 = (mt % mt_c) * TN;
        double cr[TM][TN][2], ci[TM][TN][2];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }
        // clamp block indices of partial macro tiles (results of clamped duplicates are not stored)
        int arow[TM], bcol[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) arow[i] = (min(bi0 + i, nb - 1) * 8 + fr) * LD + fc;
#pragma unroll
        for (int j = 0; j < TN; ++j) bcol[j] = fc * LD + min(bj0 + j, nb - 1) * 8 + fr;
#pragma unroll 2
        for (int k0 = 0; k0 < KP; k0 += 4) {
            cplx a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = A[arow[i] + k0];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = B[bcol[j] + k0 * LD];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    dmma8x8x4(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
                    dmma8x8x4(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
                    dmma8x8x4(cr[i][j][0], cr[i][j][1], -a[i].y, b[j].y);
                    dmma8x8x4(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
                }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                if (bi0 + i < nb && bj0 + j < nb) {
                    cplx* o = C + ((bi0 + i) * 8 + fr) * LD + (bj0 + j) * 8 + 2 * fc;
                    o[0] = cmake(cr[i][j][0], ci[i][j][0]);
                    o[1] = cmake(cr[i][j][1], ci[i][j][1]);
                }
            }
    }
}

// exact 1-norm over the D x D part of a DP-strided matrix
__device__ __forceinline__ double cta_norm1_ld(const cplx* A, const int D, const int DP, double* red) {
    double best = 0.0;
    for (int c = threadIdx.x; c < D; c += kCtaThreads) {
        double s = 0.0;
        for (int i = 0; i < D; ++i) s += cabs1(A[i * DP + c]);
        best = fmax(best, s);
    }
    best = warp_max(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    double v = red[0];
#pragma unroll
    for (int w = 1; w < kCtaThreads / 32; ++w) v = fmax(v, red[w]);
    __syncthreads();
    return v;
}

struct GemmParams {
    CtaParams c;       // same fields as the Pade CTA kernel (G, signals, hlist, sizes, outputs, ws, use_smem)
    const cplx* TR;    // [(Bm), K+1] trace shifts already subtracted from G's diagonals, or null
    int DP;            // D rounded up to a multiple of 8 (tile extent)
    int LD;            // leading dimension of every workspace matrix (DP, or DP + 4 in shared memory)
};

template <int TM, int TN>
__global__ void __launch_bounds__(kCtaThreads) pwc_t18_cta_kernel(const GemmParams gp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[kCtaThreads / 32];
    const CtaParams& p = gp.c;
    const int D = p.D, K = p.K, DP = gp.DP, LD = gp.LD;
    const int KP = (D + 3) & ~3;
    const int PP = DP * LD;
    const int tid = threadIdx.x;

    cplx* mats = p.use_smem ? reinterpret_cast<cplx*>(smem_raw) : p.ws + (size_t)blockIdx.x * kGemmSlots * PP;
    cplx* S[kGemmSlots];
#pragma unroll
    for (int i = 0; i < kGemmSlots; ++i) S[i] = mats + (size_t)i * PP;
    cplx* const P = S[9];
    const bool shifted = gp.TR != nullptr && p.hlist == nullptr;

    // zero everything once: the padding rows/columns stay zero through every product
    for (int e = tid; e < kGemmSlots * PP; e += kCtaThreads) mats[e] = cmake(0.0, 0.0);
    __syncthreads();

    const long long units = (long long)p.B * p.S;
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int b = (int)(unit / p.S);
        const int sidx = (int)(unit - (long long)b * p.S);
        const int n_begin = sidx * p.seg_len;
        const int n_end = min(p.N, n_begin + p.seg_len);
        const cplx* Gb = p.G ? p.G + (size_t)b * p.model_stride : nullptr;
        const cplx* TRb = shifted ? gp.TR + (size_t)b * (p.model_stride ? (K + 1) : 0) : nullptr;
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        const cplx hs = cmake(p.hscale_re, p.hscale_im);
        cplx mu_acc = cmake(0.0, 0.0);

        for (int n = n_begin; n < n_end; ++n) {
            cplx* A = S[0];
            // ---- assemble (D x D part; padding stays zero) ---------------------------------------
            if (p.hlist == nullptr) {
                for (int e = tid; e < D * D; e += kCtaThreads) {
                    const int i = e / D, j = e - i * D;
                    cplx v = Gb[e];
                    for (int k = 0; k < K; ++k) {
                        const double c = __ldg(sig_b + (size_t)k * p.N + n);
                        const cplx gk = Gb[(size_t)(k + 1) * D * D + e];
                        v.x = fma(c, gk.x, v.x);
                        v.y = fma(c, gk.y, v.y);
                    }
                    A[i * LD + j] = v;
                }
                if (shifted) {
                    cplx mu = TRb[0];
                    for (int k = 0; k < K; ++k) {
                        const double c = __ldg(sig_b + (size_t)k * p.N + n);
                        mu.x = fma(c, TRb[k + 1].x, mu.x);
                        mu.y = fma(c, TRb[k + 1].y, mu.y);
                    }
                    mu_acc.x += mu.x; mu_acc.y += mu.y;
                }
            } else {
                const cplx* H = p.hlist + ((size_t)b * p.N + n) * D * D;
                for (int e = tid; e < D * D; e += kCtaThreads) {
                    const int i = e / D, j = e - i * D;
                    A[i * LD + j] = cmul(hs, H[e]);
                }
            }
            __syncthreads();
            const double nrm = cta_norm1_ld(A, D, LD, red);
            const int s = squarings_for(nrm, C3B_THETA18);
            if (s > 0) {
                const double sc = pow2neg(s);
                for (int e = tid; e < PP; e += kCtaThreads) { A[e].x *= sc; A[e].y *= sc; }
                __syncthreads();
            }
            // ---- T18: A2 = S1, A3 = S2, A6 = S3 -----------------------------------------------------
            cta_zgemm<TM, TN>(S[1], A, A, DP, LD, KP);
            __syncthreads();
            cta_zgemm<TM, TN>(S[2], S[1], A, DP, LD, KP);
            __syncthreads();
            cta_zgemm<TM, TN>(S[3], S[2], S[2], DP, LD, KP);
            __syncthreads();
            // B1 -> S4, B5 -> S5, B4 -> S6, B3 -> S7, B2 -> S8
            for (int e = tid; e < PP; e += kCtaThreads) {
                const int i = e / LD, j = e - i * LD;
                const double dg = (i == j && i < D) ? 1.0 : 0.0;
                const cplx x1 = A[e], x2 = S[1][e], x3 = S[2][e], x6 = S[3][e];
                S[4][e] = cmake(C3B_T18_A11 * x1.x + C3B_T18_A21 * x2.x + C3B_T18_A31 * x3.x,
                                C3B_T18_A11 * x1.y + C3B_T18_A21 * x2.y + C3B_T18_A31 * x3.y);
                S[5][e] = cmake(C3B_T18_B24 * x2.x + C3B_T18_B34 * x3.x + C3B_T18_B64 * x6.x,
                                C3B_T18_B24 * x2.y + C3B_T18_B34 * x3.y + C3B_T18_B64 * x6.y);
                S[6][e] = cmake(C3B_T18_B03 * dg + C3B_T18_B13 * x1.x + C3B_T18_B23 * x2.x + C3B_T18_B33 * x3.x + C3B_T18_B63 * x6.x,
                                C3B_T18_B13 * x1.y + C3B_T18_B23 * x2.y + C3B_T18_B33 * x3.y + C3B_T18_B63 * x6.y);
                S[7][e] = cmake(C3B_T18_B02 * dg + C3B_T18_B12 * x1.x + C3B_T18_B22 * x2.x + C3B_T18_B32 * x3.x + C3B_T18_B62 * x6.x,
                                C3B_T18_B12 * x1.y + C3B_T18_B22 * x2.y + C3B_T18_B32 * x3.y + C3B_T18_B62 * x6.y);
                S[8][e] = cmake(C3B_T18_B11 * x1.x + C3B_T18_B21 * x2.x + C3B_T18_B31 * x3.x + C3B_T18_B61 * x6.x,
                                C3B_T18_B11 * x1.y + C3B_T18_B21 * x2.y + C3B_T18_B31 * x3.y + C3B_T18_B61 * x6.y);
            }
            __syncthreads();
            cta_zgemm<TM, TN>(S[1], S[4], S[5], DP, LD, KP);      // B1 B5
            __syncthreads();
            for (int e = tid; e < PP; e += kCtaThreads) {  // A9 -> S2, B3 + A9 -> S3
                const cplx a9 = cmake(S[1][e].x + S[6][e].x, S[1][e].y + S[6][e].y);
                S[2][e] = a9;
                S[3][e] = cmake(S[7][e].x + a9.x, S[7][e].y + a9.y);
            }
            __syncthreads();
            cta_zgemm<TM, TN>(S[1], S[3], S[2], DP, LD, KP);      // (B3 + A9) A9
            __syncthreads();
            cplx* X = S[4];
            for (int e = tid; e < PP; e += kCtaThreads) X[e] = cmake(S[1][e].x + S[8][e].x, S[1][e].y + S[8][e].y);
            __syncthreads();
            for (int i = 0; i < s; ++i) {                  // undo the scaling
                cplx* nxt = (X == S[4]) ? S[5] : S[4];
                cta_zgemm<TM, TN>(nxt, X, X, DP, LD, KP);
                __syncthreads();
                X = nxt;
            }
            if (p.dUs_out != nullptr) {
                cplx* o = p.dUs_out + ((size_t)b * p.N + n) * D * D;
                cplx phn = cmake(1.0, 0.0);
                if (shifted) {
                    cplx mu = TRb[0];
                    for (int k = 0; k < K; ++k) {
                        const double c = __ldg(sig_b + (size_t)k * p.N + n);
                        mu.x = fma(c, TRb[k + 1].x, mu.x);
                        mu.y = fma(c, TRb[k + 1].y, mu.y);
                    }
                    phn = cexp_(mu);
                }
                for (int e = tid; e < D * D; e += kCtaThreads) {
                    const int i = e / D, j = e - i * D;
                    o[e] = cmul(phn, X[i * LD + j]);
                }
            }
            if (n == n_begin) {
                for (int e = tid; e < PP; e += kCtaThreads) P[e] = X[e];
            } else {
                cta_zgemm<TM, TN>(S[6], X, P, DP, LD, KP);
                __syncthreads();
                for (int e = tid; e < PP; e += kCtaThreads) P[e] = S[6][e];
            }
            __syncthreads();
        }
        cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * D * D) : (p.seg_out + ((size_t)b * p.S + sidx) * D * D);
        if (n_end > n_begin) {
            const cplx phu = shifted ? cexp_(mu_acc) : cmake(1.0, 0.0);
            for (int e = tid; e < D * D; e += kCtaThreads) {
                const int i = e / D, j = e - i * D;
                o[e] = cmul(phu, P[i * LD + j]);
            }
        } else {
            for (int e = tid; e < D * D; e += kCtaThreads) o[e] = cmake((e / D) == (e % D) ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
    }
}

}  // namespace c3b
