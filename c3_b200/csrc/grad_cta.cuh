// Gradient of the propagators for the dimensions the warp-per-slice kernels of grad.cuh cannot hold (closed d > 16, Lindblad
// superoperators up to D = 81 and beyond): the same adjoint scheme -- backward sweep Psi_n, forward sweep M_n = F_n Psi_n,
// Frechet derivative of the degree-18 Taylor scheme on (X, dX) pairs, contraction with the control generators -- with every
// O(D^3) step on the CTA-cooperative DMMA product of pwc_gemm.cuh (cta_zgemm, 3M complex product on m8n8k4 tiles).
// Replaces tf.GradientTape through tf_propagation_vectorized / tf_propagation_lind + tf_matmul_n
// (c3/optimizers/optimizer.py:210-215; c3/libraries/propagation.py:426-440, 551-585) for these shapes.
//
// Matrices are zero-padded to DP = 8 ceil(D / 8) with leading dimension LD (DP + 4 in shared memory, DP in the per-CTA
// global workspace); the unpadded D x D arrays dUs / Psi / M in global memory are copied in and out element-wise.
#pragma once
#include "pwc_gemm.cuh"

namespace c3b {

struct GradCtaParams {
    const cplx* G;          // [K+1, D, D] trace-shifted generators (shared model)
    const double* RS;       // [K+1, D] row sums of |G_k|
    const cplx* TR;         // [K+1] trace shifts
    const double* signals;  // [B, K, N]
    const cplx* dUs;        // [B, N, D, D] slice propagators of the forward pass
    const cplx* Ubar;       // [B, D, D] cotangent of U
    cplx* PsiM;             // [B, N, D, D]: Psi_n after the backward sweep, M_n after the forward sweep
    double* alpha;          // [B] normalisation of the cotangent
    double* grad;           // [B, K, N]
    int B, K, N, D, DP, LD;
    cplx* ws;               // per-CTA workspace (global path) or null (shared memory)
    int slots;              // workspace slots per CTA
};

constexpr int kGradSweepSlots = 4;
constexpr int kGradFrechetSlots = 10;

__device__ __forceinline__ void pad_load(cplx* dst, const cplx* src, const int D, const int LD, const int tid, const int nt, const double scale = 1.0) {
    for (int e = tid; e < D * D; e += nt) {
        const int i = e / D, j = e - i * D;
        const cplx v = src[e];
        dst[i * LD + j] = cmake(v.x * scale, v.y * scale);
    }
}
__device__ __forceinline__ void pad_store(cplx* dst, const cplx* src, const int D, const int LD, const int tid, const int nt) {
    for (int e = tid; e < D * D; e += nt) {
        const int i = e / D, j = e - i * D;
        dst[e] = src[i * LD + j];
    }
}

template <int NT>
__device__ __forceinline__ cplx* grad_cta_slots(const GradCtaParams& p, unsigned char* smem_raw, const int PP) {
    cplx* mats = p.ws ? p.ws + (size_t)blockIdx.x * p.slots * PP : reinterpret_cast<cplx*>(smem_raw);
    for (int e = threadIdx.x; e < p.slots * PP; e += NT) mats[e] = cmake(0.0, 0.0);     // padding stays zero through every product
    __syncthreads();
    return mats;
}

// Backward sweep, one CTA per batch row: alpha_b = 1 / ||Ubar_b||_F, Psi_{N-1} = alpha Ubar^dag, Psi_{n-1} = Psi_n dU_n.
template <int TM, int TN, int NT>
__global__ void __launch_bounds__(NT) grad_suffix_cta_kernel(const GradCtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    const int D = p.D, DP = p.DP, LD = p.LD, PP = DP * LD, KP = (D + 3) & ~3, tid = threadIdx.x;
    const size_t dd = (size_t)D * D;
    cplx* mats = grad_cta_slots<NT>(p, smem_raw, PP);
    cplx *P = mats, *Y = mats + PP, *T = mats + 2 * PP;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const cplx* ub = p.Ubar + (size_t)b * dd;
        double nrm = 0.0;
        for (int e = tid; e < D * D; e += NT) { const cplx u = ub[e]; nrm = fma(u.x, u.x, fma(u.y, u.y, nrm)); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if ((tid & 31) == 0) red[tid >> 5] = nrm;
        __syncthreads();
        nrm = 0.0;
        for (int w = 0; w < NT / 32; ++w) nrm += red[w];
        const double a = nrm > 0.0 ? rsqrt(nrm) : 0.0;
        if (tid == 0) p.alpha[b] = a;
        for (int e = tid; e < D * D; e += NT) {          // P = alpha Ubar^dag
            const int i = e / D, j = e - i * D;
            const cplx u = ub[j * D + i];
            P[i * LD + j] = cmake(a * u.x, -a * u.y);
        }
        __syncthreads();
        for (int n = p.N - 1; n >= 0; --n) {
            pad_store(p.PsiM + ((size_t)b * p.N + n) * dd, P, D, LD, tid, NT);
            if (n > 0) {
                pad_load(Y, p.dUs + ((size_t)b * p.N + n) * dd, D, LD, tid, NT);
                __syncthreads();
                cta_zgemm<TM, TN, 0, 0, NT>(T, P, Y, DP, LD, KP);
                __syncthreads();
                cplx* t = P; P = T; T = t;
            }
        }
        __syncthreads();
    }
}

// Forward sweep, one CTA per batch row: F_0 = I, M_n = F_n Psi_n (written over Psi_n), F_{n+1} = dU_n F_n.
template <int TM, int TN, int NT>
__global__ void __launch_bounds__(NT) grad_prefix_cta_kernel(const GradCtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = p.D, DP = p.DP, LD = p.LD, PP = DP * LD, KP = (D + 3) & ~3, tid = threadIdx.x;
    const size_t dd = (size_t)D * D;
    cplx* mats = grad_cta_slots<NT>(p, smem_raw, PP);
    cplx *F = mats, *Y = mats + PP, *T = mats + 2 * PP, *M = mats + 3 * PP;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int e = tid; e < D * D; e += NT) { const int i = e / D, j = e - i * D; F[i * LD + j] = cmake(i == j ? 1.0 : 0.0, 0.0); }
        for (int n = 0; n < p.N; ++n) {
            cplx* psi = p.PsiM + ((size_t)b * p.N + n) * dd;
            pad_load(Y, psi, D, LD, tid, NT);
            __syncthreads();
            cta_zgemm<TM, TN, 0, 0, NT>(M, F, Y, DP, LD, KP);                 // M_n = F_n Psi_n
            __syncthreads();
            pad_store(psi, M, D, LD, tid, NT);
            if (n + 1 < p.N) {
                pad_load(Y, p.dUs + ((size_t)b * p.N + n) * dd, D, LD, tid, NT);
                __syncthreads();
                cta_zgemm<TM, TN, 0, 0, NT>(T, Y, F, DP, LD, KP);             // F_{n+1} = dU_n F_n
                __syncthreads();
                cplx* t = F; F = T; T = t;
            }
        }
        __syncthreads();
    }
}

// One CTA per (b, n): W_n = L(A_n, M_n), the Frechet derivative of the four-product Taylor scheme on (X, dX) pairs (the
// recurrences of grad_blk9.cuh), then grad[b,k,n] = (1 / alpha_b) Re tr(e^{mu_n} W_n G_k).  Ten matrix slots: X0..X4 values, Y0..Y4
// derivatives; the fused epilogues of cta_zgemm accumulate in place (C = A B + C).
template <int TM, int TN, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) grad_frechet_cta_kernel(const GradCtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    const int D = p.D, K = p.K, DP = p.DP, LD = p.LD, PP = DP * LD, KP = (D + 3) & ~3, RL = D * LD, tid = threadIdx.x;
    const size_t dd = (size_t)D * D;
    cplx* mats = grad_cta_slots<NT>(p, smem_raw, PP);
    cplx* X[5];
    cplx* Y[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { X[i] = mats + (size_t)i * PP; Y[i] = mats + (size_t)(5 + i) * PP; }
    const long long total = (long long)p.B * p.N;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = (int)(w / p.N), n = (int)(w - (long long)b * p.N);
        const double* sig_b = p.signals + (size_t)b * K * p.N;
        // ---- scaling from the row-sum bound (every warp computes it), trace shift ----------------------------------------
        double nb = 0.0;
        for (int r = tid & 31; r < D; r += 32) {
            double v = p.RS[r];
            for (int k = 0; k < K; ++k) v = fma(fabs(__ldg(sig_b + (size_t)k * p.N + n)), p.RS[(size_t)(k + 1) * D + r], v);
            nb = fmax(nb, v);
        }
        const int s = squarings_for(warp_max(nb), C3B_THETA15);
        const double sc = pow2neg(s);
        cplx mu = p.TR[0];
        for (int k = 0; k < K; ++k) {
            const double c = __ldg(sig_b + (size_t)k * p.N + n);
            mu.x = fma(c, p.TR[k + 1].x, mu.x);
            mu.y = fma(c, p.TR[k + 1].y, mu.y);
        }
        // ---- X0 = A_n / 2^s, Y0 = M_n / 2^s -------------------------------------------------------------------------------
        const cplx* Mg = p.PsiM + (size_t)w * dd;
        for (int e = tid; e < D * D; e += NT) {
            const int i = e / D, j = e - i * D;
            cplx v = p.G[e];
            for (int k = 0; k < K; ++k) {
                const double c = __ldg(sig_b + (size_t)k * p.N + n);
                const cplx gk = p.G[(size_t)(k + 1) * dd + e];
                v.x = fma(c, gk.x, v.x);
                v.y = fma(c, gk.y, v.y);
            }
            X[0][i * LD + j] = cmake(v.x * sc, v.y * sc);
            const cplx m = Mg[e];
            Y[0][i * LD + j] = cmake(m.x * sc, m.y * sc);
        }
        __syncthreads();
        // ---- the four-product scheme (c3b_common.cuh) and its derivative, two independent products per barrier where possible ----
        //   X0 = A, Y0 = M;   X1 = A2, Y1 = dA2 = A M + M A
        cta_zgemm<TM, TN, 0, 0, NT>(X[1], X[0], X[0], DP, LD, KP);
        cta_zgemm<TM, TN, 0, 0, NT>(Y[1], X[0], Y[0], DP, LD, KP);
        __syncthreads();
        cta_zgemm<TM, TN, 0, 0, NT, 2>(Y[1], Y[0], X[0], DP, LD, KP, Y[1]);
        __syncthreads();
        //   X2 = Q0 = a1 A2 + a2 A,  Y2 = dQ0
        for (int e = tid; e < RL; e += NT) {
            const cplx x1 = X[0][e], x2 = X[1][e], y1 = Y[0][e], y2 = Y[1][e];
            X[2][e] = cmake(C3B_T15_A1 * x2.x + C3B_T15_A2 * x1.x, C3B_T15_A1 * x2.y + C3B_T15_A2 * x1.y);
            Y[2][e] = cmake(C3B_T15_A1 * y2.x + C3B_T15_A2 * y1.x, C3B_T15_A1 * y2.y + C3B_T15_A2 * y1.y);
        }
        __syncthreads();
        //   X3 = P0 = A2 Q0,  Y3 = dP0 = dA2 Q0 + A2 dQ0
        cta_zgemm<TM, TN, 0, 0, NT>(X[3], X[1], X[2], DP, LD, KP);
        cta_zgemm<TM, TN, 0, 0, NT>(Y[3], Y[1], X[2], DP, LD, KP);
        __syncthreads();
        cta_zgemm<TM, TN, 0, 0, NT, 2>(Y[3], X[1], Y[2], DP, LD, KP, Y[3]);
        __syncthreads();
        //   X2 = L1 = P0 + b1 A2 + b2 A,  X4 = R1 = P0 + b3 A2 + b4 I,  X3 = b5 P0 (the addend of the next product); Y likewise
        for (int e = tid; e < RL; e += NT) {
            const int i = e / LD, j = e - i * LD;
            const double dg = (i == j) ? 1.0 : 0.0;
            const cplx x1 = X[0][e], x2 = X[1][e], p0 = X[3][e], y1 = Y[0][e], y2 = Y[1][e], q0 = Y[3][e];
            X[2][e] = cmake(p0.x + C3B_T15_B1 * x2.x + C3B_T15_B2 * x1.x, p0.y + C3B_T15_B1 * x2.y + C3B_T15_B2 * x1.y);
            X[4][e] = cmake(p0.x + C3B_T15_B3 * x2.x + C3B_T15_B4 * dg, p0.y + C3B_T15_B3 * x2.y);
            X[3][e] = cmake(C3B_T15_B5 * p0.x, C3B_T15_B5 * p0.y);
            Y[2][e] = cmake(q0.x + C3B_T15_B1 * y2.x + C3B_T15_B2 * y1.x, q0.y + C3B_T15_B1 * y2.y + C3B_T15_B2 * y1.y);
            Y[4][e] = cmake(q0.x + C3B_T15_B3 * y2.x, q0.y + C3B_T15_B3 * y2.y);
            Y[3][e] = cmake(C3B_T15_B5 * q0.x, C3B_T15_B5 * q0.y);
        }
        __syncthreads();
        //   X3 = P1 = L1 R1 + b5 P0,  Y3 = dP1 = dL1 R1 + L1 dR1 + b5 dP0   (in place on the addend)
        cta_zgemm<TM, TN, 0, 0, NT, 2>(X[3], X[2], X[4], DP, LD, KP, X[3]);
        cta_zgemm<TM, TN, 0, 0, NT, 2>(Y[3], Y[2], X[4], DP, LD, KP, Y[3]);
        __syncthreads();
        cta_zgemm<TM, TN, 0, 0, NT, 2>(Y[3], X[2], Y[4], DP, LD, KP, Y[3]);
        __syncthreads();
        //   P0 = L1 - b1 A2 - b2 A is recovered (|L1| ~ 0.4 |A|: the cancellation costs 1e-16 absolute);
        //   X2 = L2, X4 = R2, X3 = E0 and their derivatives in the Y slots
        for (int e = tid; e < RL; e += NT) {
            const int i = e / LD, j = e - i * LD;
            const double dg = (i == j) ? 1.0 : 0.0;
            const cplx x1 = X[0][e], x2 = X[1][e], l1 = X[2][e], p1 = X[3][e];
            const cplx y1 = Y[0][e], y2 = Y[1][e], dl1 = Y[2][e], q1 = Y[3][e];
            const cplx p0 = cmake(l1.x - C3B_T15_B1 * x2.x - C3B_T15_B2 * x1.x, l1.y - C3B_T15_B1 * x2.y - C3B_T15_B2 * x1.y);
            const cplx q0 = cmake(dl1.x - C3B_T15_B1 * y2.x - C3B_T15_B2 * y1.x, dl1.y - C3B_T15_B1 * y2.y - C3B_T15_B2 * y1.y);
            X[2][e] = cmake(p1.x + C3B_T15_C1 * x2.x + C3B_T15_C2 * x1.x, p1.y + C3B_T15_C1 * x2.y + C3B_T15_C2 * x1.y);
            X[4][e] = cmake(p1.x + C3B_T15_C3 * p0.x + C3B_T15_C4 * x1.x, p1.y + C3B_T15_C3 * p0.y + C3B_T15_C4 * x1.y);
            X[3][e] = cmake(C3B_T15_C9 * p1.x + C3B_T15_C5 * p0.x + C3B_T15_C6 * x2.x + C3B_T15_C7 * x1.x + C3B_T15_C8 * dg,
                            C3B_T15_C9 * p1.y + C3B_T15_C5 * p0.y + C3B_T15_C6 * x2.y + C3B_T15_C7 * x1.y);
            Y[2][e] = cmake(q1.x + C3B_T15_C1 * y2.x + C3B_T15_C2 * y1.x, q1.y + C3B_T15_C1 * y2.y + C3B_T15_C2 * y1.y);
            Y[4][e] = cmake(q1.x + C3B_T15_C3 * q0.x + C3B_T15_C4 * y1.x, q1.y + C3B_T15_C3 * q0.y + C3B_T15_C4 * y1.y);
            Y[3][e] = cmake(C3B_T15_C9 * q1.x + C3B_T15_C5 * q0.x + C3B_T15_C6 * y2.x + C3B_T15_C7 * y1.x,
                            C3B_T15_C9 * q1.y + C3B_T15_C5 * q0.y + C3B_T15_C6 * y2.y + C3B_T15_C7 * y1.y);
        }
        __syncthreads();
        //   X0 = T = L2 R2 + E0 (needed only to undo a scaling);  Y0 = dT = dL2 R2 + L2 dR2 + dE0
        if (s > 0) cta_zgemm<TM, TN, 0, 0, NT, 2>(X[0], X[2], X[4], DP, LD, KP, X[3]);
        cta_zgemm<TM, TN, 0, 0, NT, 2>(Y[0], Y[2], X[4], DP, LD, KP, Y[3]);
        __syncthreads();
        cta_zgemm<TM, TN, 0, 0, NT, 2>(Y[0], X[2], Y[4], DP, LD, KP, Y[0]);
        __syncthreads();
        cplx *Xc = X[0], *dXc = Y[0], *Xn = X[1], *dXn = Y[1];
        for (int q = 0; q < s; ++q) {                           // (X, dX) <- (X X, dX X + X dX)
            cta_zgemm<TM, TN, 0, 0, NT>(dXn, dXc, Xc, DP, LD, KP);
            if (q + 1 < s) cta_zgemm<TM, TN, 0, 0, NT>(Xn, Xc, Xc, DP, LD, KP);
            __syncthreads();
            cta_zgemm<TM, TN, 0, 0, NT, 2>(dXn, Xc, dXc, DP, LD, KP, dXn);
            __syncthreads();
            cplx* t = Xc; Xc = Xn; Xn = t;
            t = dXc; dXc = dXn; dXn = t;
        }
        // ---- contraction with the control generators: (1 / alpha) Re tr(e^mu dX (G_k + t_k I)) ---------------------------------------
        const cplx ph = cexp_(mu);
        const double a = p.alpha[b];
        const double inv_a = a > 0.0 ? 1.0 / a : 0.0;
        for (int k = 0; k < K; ++k) {
            const cplx tk = p.TR[k + 1];
            double acc = 0.0;
            for (int e = tid; e < D * D; e += NT) {
                const int i = e / D, j = e - i * D;
                const cplx wv = cmul(ph, dXc[i * LD + j]);
                cplx g = p.G[(size_t)(k + 1) * dd + (size_t)j * D + i];
                if (i == j) { g.x += tk.x; g.y += tk.y; }
                acc = fma(wv.x, g.x, fma(-wv.y, g.y, acc));     // Re(w_ij g_ji)
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if ((tid & 31) == 0) red[tid >> 5] = acc;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w2 = 0; w2 < NT / 32; ++w2) t += red[w2];
                p.grad[((size_t)b * K + k) * p.N + n] = t * inv_a;
            }
            __syncthreads();
        }
    }
}

}  // namespace c3b
