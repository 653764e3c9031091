// Fused gradient kernel for closed systems with 16 < d <= 32 (BASELINE config 5: the d = 27 tunable coupler) and Hermitian
// Hamiltonians -- the scheme of grad_blk9.cuh on the CTA-cooperative DMMA product of pwc_gemm.cuh (SURVEY.md section 8f, f-1;
// replaces tf.GradientTape through tf_propagation_vectorized + tf_matmul_n, c3/optimizers/optimizer.py:210-215,
// c3/libraries/propagation.py:426-440):
//
//   dL/dc_k[n] = Re tr( L(A_n, Y_n) dU_n^dag G_k ),      Y_{n+1} = dU_n Y_n dU_n^dag,      Y_0 = Ubar^dag U
//
// No stored slice propagators and no sweeps: a CTA walks a chunk of CL consecutive slices of one batch row, Y at the head of
// the chunk comes from the chunk products of one forward launch (grad9_prefix_kernel + grad9_ybound_kernel).  Per slice the Frechet derivative of
// the four-product Taylor scheme in direction Y runs next to the scheme: 15 products (+ 3 per squaring) against 6 of a
// forward slice, two independent products per barrier where the data flow allows.
//
// SEVEN matrix slots (7 x 16 KB at DP = 32: two CTAs per SM, like the forward kernel; the stored-propagator Frechet kernel
// needs ten and runs one CTA per SM).  What makes seven enough: operands are overwritten as soon as they are dead, the own
// values of A2 / P0 (dA2 / dP0) that the later combinations need are recovered from L1, R1', A (dL1, dR1, Y) as in
// grad_blk9.cuh, and A itself is re-assembled from the generators instead of being kept.
#pragma once
#include "pwc_gemm.cuh"

namespace c3b {

constexpr int kGradUSlots = 7;

template <int TM, int TN, int DPT, int KST, int NT>
__global__ void __launch_bounds__(NT, 2) grad_unitary_cta_kernel(const GradUParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    constexpr bool SWZ = (DPT == 32);
    const int D = p.D, K = p.K;
    const int DP = DPT > 0 ? DPT : p.DP;
    const int LD = DPT > 0 ? (SWZ ? DPT : DPT + 4) : p.LD;
    const int PP = DP * LD, KP = (D + 3) & ~3, tid = threadIdx.x;
    const size_t dd = (size_t)D * D;
    cplx* mats = reinterpret_cast<cplx*>(smem_raw);
    for (int e = tid; e < kGradUSlots * PP; e += NT) mats[e] = cmake(0.0, 0.0);       // padding stays zero through every product
    __syncthreads();
    cplx* S[kGradUSlots];
#pragma unroll
    for (int i = 0; i < kGradUSlots; ++i) S[i] = mats + (size_t)i * PP;
    constexpr double kI = 1.0 / (C3B_T15_B1 - C3B_T15_B3);
    auto mm = [&](cplx* C, const cplx* A, const cplx* B) { cta_zgemm<TM, TN, DPT, KST, NT, 0>(C, A, B, DP, LD, KP); };
    auto mma = [&](cplx* C, const cplx* A, const cplx* B) { cta_zgemm<TM, TN, DPT, KST, NT, 2>(C, A, B, DP, LD, KP, C); };   // C += A B

    const long long total = (long long)p.B * p.Q;
    for (long long unit = blockIdx.x; unit < total; unit += gridDim.x) {
        const int b = (int)(unit / p.Q), q = (int)(unit - (long long)b * p.Q);
        const int n0 = q * p.CL, n_end = min(p.N, n0 + p.CL);
        const double* sig_b = p.signals + (size_t)b * K * p.N;
        cplx* const sY = S[1];                  // the running Y never moves
        {
            const cplx* Yg = p.Ybound + ((size_t)b * p.Q + q) * dd;
            for (int e = tid; e < D * D; e += NT) { const int i = e / D, j = e - i * D; sY[mat_idx<SWZ>(i, j, LD)] = Yg[e]; }
        }
        for (int n = n0; n < n_end; ++n) {
            // slot roles of this slice (the squarings swap some of them locally)
            cplx *sA = S[0], *f1 = S[2], *f2 = S[3], *f3 = S[4], *f4 = S[5], *f5 = S[6];
            // ---- scaling from the row-sum bound (every warp computes it) ------------------------------------------------------
            double nb = 0.0;
            for (int r = tid & 31; r < D; r += 32) {
                double v = p.RS[r];
                for (int k = 0; k < K; ++k) v = fma(fabs(__ldg(sig_b + (size_t)k * p.N + n)), p.RS[(size_t)(k + 1) * D + r], v);
                nb = fmax(nb, v);
            }
            const int s = squarings_for(warp_max(nb), C3B_THETA15);
            const double sc = pow2neg(s);
            // A_n / 2^s, element (i, j): re-assembled wherever a combination needs it (the generators sit in L1 / L2)
            auto a_elem = [&](const int e) {
                cplx v = p.G[e];
                for (int k = 0; k < K; ++k) {
                    const double c = __ldg(sig_b + (size_t)k * p.N + n);
                    const cplx gk = p.G[(size_t)(k + 1) * dd + e];
                    v.x = fma(c, gk.x, v.x);
                    v.y = fma(c, gk.y, v.y);
                }
                return cmake(v.x * sc, v.y * sc);
            };
            for (int e = tid; e < D * D; e += NT) { const int i = e / D, j = e - i * D; sA[mat_idx<SWZ>(i, j, LD)] = a_elem(e); }
            __syncthreads();
            // ---- A2 = A A -> f1;  dA2 = A Y + Y A -> f2 --------------------------------------------------------------------------
            mm(f1, sA, sA);
            mm(f2, sA, sY);
            __syncthreads();
            mma(f2, sY, sA);
            __syncthreads();
            // ---- Q0 = a1 A2 + a2 A -> f3,  dQ0 = a1 dA2 + a2 Y -> f4 ---------------------------------------------------------------
            for (int e = tid; e < D * D; e += NT) {
                const int i = e / D, j = e - i * D, x = mat_idx<SWZ>(i, j, LD);
                const cplx a = sA[x], a2 = f1[x], y = sY[x], da2 = f2[x];
                f3[x] = cmake(C3B_T15_A1 * a2.x + C3B_T15_A2 * a.x, C3B_T15_A1 * a2.y + C3B_T15_A2 * a.y);
                f4[x] = cmake(C3B_T15_A1 * da2.x + C3B_T15_A2 * y.x, C3B_T15_A1 * da2.y + C3B_T15_A2 * y.y);
            }
            __syncthreads();
            // ---- P0 = A2 Q0 -> sA (A is re-assembled from here on);  dP0 = dA2 Q0 + A2 dQ0 -> f5 -------------------------------------
            mm(sA, f1, f3);
            mm(f5, f2, f3);
            __syncthreads();
            mma(f5, f1, f4);
            __syncthreads();
            // ---- L1 -> f3, R1' -> f4, dL1 -> f1, dR1 -> f2 (in place over Q0, dQ0, A2, dA2) ----------------------------------------
            for (int e = tid; e < D * D; e += NT) {
                const int i = e / D, j = e - i * D, x = mat_idx<SWZ>(i, j, LD);
                const cplx a = a_elem(e), a2 = f1[x], p0 = sA[x], y = sY[x], da2 = f2[x], dp0 = f5[x];
                f3[x] = cmake(p0.x + C3B_T15_B1 * a2.x + C3B_T15_B2 * a.x, p0.y + C3B_T15_B1 * a2.y + C3B_T15_B2 * a.y);
                f4[x] = cmake(p0.x + C3B_T15_B3 * a2.x, p0.y + C3B_T15_B3 * a2.y);
                f1[x] = cmake(dp0.x + C3B_T15_B1 * da2.x + C3B_T15_B2 * y.x, dp0.y + C3B_T15_B1 * da2.y + C3B_T15_B2 * y.y);
                f2[x] = cmake(dp0.x + C3B_T15_B3 * da2.x, dp0.y + C3B_T15_B3 * da2.y);
            }
            __syncthreads();
            // ---- L1 R1' -> sA;  dL1 R1' + L1 dR1 -> f5 ------------------------------------------------------------------------------
            mm(sA, f3, f4);
            mm(f5, f1, f4);
            __syncthreads();
            mma(f5, f3, f2);
            __syncthreads();
            // ---- P1, dP1 and the last combinations, in place: L2 -> f3, R2 -> f4, E0 -> sA;  dL2 -> f1, dR2 -> f2, dE0 -> f5 -----------
            for (int e = tid; e < D * D; e += NT) {
                const int i = e / D, j = e - i * D, x = mat_idx<SWZ>(i, j, LD);
                const cplx a = a_elem(e), l1 = f3[x], r1 = f4[x], c = sA[x];
                const cplx a2 = cmake((l1.x - r1.x - C3B_T15_B2 * a.x) * kI, (l1.y - r1.y - C3B_T15_B2 * a.y) * kI);
                const cplx p0 = cmake(r1.x - C3B_T15_B3 * a2.x, r1.y - C3B_T15_B3 * a2.y);
                const cplx p1 = cmake(c.x + C3B_T15_B4 * l1.x + C3B_T15_B5 * p0.x, c.y + C3B_T15_B4 * l1.y + C3B_T15_B5 * p0.y);
                const double dg = (i == j) ? 1.0 : 0.0;
                f3[x] = cmake(p1.x + C3B_T15_C1 * a2.x + C3B_T15_C2 * a.x, p1.y + C3B_T15_C1 * a2.y + C3B_T15_C2 * a.y);
                f4[x] = cmake(p1.x + C3B_T15_C3 * p0.x + C3B_T15_C4 * a.x, p1.y + C3B_T15_C3 * p0.y + C3B_T15_C4 * a.y);
                sA[x] = cmake(C3B_T15_C9 * p1.x + C3B_T15_C5 * p0.x + C3B_T15_C6 * a2.x + C3B_T15_C7 * a.x + C3B_T15_C8 * dg,
                              C3B_T15_C9 * p1.y + C3B_T15_C5 * p0.y + C3B_T15_C6 * a2.y + C3B_T15_C7 * a.y);
                const cplx y = sY[x], dl1 = f1[x], dr1 = f2[x], dc = f5[x];
                const cplx da2 = cmake((dl1.x - dr1.x - C3B_T15_B2 * y.x) * kI, (dl1.y - dr1.y - C3B_T15_B2 * y.y) * kI);
                const cplx dp0 = cmake(dr1.x - C3B_T15_B3 * da2.x, dr1.y - C3B_T15_B3 * da2.y);
                const cplx dp1 = cmake(dc.x + C3B_T15_B4 * dl1.x + C3B_T15_B5 * dp0.x, dc.y + C3B_T15_B4 * dl1.y + C3B_T15_B5 * dp0.y);
                f1[x] = cmake(dp1.x + C3B_T15_C1 * da2.x + C3B_T15_C2 * y.x, dp1.y + C3B_T15_C1 * da2.y + C3B_T15_C2 * y.y);
                f2[x] = cmake(dp1.x + C3B_T15_C3 * dp0.x + C3B_T15_C4 * y.x, dp1.y + C3B_T15_C3 * dp0.y + C3B_T15_C4 * y.y);
                // dT of the scaled slice is linear in its direction Y / 2^s: the factor goes on every term of dT
                f5[x] = cmake((C3B_T15_C9 * dp1.x + C3B_T15_C5 * dp0.x + C3B_T15_C6 * da2.x + C3B_T15_C7 * y.x),
                              (C3B_T15_C9 * dp1.y + C3B_T15_C5 * dp0.y + C3B_T15_C6 * da2.y + C3B_T15_C7 * y.y));
            }
            __syncthreads();
            // ---- T = E0 + L2 R2 (in place on sA);  dT = dE0 + dL2 R2 + L2 dR2 (in place on f5) ----------------------------------------
            mma(sA, f3, f4);
            mma(f5, f1, f4);
            __syncthreads();
            mma(f5, f3, f2);
            __syncthreads();
            if (s > 0) {
                for (int e = tid; e < D * D; e += NT) { const int i = e / D, j = e - i * D, x = mat_idx<SWZ>(i, j, LD); f5[x].x *= sc; f5[x].y *= sc; }
                __syncthreads();
            }
            cplx *sT = sA, *sdT = f5;               // f1 .. f4 are free
            for (int i = 0; i < s; ++i) {           // (T, dT) <- (T T, dT T + T dT)
                mm(f1, sT, sT);
                mm(f2, sdT, sT);
                __syncthreads();
                mma(f2, sT, sdT);
                __syncthreads();
                cplx* t = sT; sT = f1; f1 = t;
                t = sdT; sdT = f2; f2 = t;
            }
            // ---- T^dag -> f3 ---------------------------------------------------------------------------------------------------------
            for (int e = tid; e < D * D; e += NT) {
                const int i = e / D, j = e - i * D;
                const cplx t = sT[mat_idx<SWZ>(j, i, LD)];
                f3[mat_idx<SWZ>(i, j, LD)] = cmake(t.x, -t.y);
            }
            __syncthreads();
            // ---- V = dT T^dag -> f4;  T Y -> f1 ---------------------------------------------------------------------------------------
            mm(f4, sdT, f3);
            mm(f1, sT, sY);
            __syncthreads();
            // ---- the next Y = (T Y) T^dag -> sY, and the K contractions Re tr(V (G_k + t_k I)) ------------------------------------------
            mm(sY, f1, f3);
            for (int k = 0; k < K; ++k) {
                const cplx tk = p.TR[k + 1];
                double acc = 0.0;
                for (int e = tid; e < D * D; e += NT) {
                    const int i = e / D, j = e - i * D;
                    const cplx v = f4[mat_idx<SWZ>(i, j, LD)];
                    cplx g = p.G[(size_t)(k + 1) * dd + (size_t)j * D + i];
                    if (i == j) { g.x += tk.x; g.y += tk.y; }
                    acc = fma(v.x, g.x, fma(-v.y, g.y, acc));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if ((tid & 31) == 0) red[tid >> 5] = acc;
                __syncthreads();
                if (tid == 0) {
                    double t = 0.0;
                    for (int w = 0; w < NT / 32; ++w) t += red[w];
                    p.grad[((size_t)b * K + k) * p.N + n] = t;
                }
                __syncthreads();
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

}  // namespace c3b
