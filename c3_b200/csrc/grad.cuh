// Gradient of a real scalar loss through the closed-system PWC propagator with respect to the control
// fields (what tf.GradientTape provides to the reference's default
// L-BFGS optimiser, c3/optimizers/optimizer.py:210-215, 277-313).
//
//   U = dU_{N-1} ... dU_0,  dU_n = exp(A_n),  A_n = G_0 + sum_k c_k[n] G_k,  G_k = -i dt H_k
//   dL = Re tr(Ubar^dag dU)                       (Ubar = torch's grad_output for U)
//   dL/dc_k[n] = Re tr( M_n  L(A_n, G_k) ),      M_n = F_n Ubar^dag R_n,
//                F_n = dU_{n-1} ... dU_0,  R_n = dU_{N-1} ... dU_{n+1},  L = Frechet derivative of exp.
// Since tr(M L(A, E)) = tr(L(A, M) E)  (L(A,E) = int_0^1 e^{sA} E e^{(1-s)A} ds), ONE Frechet derivative
// per slice, in direction M_n, serves all K controls:  dL/dc_k[n] = Re tr(W_n G_k),  W_n = L(A_n, M_n),
// and W_n is the upper-right block of exp([[A_n, M_n], [0, A_n]]) -- computed by the SAME fused
// exponential kernels on explicit 2d x 2d matrices (H-list mode), so no second expm implementation
// exists.  This file holds the three small kernels around that call; the orchestration (chunking over
// the batch to bound memory) is c3b_pwc_closed_grad in c3b_api.cu.
#pragma once
#include "c3b_common.cuh"

namespace c3b {

// one warp computes C = X * Y (d x d, row-major, shared memory), lanes strided over the entries
__device__ __forceinline__ void warp_mm(cplx* __restrict__ C, const cplx* __restrict__ X, const cplx* __restrict__ Y,
                                        const int d, const int lane) {
    for (int e = lane; e < d * d; e += 32) {
        const int i = e / d, j = e - i * d;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k < d; ++k) cfma(acc, X[i * d + k], Y[k * d + j]);
        C[e] = acc;
    }
}

// Psi_n = alpha Ubar^dag R_n for every n:  Psi_{N-1} = alpha Ubar^dag,  Psi_{n-1} = Psi_n dU_n.
// alpha_b = 1 / ||Ubar_b||_F keeps the augmented matrices O(1); it is divided out in grad_contract.
// One warp per batch row; dynamic smem = warps * 3 * d*d * 16 bytes.
__global__ void grad_suffix_kernel(const cplx* __restrict__ dUs, const cplx* __restrict__ Ubar, cplx* __restrict__ Psi,
                                   double* __restrict__ alpha, const int B, const int N, const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int b = blockIdx.x * wpb + warp;
    if (b >= B) return;
    const int dd = d * d;
    cplx* P = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 3 * dd;
    cplx* Y = P + dd;
    cplx* T = P + 2 * dd;
    double nrm = 0.0;
    for (int e = lane; e < dd; e += 32) { const cplx u = Ubar[(size_t)b * dd + e]; nrm = fma(u.x, u.x, fma(u.y, u.y, nrm)); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    const double a = nrm > 0.0 ? rsqrt(nrm) : 0.0;
    if (lane == 0) alpha[b] = a;
    for (int e = lane; e < dd; e += 32) {   // P = alpha * Ubar^dag
        const int i = e / d, j = e - i * d;
        const cplx u = Ubar[(size_t)b * dd + j * d + i];
        P[e] = cmake(a * u.x, -a * u.y);
    }
    __syncwarp();
    for (int n = N - 1; n >= 0; --n) {
        cplx* out = Psi + ((size_t)b * N + n) * dd;
        const cplx* dU = dUs + ((size_t)b * N + n) * dd;
        for (int e = lane; e < dd; e += 32) { out[e] = P[e]; Y[e] = dU[e]; }
        __syncwarp();
        if (n > 0) {
            warp_mm(T, P, Y, d, lane);
            __syncwarp();
            cplx* t = P; P = T; T = t;
        }
    }
}

// Forward sweep: F_0 = I, M_n = F_n Psi_n, F_{n+1} = dU_n F_n, and the augmented explicit "Hamiltonian"
//   Haug[b,n] = [[H_n, M_n / hscale], [0, H_n]],   H_n = h0 + sum_k c_k[n] h_k,   hscale = -i dt,
// so that exp(hscale * Haug) = [[dU_n, W_n], [0, dU_n]].
__global__ void grad_prefix_kernel(const cplx* __restrict__ dUs, const cplx* __restrict__ Psi, const cplx* __restrict__ h0,
                                   const cplx* __restrict__ hks, const double* __restrict__ signals, cplx* __restrict__ Haug,
                                   const double dt, const int B, const int K, const int N, const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int b = blockIdx.x * wpb + warp;
    if (b >= B) return;
    const int dd = d * d, d2 = 2 * d;
    cplx* F = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 4 * dd;
    cplx* Y = F + dd;
    cplx* T = F + 2 * dd;
    cplx* M = F + 3 * dd;
    for (int e = lane; e < dd; e += 32) F[e] = cmake((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
    __syncwarp();
    const double idt = 1.0 / dt;
    for (int n = 0; n < N; ++n) {
        const cplx* psi = Psi + ((size_t)b * N + n) * dd;
        const cplx* dU = dUs + ((size_t)b * N + n) * dd;
        for (int e = lane; e < dd; e += 32) Y[e] = psi[e];
        __syncwarp();
        warp_mm(M, F, Y, d, lane);                     // M_n = F_n Psi_n
        __syncwarp();
        cplx* out = Haug + ((size_t)b * N + n) * d2 * d2;
        for (int e = lane; e < dd; e += 32) {
            const int i = e / d, j = e - i * d;
            cplx h = h0[e];
            for (int k = 0; k < K; ++k) {
                const double c = __ldg(signals + ((size_t)b * K + k) * N + n);
                const cplx hk = hks[(size_t)k * dd + e];
                h.x = fma(c, hk.x, h.x);
                h.y = fma(c, hk.y, h.y);
            }
            const cplx m = M[e];
            out[i * d2 + j] = h;                                   // upper-left
            out[(i + d) * d2 + (j + d)] = h;                       // lower-right
            out[i * d2 + (j + d)] = cmake(-m.y * idt, m.x * idt);  // M / (-i dt) = i M / dt
            out[(i + d) * d2 + j] = cmake(0.0, 0.0);
            Y[e] = dU[e];
        }
        __syncwarp();
        if (n + 1 < N) {
            warp_mm(T, Y, F, d, lane);                 // F_{n+1} = dU_n F_n
            __syncwarp();
            cplx* t = F; F = T; T = t;
        }
    }
}

// grad[b,k,n] = (1/alpha_b) Re sum_ij W[i,j] G_k[j,i],  W = upper-right block of dUaug[b,n],  G_k = -i dt h_k.
// One warp per (b, n).
__global__ void grad_contract_kernel(const cplx* __restrict__ dUaug, const cplx* __restrict__ hks,
                                     const double* __restrict__ alpha, double* __restrict__ grad, const double dt,
                                     const int B, const int K, const int N, const int d) {
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)B * N) return;
    const int b = (int)(w / N), n = (int)(w - (long long)b * N);
    const int d2 = 2 * d, dd = d * d;
    const cplx* W = dUaug + (size_t)w * d2 * d2 + d;   // upper-right block, row stride d2
    const double a = alpha[b];
    const double inv_a = a > 0.0 ? 1.0 / a : 0.0;
    for (int k = 0; k < K; ++k) {
        double acc = 0.0;
        for (int e = lane; e < dd; e += 32) {
            const int i = e / d, j = e - i * d;
            const cplx wv = W[i * d2 + j];
            const cplx h = hks[(size_t)k * dd + j * d + i];
            // Re( w * (-i dt h) ) = dt * Re( w * (h.y - i h.x) ) = dt (w.x h.y + w.y h.x)
            acc = fma(wv.x, h.y, fma(wv.y, h.x, acc));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) grad[((size_t)b * K + k) * N + n] = acc * dt * inv_a;
    }
}

// =====================================================================================================================
// Second generation of the gradient path (shipped default for d <= 16): no augmented 2d x 2d exponential.
// The Frechet derivative of the degree-18 Taylor scheme is evaluated directly on the (X, dX) pairs of the
// block-triangular product  [[X, dX], [0, X]] [[Y, dY], [0, Y]] = [[XY, X dY + dX Y], [0, XY]]:  14 d x d products per
// slice instead of 5 products of 2d x 2d (= 40 d x d products, padded to 8-wide DMMA tiles), the contraction with the
// K control generators is fused (K doubles leave the SM per slice instead of a 2d x 2d matrix), and the generators are
// the trace-shifted ones of the forward pass:  L(A, M) = e^mu L(A - mu I, M).
//   A2 = A A              dA2 = A M + M A
//   A3 = A2 A             dA3 = dA2 A + A2 M
//   A6 = A3 A3            dA6 = dA3 A3 + A3 dA3
//   B_i = lin(A, A2, A3, A6)        dB_i = lin(M, dA2, dA3, dA6)
//   A9 = B4 + B1 B5       dA9 = dB4 + dB1 B5 + B1 dB5
//   T18 = B2 + (B3 + A9) A9         dT18 = dB2 + (dB3 + dA9) A9 + (B3 + A9) dA9
// =====================================================================================================================

// C (+)= X * Y for d x d matrices in shared memory; a lane owns chunks of TC consecutive entries of one row, so every
// loaded element of X feeds TC complex MACs (4.5x fewer shared loads than one entry per lane at d = 9, TC = 3).
template <int TC>
__device__ __forceinline__ void warp_mm_chunk(cplx* __restrict__ C, const cplx* __restrict__ X, const cplx* __restrict__ Y,
                                              const int d, const int lane, const bool accumulate) {
    const int cpr = d / TC;
    for (int q = lane; q < d * cpr; q += 32) {
        const int i = q / cpr, j0 = (q - i * cpr) * TC;
        cplx c[TC];
#pragma unroll
        for (int t = 0; t < TC; ++t) c[t] = accumulate ? C[i * d + j0 + t] : cmake(0.0, 0.0);
        for (int k = 0; k < d; ++k) {
            const cplx x = X[i * d + k];
#pragma unroll
            for (int t = 0; t < TC; ++t) cfma(c[t], x, Y[k * d + j0 + t]);
        }
#pragma unroll
        for (int t = 0; t < TC; ++t) C[i * d + j0 + t] = c[t];
    }
}

__device__ __forceinline__ void warp_mm2(cplx* C, const cplx* X, const cplx* Y, const int d, const int lane, const bool accumulate = false) {
    if (d % 3 == 0) warp_mm_chunk<3>(C, X, Y, d, lane, accumulate);
    else if (d % 2 == 0) warp_mm_chunk<2>(C, X, Y, d, lane, accumulate);
    else warp_mm_chunk<1>(C, X, Y, d, lane, accumulate);
}

// Backward sweep, as grad_suffix_kernel but with the chunked product.
__global__ void grad_suffix2_kernel(const cplx* __restrict__ dUs, const cplx* __restrict__ Ubar, cplx* __restrict__ Psi,
                                    double* __restrict__ alpha, const int B, const int N, const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int b = blockIdx.x * wpb + warp;
    if (b >= B) return;
    const int dd = d * d;
    cplx* P = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 3 * dd;
    cplx* Y = P + dd;
    cplx* T = P + 2 * dd;
    double nrm = 0.0;
    for (int e = lane; e < dd; e += 32) { const cplx u = Ubar[(size_t)b * dd + e]; nrm = fma(u.x, u.x, fma(u.y, u.y, nrm)); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    const double a = nrm > 0.0 ? rsqrt(nrm) : 0.0;
    if (lane == 0) alpha[b] = a;
    for (int e = lane; e < dd; e += 32) {   // P = alpha * Ubar^dag
        const int i = e / d, j = e - i * d;
        const cplx u = Ubar[(size_t)b * dd + j * d + i];
        P[e] = cmake(a * u.x, -a * u.y);
    }
    __syncwarp();
    for (int n = N - 1; n >= 0; --n) {
        cplx* out = Psi + ((size_t)b * N + n) * dd;
        const cplx* dU = dUs + ((size_t)b * N + n) * dd;
        for (int e = lane; e < dd; e += 32) { out[e] = P[e]; Y[e] = dU[e]; }
        __syncwarp();
        if (n > 0) {
            warp_mm2(T, P, Y, d, lane);
            __syncwarp();
            cplx* t = P; P = T; T = t;
        }
    }
}

// Forward sweep: F_0 = I, M_n = F_n Psi_n (written over Psi_n), F_{n+1} = dU_n F_n.
__global__ void grad_prefix2_kernel(const cplx* __restrict__ dUs, cplx* __restrict__ PsiM, const int B, const int N, const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int b = blockIdx.x * wpb + warp;
    if (b >= B) return;
    const int dd = d * d;
    cplx* F = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 4 * dd;
    cplx* Y = F + dd;
    cplx* T = F + 2 * dd;
    cplx* M = F + 3 * dd;
    for (int e = lane; e < dd; e += 32) F[e] = cmake((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
    __syncwarp();
    for (int n = 0; n < N; ++n) {
        cplx* psi = PsiM + ((size_t)b * N + n) * dd;
        const cplx* dU = dUs + ((size_t)b * N + n) * dd;
        for (int e = lane; e < dd; e += 32) Y[e] = psi[e];
        __syncwarp();
        warp_mm2(M, F, Y, d, lane);                    // M_n = F_n Psi_n
        __syncwarp();
        for (int e = lane; e < dd; e += 32) { psi[e] = M[e]; Y[e] = dU[e]; }
        __syncwarp();
        if (n + 1 < N) {
            warp_mm2(T, Y, F, d, lane);                // F_{n+1} = dU_n F_n
            __syncwarp();
            cplx* t = F; F = T; T = t;
        }
    }
}

constexpr int kFrechetBufs = 16;

// One warp per (b, n): W_n = L(A_n, M_n) by the Taylor-scheme derivative above, then
//   grad[b,k,n] = (1/alpha_b) Re tr(W_n G_k),  G_k = dA/dc_k  (-i dt h_k for the closed system, the commutator
//   superoperator -i dt (h_k (x) I - I (x) h_k^T) for Lindblad, where d is the superoperator dimension).
// G / RS / TR are the trace-shifted generators, their row sums and the shifts of the forward pass (RowsParams
// conventions; the unshifted G_k is G[k+1] + TR[k+1] I); Mn [B,N,d,d] from grad_prefix2_kernel.
// Dynamic smem = warps * 16 * d*d * 16 bytes.
__global__ void grad_frechet_kernel(const cplx* __restrict__ G, const double* __restrict__ RS, const cplx* __restrict__ TR,
                                    const double* __restrict__ signals, const cplx* __restrict__ Mn,
                                    const double* __restrict__ alpha, double* __restrict__ grad, const int B, const int K,
                                    const int N, const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int dd = d * d;
    cplx* buf = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * kFrechetBufs * dd;
    cplx* const bA = buf;            cplx* const bM = buf + dd;
    cplx* const bA2 = buf + 2 * dd;  cplx* const bdA2 = buf + 3 * dd;
    cplx* const bA3 = buf + 4 * dd;  cplx* const bdA3 = buf + 5 * dd;
    cplx* const bA6 = buf + 6 * dd;  cplx* const bdA6 = buf + 7 * dd;
    cplx* const bB1 = buf + 8 * dd;  cplx* const bB5 = buf + 9 * dd;
    cplx* const bdB1 = buf + 10 * dd; cplx* const bdB5 = buf + 11 * dd;
    cplx* const bA9 = buf + 12 * dd; cplx* const bdA9 = buf + 13 * dd;
    cplx* const bS = buf + 14 * dd;  cplx* const bdS = buf + 15 * dd;
    const long long total = (long long)B * N;
    for (long long w = (long long)blockIdx.x * wpb + warp; w < total; w += (long long)gridDim.x * wpb) {
        const int b = (int)(w / N), n = (int)(w - (long long)b * N);
        const double* sig_b = signals + (size_t)b * K * N;
        // ---- assemble the (trace-shifted) slice generator, its norm bound, mu_n, and load the direction ---------
        double nb = 0.0;
        for (int r = lane; r < d; r += 32) {
            double v = RS[r];
            for (int k = 0; k < K; ++k) v = fma(fabs(__ldg(sig_b + (size_t)k * N + n)), RS[(k + 1) * d + r], v);
            nb = fmax(nb, v);
        }
        nb = warp_max(nb);
        const int s = squarings_for(nb, C3B_THETA18);
        const double sc = pow2neg(s);
        cplx mu = TR ? TR[0] : cmake(0.0, 0.0);
        for (int k = 0; k < K && TR; ++k) {
            const double c = __ldg(sig_b + (size_t)k * N + n);
            mu.x = fma(c, TR[k + 1].x, mu.x);
            mu.y = fma(c, TR[k + 1].y, mu.y);
        }
        const cplx* Mg = Mn + (size_t)w * dd;
        for (int e = lane; e < dd; e += 32) {
            cplx v = G[e];
            for (int k = 0; k < K; ++k) {
                const double c = __ldg(sig_b + (size_t)k * N + n);
                const cplx gk = G[(size_t)(k + 1) * dd + e];
                v.x = fma(c, gk.x, v.x);
                v.y = fma(c, gk.y, v.y);
            }
            bA[e] = cmake(v.x * sc, v.y * sc);
            const cplx m = Mg[e];
            bM[e] = cmake(m.x * sc, m.y * sc);
        }
        __syncwarp();
        // ---- powers and their derivatives ----------------------------------------------------------------------
        warp_mm2(bA2, bA, bA, d, lane);
        warp_mm2(bdA2, bA, bM, d, lane);
        __syncwarp();
        warp_mm2(bdA2, bM, bA, d, lane, true);
        warp_mm2(bA3, bA2, bA, d, lane);
        __syncwarp();
        warp_mm2(bdA3, bdA2, bA, d, lane);
        __syncwarp();
        warp_mm2(bdA3, bA2, bM, d, lane, true);
        warp_mm2(bA6, bA3, bA3, d, lane);
        __syncwarp();
        warp_mm2(bdA6, bdA3, bA3, d, lane);
        __syncwarp();
        warp_mm2(bdA6, bA3, bdA3, d, lane, true);
        __syncwarp();
        // ---- combinations: B1, B5, dB1, dB5 to their own buffers; B4 -> bA, B3 -> bA2, B2 -> bA3,
        //      dB4 -> bM, dB3 -> bdA2, dB2 -> bdA3 (in place, entry by entry) -------------------------------------
        for (int e = lane; e < dd; e += 32) {
            const double dg = (e / d) == (e % d) ? 1.0 : 0.0;
            const cplx x1 = bA[e], x2 = bA2[e], x3 = bA3[e], x6 = bA6[e];
            const cplx y1 = bM[e], y2 = bdA2[e], y3 = bdA3[e], y6 = bdA6[e];
            bB1[e] = cmake(C3B_T18_A11 * x1.x + C3B_T18_A21 * x2.x + C3B_T18_A31 * x3.x,
                           C3B_T18_A11 * x1.y + C3B_T18_A21 * x2.y + C3B_T18_A31 * x3.y);
            bdB1[e] = cmake(C3B_T18_A11 * y1.x + C3B_T18_A21 * y2.x + C3B_T18_A31 * y3.x,
                            C3B_T18_A11 * y1.y + C3B_T18_A21 * y2.y + C3B_T18_A31 * y3.y);
            bB5[e] = cmake(C3B_T18_B24 * x2.x + C3B_T18_B34 * x3.x + C3B_T18_B64 * x6.x,
                           C3B_T18_B24 * x2.y + C3B_T18_B34 * x3.y + C3B_T18_B64 * x6.y);
            bdB5[e] = cmake(C3B_T18_B24 * y2.x + C3B_T18_B34 * y3.x + C3B_T18_B64 * y6.x,
                            C3B_T18_B24 * y2.y + C3B_T18_B34 * y3.y + C3B_T18_B64 * y6.y);
            bA[e] = cmake(C3B_T18_B03 * dg + C3B_T18_B13 * x1.x + C3B_T18_B23 * x2.x + C3B_T18_B33 * x3.x + C3B_T18_B63 * x6.x,
                          C3B_T18_B13 * x1.y + C3B_T18_B23 * x2.y + C3B_T18_B33 * x3.y + C3B_T18_B63 * x6.y);
            bM[e] = cmake(C3B_T18_B13 * y1.x + C3B_T18_B23 * y2.x + C3B_T18_B33 * y3.x + C3B_T18_B63 * y6.x,
                          C3B_T18_B13 * y1.y + C3B_T18_B23 * y2.y + C3B_T18_B33 * y3.y + C3B_T18_B63 * y6.y);
            bA2[e] = cmake(C3B_T18_B02 * dg + C3B_T18_B12 * x1.x + C3B_T18_B22 * x2.x + C3B_T18_B32 * x3.x + C3B_T18_B62 * x6.x,
                           C3B_T18_B12 * x1.y + C3B_T18_B22 * x2.y + C3B_T18_B32 * x3.y + C3B_T18_B62 * x6.y);
            bdA2[e] = cmake(C3B_T18_B12 * y1.x + C3B_T18_B22 * y2.x + C3B_T18_B32 * y3.x + C3B_T18_B62 * y6.x,
                            C3B_T18_B12 * y1.y + C3B_T18_B22 * y2.y + C3B_T18_B32 * y3.y + C3B_T18_B62 * y6.y);
            bA3[e] = cmake(C3B_T18_B11 * x1.x + C3B_T18_B21 * x2.x + C3B_T18_B31 * x3.x + C3B_T18_B61 * x6.x,
                           C3B_T18_B11 * x1.y + C3B_T18_B21 * x2.y + C3B_T18_B31 * x3.y + C3B_T18_B61 * x6.y);
            bdA3[e] = cmake(C3B_T18_B11 * y1.x + C3B_T18_B21 * y2.x + C3B_T18_B31 * y3.x + C3B_T18_B61 * y6.x,
                            C3B_T18_B11 * y1.y + C3B_T18_B21 * y2.y + C3B_T18_B31 * y3.y + C3B_T18_B61 * y6.y);
        }
        __syncwarp();
        // ---- A9 = B4 + B1 B5 (bA9),  dA9 = dB4 + dB1 B5 + B1 dB5 (bdA9) ---------------------------------------------
        for (int e = lane; e < dd; e += 32) { bA9[e] = bA[e]; bdA9[e] = bM[e]; }
        __syncwarp();
        warp_mm2(bA9, bB1, bB5, d, lane, true);
        warp_mm2(bdA9, bdB1, bB5, d, lane, true);
        __syncwarp();
        warp_mm2(bdA9, bB1, bdB5, d, lane, true);
        __syncwarp();
        // ---- S = B3 + A9, dS = dB3 + dA9;  X = B2 + S A9 (bA6),  dX = dB2 + dS A9 + S dA9 (bdA6) -------------------
        for (int e = lane; e < dd; e += 32) {
            bS[e] = cmake(bA2[e].x + bA9[e].x, bA2[e].y + bA9[e].y);
            bdS[e] = cmake(bdA2[e].x + bdA9[e].x, bdA2[e].y + bdA9[e].y);
            bA6[e] = bA3[e];
            bdA6[e] = bdA3[e];
        }
        __syncwarp();
        if (s > 0) warp_mm2(bA6, bS, bA9, d, lane, true);      // T18 itself is only needed to undo the scaling
        warp_mm2(bdA6, bdS, bA9, d, lane, true);
        __syncwarp();
        warp_mm2(bdA6, bS, bdA9, d, lane, true);
        __syncwarp();
        cplx* X = bA6;
        cplx* dX = bdA6;
        cplx* X2 = bB1;
        cplx* dX2 = bdB1;
        for (int q = 0; q < s; ++q) {                           // (X, dX) <- (X X, dX X + X dX)
            warp_mm2(dX2, dX, X, d, lane);
            if (q + 1 < s) warp_mm2(X2, X, X, d, lane);
            __syncwarp();
            warp_mm2(dX2, X, dX, d, lane, true);
            __syncwarp();
            cplx* t = X; X = X2; X2 = t;
            t = dX; dX = dX2; dX2 = t;
        }
        // ---- contraction with the control generators ---------------------------------------------------------------
        const cplx ph = cexp_(mu);
        const double a = alpha[b];
        const double inv_a = a > 0.0 ? 1.0 / a : 0.0;
        for (int k = 0; k < K; ++k) {
            const cplx tk = TR ? TR[k + 1] : cmake(0.0, 0.0);
            double acc = 0.0;
            for (int e = lane; e < dd; e += 32) {
                const int i = e / d, j = e - i * d;
                const cplx wv = cmul(ph, dX[e]);
                cplx g = G[(size_t)(k + 1) * dd + j * d + i];
                if (i == j) { g.x += tk.x; g.y += tk.y; }
                acc = fma(wv.x, g.x, fma(-wv.y, g.y, acc));     // Re(w_ij g_ji)
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) grad[((size_t)b * K + k) * N + n] = acc * inv_a;
        }
        __syncwarp();
    }
}

}  // namespace c3b
