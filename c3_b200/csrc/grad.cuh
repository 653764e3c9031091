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

}  // namespace c3b
