// Ordered matrix products on the device:
//   * tf_matmul_left / tf_matmul_n semantics  U = M_{L-1} ... M_1 M_0   (c3/utils/tf_utils.py:120-193)
//   * evaluate_sequences: gather gate matrices by index, same product, empty -> identity
//                                                                  (c3/libraries/propagation.py:588-627)
//   * batched Kronecker product for the superoperator helpers tf_kron/tf_spre/tf_spost
//                                                                  (c3/utils/tf_utils.py:257-280)
// The reference's pairwise tree and its sequential fold are the same product up to fp64
// re-association; here each CTA folds a contiguous segment sequentially and a second launch
// folds the segment results.
#pragma once
#include "pwc_cta.cuh"

namespace c3b {


template <int CT, int TR, int TC>
__global__ void __launch_bounds__(kCtaThreads) product_kernel(const ProductParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = p.D, DD = D * D, tid = threadIdx.x;
    cplx* P = p.use_smem ? reinterpret_cast<cplx*>(smem_raw) : p.ws + (size_t)blockIdx.x * 2 * DD;
    cplx* T = P + DD;
    const long long units = (long long)p.B * p.S;
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int b = (int)(unit / p.S);
        const int sidx = (int)(unit - (long long)b * p.S);
        const int len = p.lens ? min(p.lens[b], p.M) : p.M;
        const int m0 = sidx * p.seg_len;
        const int m1 = min(len, m0 + p.seg_len);
        cplx* o = p.out + (size_t)unit * DD;
        if (m1 <= m0) {
            for (int e = tid; e < DD; e += kCtaThreads) o[e] = cmake((e / D) == (e % D) ? 1.0 : 0.0, 0.0);
            continue;
        }
        auto mat = [&](int m) -> const cplx* {
            return p.idx ? p.mats + (size_t)p.idx[(size_t)b * p.M + m] * DD : p.mats + ((size_t)b * p.M + m) * DD;
        };
        {
            const cplx* first = mat(m0);
            for (int e = tid; e < DD; e += kCtaThreads) P[e] = first[e];
        }
        __syncthreads();
        for (int m = m0 + 1; m < m1; ++m) {
            cta_gemm<CT, TR, TC>(T, mat(m), P, D);
            __syncthreads();
            cplx* t = P; P = T; T = t;
        }
        for (int e = tid; e < DD; e += kCtaThreads) o[e] = P[e];
        __syncthreads();
    }
}

// out[b] = A[b] (x) B[b]; a_bstride / b_bstride may be 0 to broadcast one operand.
__global__ void kron_kernel(const cplx* __restrict__ A, const cplx* __restrict__ Bm, cplx* __restrict__ out,
                            int batch, int ra, int ca, int rb, int cb, long long a_bstride, long long b_bstride) {
    const long long per = (long long)ra * rb * ca * cb;
    const long long total = per * batch;
    const int ncol = ca * cb;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / per);
        const long long rem = e - b * per;
        const int I = (int)(rem / ncol), J = (int)(rem - (long long)I * ncol);
        const int i1 = I / rb, i2 = I - i1 * rb, j1 = J / cb, j2 = J - j1 * cb;
        out[e] = cmul(A[b * a_bstride + (long long)i1 * ca + j1], Bm[b * b_bstride + (long long)i2 * cb + j2]);
    }
}

// ---- generator set-up (once per launch; never on the per-slice path) ------------------------

// G_k = (-i dt) H_k for the closed system.  h0 [(Bm), d, d], hks [(Bm), K, d, d].
__global__ void setup_closed_kernel(const cplx* __restrict__ h0, const cplx* __restrict__ hks, cplx* __restrict__ G,
                                    int Bm, int K, int d, double dt) {
    const long long per = (long long)(K + 1) * d * d;
    const long long total = per * Bm;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int bm = (int)(e / per);
        const long long rem = e - bm * per;
        const int k = (int)(rem / (d * d));
        const int off = (int)(rem - (long long)k * d * d);
        const cplx h = (k == 0) ? h0[(long long)bm * d * d + off] : hks[((long long)bm * K + (k - 1)) * d * d + off];
        G[e] = cmake(h.y * dt, -h.x * dt);
    }
}

// Lindblad generators, D = d^2, row-major Kronecker convention of tf_kron:
//   G_0 = dt [ -i (h0 (x) I - I (x) h0^T) + sum_c ( L_c (x) L_c^* - 1/2 M (x) I - 1/2 I (x) M^T ) ],  M = sum_c L_c^dag L_c
//   G_k = -i dt ( h_k (x) I - I (x) h_k^T )
// (c3/libraries/propagation.py:563-582 builds the same operator from Kronecker temporaries.)
__global__ void setup_lindblad_kernel(const cplx* __restrict__ h0, const cplx* __restrict__ hks,
                                      const cplx* __restrict__ col, cplx* __restrict__ G, int Bm, int K, int C,
                                      int d, double dt) {
    const int D = d * d;
    const long long per = (long long)(K + 1) * D * D;
    const long long total = per * Bm;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int bm = (int)(e / per);
        const long long rem = e - bm * per;
        const int k = (int)(rem / ((long long)D * D));
        const int off = (int)(rem - (long long)k * D * D);
        const int I = off / D, J = off - I * D;
        const int i1 = I / d, i2 = I - i1 * d, j1 = J / d, j2 = J - j1 * d;
        const cplx* h = (k == 0) ? h0 + (long long)bm * d * d : hks + ((long long)bm * K + (k - 1)) * d * d;
        cplx comm = cmake(0.0, 0.0);  // (h (x) I - I (x) h^T)[I,J]
        if (i2 == j2) { const cplx v = h[i1 * d + j1]; comm.x += v.x; comm.y += v.y; }
        if (i1 == j1) { const cplx v = h[j2 * d + i2]; comm.x -= v.x; comm.y -= v.y; }
        cplx g = cmake(comm.y * dt, -comm.x * dt);  // -i dt comm
        if (k == 0 && col != nullptr) {
            const cplx* cb = col + (long long)bm * C * d * d;
            cplx diss = cmake(0.0, 0.0);
            for (int c = 0; c < C; ++c) {
                const cplx* Lc = cb + (long long)c * d * d;
                const cplx a = Lc[i1 * d + j1];
                const cplx bconj = cmake(Lc[i2 * d + j2].x, -Lc[i2 * d + j2].y);
                cfma(diss, a, bconj);
                if (i2 == j2) {  // -1/2 M[i1,j1]
                    cplx m = cmake(0.0, 0.0);
                    for (int x = 0; x < d; ++x) cfma(m, cmake(Lc[x * d + i1].x, -Lc[x * d + i1].y), Lc[x * d + j1]);
                    diss.x -= 0.5 * m.x; diss.y -= 0.5 * m.y;
                }
                if (i1 == j1) {  // -1/2 M[j2,i2]
                    cplx m = cmake(0.0, 0.0);
                    for (int x = 0; x < d; ++x) cfma(m, cmake(Lc[x * d + j2].x, -Lc[x * d + j2].y), Lc[x * d + i2]);
                    diss.x -= 0.5 * m.x; diss.y -= 0.5 * m.y;
                }
            }
            g.x = fma(dt, diss.x, g.x);
            g.y = fma(dt, diss.y, g.y);
        }
        G[e] = g;
    }
}

// Trace shift (Higham's preprocessing): t_m = tr(G_m)/D is subtracted from the diagonal of G_m and
// stored in TR[m]; the kernels re-apply exp(t_0 + sum_k c_k t_k).  One warp per matrix.
__global__ void trace_shift_kernel(cplx* __restrict__ G, cplx* __restrict__ TR, long long nmat, int D) {
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= nmat) return;
    cplx* g = G + m * (long long)D * D;
    double tr = 0.0, ti = 0.0;
    for (int j = lane; j < D; j += 32) { tr += g[(long long)j * D + j].x; ti += g[(long long)j * D + j].y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tr += __shfl_xor_sync(0xffffffffu, tr, o);
        ti += __shfl_xor_sync(0xffffffffu, ti, o);
    }
    tr /= D; ti /= D;
    for (int j = lane; j < D; j += 32) { g[(long long)j * D + j].x -= tr; g[(long long)j * D + j].y -= ti; }
    if (lane == 0) TR[m] = cmake(tr, ti);
}

// RS[m, r] = sum_j |G[m, r, j]|  for m over (Bm * (K+1)) matrices; one warp per row.
__global__ void rowsum_kernel(const cplx* __restrict__ G, double* __restrict__ RS, long long nrows, int D) {
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    double s = 0.0;
    for (int j = lane; j < D; j += 32) s += cabs1(G[row * D + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) RS[row] = s;
}

}  // namespace c3b
