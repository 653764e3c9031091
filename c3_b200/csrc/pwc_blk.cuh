// Block-layout PWC propagator kernel for small Hilbert dimensions (D = NB * BS).
//
// Fused contract (assemble -> expm -> ordered product, replacing c3/libraries/propagation.py:426-440,460-515 and
// c3/utils/tf_utils.py:120-193): a group of NB x NB lanes owns one matrix, lane (bi, bj) holds the BS x BS block
// (bi, bj) of every register-resident matrix (4*BS*BS registers each).  A product streams BOTH operands from shared
// memory: per k, BS elements of X's block-row and BS of Y's block-column feed BS*BS complex MACs, i.e. BS/2 cfma per
// 16-byte operand instead of 1 in a one-row-per-lane layout -- the shared-memory wavefront pipe (128 lane-bytes/clk/SM,
// every LDS.128 = 4 wavefronts) is what bounds these kernels (profiles/README_r01.md).  The exponential is a Taylor
// polynomial evaluated in four products (c3b_common.cuh): no division, no shuffles, no pivot chain.
#pragma once
#include "c3b_common.cuh"
#include "c3b_params.cuh"

namespace c3b {

template <int D, int BS>
struct BlkLayout {
    static constexpr int NB = D / BS;
    static constexpr int LPM = NB * NB;             // lanes per matrix
    static constexpr int MPW = 32 / LPM;            // matrices per warp
    // Leading dimension of the per-matrix shared buffers.  With LD = 3 (mod 8) [D=9: 11] the NB*NB
    // lanes of a group hit distinct 16-byte bank slots for block stores, block-row and block-column
    // loads; with the group stride = 4 (mod 8) neighbouring groups in a quarter-warp do not collide
    // either (ncu before: 5.7 wavefronts per LDS.128 and 9 per STS.128 instead of 4).
    static constexpr int LD = (D % 8 == 1) ? D + 2 : ((D % 8 == 3) ? D : ((D % 2 == 0) ? D + 1 : D));
    static constexpr int BUF = D * LD;
    static constexpr int GROUP_PAD = (4 - (4 * BUF) % 8 + 8) % 8;
    static constexpr int GROUP_ELEMS = 4 * BUF + GROUP_PAD; // bufA, bufA2, bufX, bufP
    // D = 9 (3 groups of 9 lanes): group base offsets {0,5,3} (mod 8) and a rotation of the lane ->
    // block assignment inside group 1 make every quarter-warp hit 8 distinct 16-byte bank slots for
    // block stores AND both operand loads (brute-force checked: 4 wavefronts per LDS/STS.128, the minimum).
    static constexpr bool TUNED9 = (D == 9 && BS == 3);
    static constexpr int WARP_ELEMS = TUNED9 ? 1208 : MPW * GROUP_ELEMS;
    __host__ __device__ static constexpr int group_off(int g) { return TUNED9 ? (g == 0 ? 0 : (g == 1 ? 405 : 803)) : g * GROUP_ELEMS; }
    __host__ __device__ static constexpr int rot(int g) { return (TUNED9 && g == 1) ? 2 : 0; }
    static_assert(D % BS == 0, "D must be a multiple of the block size");
    __host__ __device__ static size_t smem_bytes(int K, int warps) {
        size_t model = (size_t)(K + 1) * D * LD * sizeof(cplx) + (size_t)(((K + 1) * D + 1) & ~1) * sizeof(double);
        return model + (size_t)warps * WARP_ELEMS * sizeof(cplx);
    }
};

// C(block) = X(block-row) * Y(block-column); Xr = &X[r0*D], Yc = &Y[c0]
template <int D, int BS, int LD>
__device__ __forceinline__ void mm_blk(const cplx* __restrict__ Xr, const cplx* __restrict__ Yc, cplx (&c)[BS][BS]) {
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) c[a][b] = cmake(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        cplx x[BS], y[BS];
#pragma unroll
        for (int a = 0; a < BS; ++a) x[a] = Xr[a * LD + k];
#pragma unroll
        for (int b = 0; b < BS; ++b) y[b] = Yc[k * LD + b];
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b) cfma(c[a][b], x[a], y[b]);
    }
}

template <int D, int BS, int LD>
__device__ __forceinline__ void store_blk(cplx* __restrict__ Mrc, const cplx (&x)[BS][BS], bool pred) {
    if (pred) {
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b) Mrc[a * LD + b] = x[a][b];
    }
}

__device__ __forceinline__ cplx shfl_c(const cplx v, const int src) {
    cplx r;
    r.x = __shfl_sync(0xffffffffu, v.x, src);
    r.y = __shfl_sync(0xffffffffu, v.y, src);
    return r;
}

// =============================================================================================
// Same mapping, exponential by the degree-15+ Taylor scheme in 4 products (c3b_common.cuh) instead
// of Pade + Gauss-Jordan: no division, no shuffles, no serial pivot chain -- every phase is a dense
// block product or an element-wise combination, which is what a kernel with 2 warps per
// scheduler needs (ncu: the Gauss-Jordan sweep took 33 % of the Pade kernel's time at 1/6 of its
// flops).  The generators arrive TRACE-SHIFTED (G_k - tr(G_k)/d I, Higham's preprocessing step):
// for lab-frame Hamiltonians with a large diagonal this halves ||A|| (headline: 1.33 -> 0.75), so
// the slice needs no squaring; exp(mu_n) factors are accumulated as one complex sum per lane group
// and applied once per segment (per slice only when the partial propagators are stored).
// =============================================================================================
template <int D, int BS, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) pwc_blk_taylor_kernel(const RowsParams p, unsigned int* __restrict__ counter) {
    using L = BlkLayout<D, BS>;
    constexpr int NB = L::NB, LPM = L::LPM, MPW = L::MPW, LD = L::LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);                 // [(K+1), D, LD] zero padded
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * D * LD);
    cplx* sWarps = reinterpret_cast<cplx*>(sRS + (((K + 1) * D + 1) & ~1));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool hmode = p.hlist != nullptr;

    if (!hmode) {
        for (int idx = tid; idx < (K + 1) * D * D; idx += WARPS * 32) {
            const int k = idx / (D * D);
            const int rem = idx - k * D * D;
            const int r = rem / D, j = rem - r * D;
            cplx v = cmake(0.0, 0.0);
            if (r < d && j < d) v = p.G[(size_t)k * d * d + r * d + j];
            sG[(k * D + r) * LD + j] = v;
        }
        for (int idx = tid; idx < (K + 1) * D; idx += WARPS * 32) {
            const int k = idx / D, r = idx - k * D;
            sRS[idx] = (r < d) ? p.RS[k * d + r] : 0.0;
        }
    }
    __syncthreads();

    const int g_raw = lane / LPM;
    const bool lane_on = g_raw < MPW;
    const int g = lane_on ? g_raw : MPW - 1;
    const int li = ((lane_on ? (lane - g_raw * LPM) : LPM - 1) + L::rot(g)) % LPM;   // block index owned by this lane
    const int bi = li / NB, bj = li - bi * NB;
    const int r0 = bi * BS, c0 = bj * BS;
    const bool on_diag = (bi == bj);

    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + L::group_off(g);
    cplx* bufA = gbase;
    cplx* bufB = gbase + L::BUF;      // A^2, then B1, then the left operands of the later products
    cplx* bufX = gbase + 2 * L::BUF;  // A^3, then B5, then A9
    cplx* bufP = gbase + 3 * L::BUF;
    const int rc_off = r0 * LD + c0;
    const int mg_off = r0 * LD + c0;
    const cplx hs = cmake(p.hscale_re, p.hscale_im);
    const long long total_units = (long long)p.B * p.S;
    const bool shifted = (p.TR != nullptr) && !hmode;

    for (;;) {
        unsigned int unit_u = 0;
        if (lane == 0) unit_u = atomicAdd(counter, 1u);
        unit_u = __shfl_sync(0xffffffffu, unit_u, 0);
        const long long unit = unit_u;
        if (unit >= total_units) break;
        const int b = (int)(unit / p.S);
        const int sidx = (int)(unit - (long long)b * p.S);
        const int n_begin = sidx * p.seg_len;
        const int n_end = min(p.N, n_begin + p.seg_len);
        const int len = n_end - n_begin;
        const int cl = (len + MPW - 1) / MPW;
        const int my_begin = n_begin + g * cl;
        const int my_end = min(n_end, my_begin + cl);
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        cplx mu_acc = cmake(0.0, 0.0);    // sum of the trace shifts of this group's slices

#pragma unroll 1
        for (int it = 0; it < cl; ++it) {
            const int n = my_begin + it;
            const bool on = lane_on && (n < my_end);

            cplx R1[BS][BS];   // A, later B2
            cplx mu = cmake(0.0, 0.0);
            double nb = 0.0;
            if (!hmode) {
                double nba[BS];
#pragma unroll
                for (int a = 0; a < BS; ++a) {
                    nba[a] = on ? sRS[r0 + a] : 0.0;
#pragma unroll
                    for (int c = 0; c < BS; ++c) R1[a][c] = on ? sG[mg_off + a * LD + c] : cmake(0.0, 0.0);
                }
                if (shifted && on) mu = p.TR[0];
                for (int k = 0; k < K; ++k) {
                    const double cs = on ? __ldg(sig_b + (size_t)k * p.N + n) : 0.0;
                    const cplx* gk = sG + (k + 1) * D * LD + mg_off;
#pragma unroll
                    for (int a = 0; a < BS; ++a) {
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            const cplx gv = gk[a * LD + c];
                            R1[a][c].x = fma(cs, gv.x, R1[a][c].x);
                            R1[a][c].y = fma(cs, gv.y, R1[a][c].y);
                        }
                        nba[a] = fma(fabs(cs), sRS[(k + 1) * D + r0 + a], nba[a]);
                    }
                    if (shifted) { const cplx t = p.TR[k + 1]; mu.x = fma(cs, t.x, mu.x); mu.y = fma(cs, t.y, mu.y); }
                }
#pragma unroll
                for (int a = 0; a < BS; ++a) nb = fmax(nb, nba[a]);
            } else {
#pragma unroll
                for (int a = 0; a < BS; ++a) {
                    const int row = r0 + a;
                    double rs = 0.0;
#pragma unroll
                    for (int c = 0; c < BS; ++c) {
                        cplx h = cmake(0.0, 0.0);
                        if (on && row < d && c0 + c < d)
                            h = p.hlist[((size_t)b * p.N + n) * d * d + (size_t)row * d + c0 + c];
                        R1[a][c] = cmul(hs, h);
                        rs += cabs1(R1[a][c]);
                    }
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < NB; ++q) tot += __shfl_sync(0xffffffffu, rs, g * LPM + (bi * NB + q - L::rot(g) + LPM) % LPM);
                    nb = fmax(nb, tot);
                }
            }
            nb = warp_max(nb);
            mu_acc.x += mu.x; mu_acc.y += mu.y;

            const int s = squarings_for(nb, C3B_THETA15);
            if (s > 0) {
                const double sc = pow2neg(s);
#pragma unroll
                for (int a = 0; a < BS; ++a)
#pragma unroll
                    for (int c = 0; c < BS; ++c) { R1[a][c].x *= sc; R1[a][c].y *= sc; }
            }
            store_blk<D, BS, LD>(bufA + rc_off, R1, lane_on);
            __syncwarp();

            cplx R2[BS][BS], R3[BS][BS], R4[BS][BS], C[BS][BS];
            // phases (degree-15+ scheme in four products, c3b_common.cuh):
            //   0 A2 (-> Q0) | 1 P0 = A2 Q0 (-> L1, R1) | 2 L1 R1 (-> P1, L2, R2, E0) | 3 T = L2 R2 + E0 | s squarings | product
            const int ph_lastsq = 3 + s;
            const int ph_last = ph_lastsq + (it > 0 ? 1 : 0);
            const cplx* Xr = bufA + r0 * LD;
            const cplx* Yc = bufA + c0;

#pragma unroll 1
            for (int ph = 0; ph <= ph_last; ++ph) {
                mm_blk<D, BS, LD>(Xr, Yc, C);
                if (ph == 0) {                                  // C = A^2;  Q0 = a1 A2 + a2 A
#pragma unroll
                    for (int a = 0; a < BS; ++a)
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            R2[a][c] = C[a][c];
                            const cplx q = cmake(C3B_T15_A1 * C[a][c].x + C3B_T15_A2 * R1[a][c].x, C3B_T15_A1 * C[a][c].y + C3B_T15_A2 * R1[a][c].y);
                            if (lane_on) {
                                bufB[rc_off + a * LD + c] = C[a][c];    // left operand A2
                                bufX[rc_off + a * LD + c] = q;          // right operand Q0
                            }
                        }
                    __syncwarp();
                    Xr = bufB + r0 * LD; Yc = bufX + c0;
                } else if (ph == 1) {                           // C = P0 = A2 Q0;  L1 = P0 + b1 A2 + b2 A,  R1 = P0 + b3 A2 + b4 I
                    __syncwarp();                               // bufB (A2) and bufX (Q0) fully read
#pragma unroll
                    for (int a = 0; a < BS; ++a)
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            const cplx x1 = R1[a][c], x2 = R2[a][c], q = C[a][c];
                            const double dg = (on_diag && a == c) ? 1.0 : 0.0;
                            R3[a][c] = q;
                            if (lane_on) {
                                bufB[rc_off + a * LD + c] = cmake(q.x + C3B_T15_B1 * x2.x + C3B_T15_B2 * x1.x, q.y + C3B_T15_B1 * x2.y + C3B_T15_B2 * x1.y);
                                bufX[rc_off + a * LD + c] = cmake(q.x + C3B_T15_B3 * x2.x + C3B_T15_B4 * dg, q.y + C3B_T15_B3 * x2.y);
                            }
                        }
                    __syncwarp();
                    Xr = bufB + r0 * LD; Yc = bufX + c0;
                } else if (ph == 2) {                           // C = L1 R1;  P1 = C + b5 P0 -> L2, R2 (published), E0 (kept)
                    __syncwarp();
#pragma unroll
                    for (int a = 0; a < BS; ++a)
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            const cplx x1 = R1[a][c], x2 = R2[a][c], q = R3[a][c];
                            const double dg = (on_diag && a == c) ? 1.0 : 0.0;
                            const cplx p1 = cmake(C[a][c].x + C3B_T15_B5 * q.x, C[a][c].y + C3B_T15_B5 * q.y);
                            R4[a][c] = cmake(C3B_T15_C9 * p1.x + C3B_T15_C5 * q.x + C3B_T15_C6 * x2.x + C3B_T15_C7 * x1.x + C3B_T15_C8 * dg,
                                             C3B_T15_C9 * p1.y + C3B_T15_C5 * q.y + C3B_T15_C6 * x2.y + C3B_T15_C7 * x1.y);
                            if (lane_on) {
                                bufB[rc_off + a * LD + c] = cmake(p1.x + C3B_T15_C1 * x2.x + C3B_T15_C2 * x1.x, p1.y + C3B_T15_C1 * x2.y + C3B_T15_C2 * x1.y);
                                bufX[rc_off + a * LD + c] = cmake(p1.x + C3B_T15_C3 * q.x + C3B_T15_C4 * x1.x, p1.y + C3B_T15_C3 * q.y + C3B_T15_C4 * x1.y);
                            }
                        }
                    __syncwarp();
                    Xr = bufB + r0 * LD; Yc = bufX + c0;
                } else {
                    if (ph == 3) {                              // C = L2 R2  ->  T = C + E0
#pragma unroll
                        for (int a = 0; a < BS; ++a)
#pragma unroll
                            for (int c = 0; c < BS; ++c) { C[a][c].x += R4[a][c].x; C[a][c].y += R4[a][c].y; }
                    }
                    if (ph <= ph_lastsq) {
                        // C = exp(A_n / 2^s)^(2^(ph-3)); publish as the next left operand
                        __syncwarp();
                        store_blk<D, BS, LD>(bufB + rc_off, C, lane_on);
                        __syncwarp();
                        Xr = bufB + r0 * LD;
                        Yc = (ph < ph_lastsq) ? (bufB + c0) : (bufP + c0);
                        if (ph == ph_lastsq) {
                            if (p.dUs_out != nullptr && on) {
                                const cplx ph_n = shifted ? cexp_(mu) : cmake(1.0, 0.0);
#pragma unroll
                                for (int a = 0; a < BS; ++a) {
                                    const int row = r0 + a;
                                    if (row < d) {
                                        cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)row * d + c0;
#pragma unroll
                                        for (int c = 0; c < BS; ++c)
                                            if (c0 + c < d) o[c] = shifted ? cmul(ph_n, C[a][c]) : C[a][c];
                                    }
                                }
                            }
                            if (it == 0) store_blk<D, BS, LD>(bufP + rc_off, C, lane_on);
                        }
                    } else {                                    // C = dU_n * P
                        __syncwarp();
                        store_blk<D, BS, LD>(bufP + rc_off, C, lane_on);
                    }
                }
            }
        }
        __syncwarp();

        // ---- re-apply this group's accumulated shift: P_g <- exp(sum mu) P_g -------------------------
        if (shifted) {
            const cplx ph_g = cexp_(mu_acc);
#pragma unroll
            for (int a = 0; a < BS; ++a)
#pragma unroll
                for (int c = 0; c < BS; ++c) {
                    const cplx v = bufP[rc_off + a * LD + c];
                    if (lane_on) bufP[rc_off + a * LD + c] = cmul(ph_g, v);
                }
            __syncwarp();
        }

        // ---- fold the group products P_{MPW-1} ... P_1 P_0 as a pairwise tree: at stride s group g (g % 2s == 0)
        // forms P_{g+s} P_g -- log2(MPW) dependent products instead of MPW - 1 (d <= 3: 5 instead of 31, which was most
        // of a latency-bound B = 1 call).  Levels ping-pong between the group's P / X / A buffers.
        cplx* wbase = sWarps + (size_t)warp * L::WARP_ELEMS;
        int src_slot = 3;
#pragma unroll 1
        for (int stride = 1; stride < MPW; stride <<= 1) {
            const int dst_slot = (src_slot == 2) ? 0 : 2;
            if (lane_on && (g % (2 * stride)) == 0) {
                const cplx* own = gbase + src_slot * L::BUF;
                cplx T[BS][BS];
                if (g + stride < MPW) {
                    mm_blk<D, BS, LD>(wbase + L::group_off(g + stride) + src_slot * L::BUF + r0 * LD, own + c0, T);
                } else {
#pragma unroll
                    for (int a = 0; a < BS; ++a)
#pragma unroll
                        for (int c = 0; c < BS; ++c) T[a][c] = own[rc_off + a * LD + c];
                }
                store_blk<D, BS, LD>(gbase + dst_slot * L::BUF + rc_off, T, true);
            }
            __syncwarp();
            src_slot = dst_slot;
        }
        const cplx* cur = wbase + L::group_off(0) + src_slot * L::BUF;
        if (lane_on && g == 0) {
            cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
            for (int a = 0; a < BS; ++a) {
                const int row = r0 + a;
                if (row < d) {
#pragma unroll
                    for (int c = 0; c < BS; ++c)
                        if (c0 + c < d) o[row * d + c0 + c] = cur[(r0 + a) * LD + c0 + c];
                }
            }
        }
        __syncwarp();
    }
}

// =============================================================================================
// evaluate_sequences for small dimensions (c3/libraries/propagation.py:588-627): U_seq = G[i_{L-1}] ... G[i_1] G[i_0],
// empty sequence -> identity.  The CTA-per-sequence product_kernel pays a __syncthreads round per factor (4096 sequences
// of ~45 gates at d = 9: 0.45 ms, all latency); here a lane group of NB x NB lanes owns one sequence (MPW sequences per
// warp), the gate table sits zero-padded in shared memory, the running product ping-pongs between two per-group buffers
// and only __syncwarp separates the factors.
// =============================================================================================
template <int D, int BS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) seq_product_blk_kernel(const cplx* __restrict__ gates, const int Gn,
                                                                     const int* __restrict__ idx, const int* __restrict__ lens,
                                                                     const int S, const int Lmax, const int d,
                                                                     cplx* __restrict__ out) {
    using L = BlkLayout<D, BS>;
    constexpr int NB = L::NB, LPM = L::LPM, MPW = L::MPW, LD = L::LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sGates = reinterpret_cast<cplx*>(smem_raw);                       // [Gn, D, LD] zero padded
    cplx* sWarps = sGates + (size_t)Gn * L::BUF;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < Gn * L::BUF; e += WARPS * 32) {
        const int gi = e / L::BUF, rem = e - gi * L::BUF;
        const int r = rem / LD, j = rem - r * LD;
        sGates[e] = (r < d && j < d) ? gates[(size_t)gi * d * d + r * d + j] : cmake(0.0, 0.0);
    }
    __syncthreads();
    const int g_raw = lane / LPM;
    const bool lane_on = g_raw < MPW;
    const int g = lane_on ? g_raw : MPW - 1;
    const int li = ((lane_on ? (lane - g_raw * LPM) : LPM - 1) + L::rot(g)) % LPM;
    const int bi = li / NB, bj = li - bi * NB;
    const int r0 = bi * BS, c0 = bj * BS;
    const int rc_off = r0 * LD + c0;
    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + L::group_off(g);
    const long long nwarp_units = ((long long)S + MPW - 1) / MPW;
    for (long long wu = (long long)blockIdx.x * WARPS + warp; wu < nwarp_units; wu += (long long)gridDim.x * WARPS) {
        const long long seq = wu * MPW + g;
        const bool have = lane_on && seq < S;
        const int len = have ? min(lens[seq], Lmax) : 0;
        int maxlen = len;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
        cplx* P = gbase;                 // running product
        cplx* T = gbase + L::BUF;
        cplx C[BS][BS];
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int c = 0; c < BS; ++c) C[a][c] = cmake((bi == bj && a == c && r0 + a < d) ? 1.0 : 0.0, 0.0);
        store_blk<D, BS, LD>(P + rc_off, C, lane_on);
        __syncwarp();
        const int* my_idx = idx + (size_t)(have ? seq : 0) * Lmax;
        for (int m = 0; m < maxlen; ++m) {
            const bool active = m < len;
            int gi = active ? __ldg(my_idx + m) : 0;
            gi = min(max(gi, 0), Gn - 1);
            mm_blk<D, BS, LD>(sGates + (size_t)gi * L::BUF + r0 * LD, P + c0, C);
            if (active) {                // group-uniform: the group's buffers swap roles
                store_blk<D, BS, LD>(T + rc_off, C, lane_on);
                cplx* t = P; P = T; T = t;
            }
            __syncwarp();
        }
        if (have) {
#pragma unroll
            for (int a = 0; a < BS; ++a) {
                const int row = r0 + a;
                if (row < d) {
#pragma unroll
                    for (int c = 0; c < BS; ++c)
                        if (c0 + c < d) out[(size_t)seq * d * d + row * d + c0 + c] = P[rc_off + a * LD + c];
                }
            }
        }
        __syncwarp();
    }
}

// Ordered product of M matrices per batch row for small dimensions, out[b] = mats[b,M-1] ... mats[b,0]
// (tf_matmul_left / tf_matmul_n, c3/utils/tf_utils.py:120-193; also the fold of the fused kernels' segment products):
// a lane group per batch row, the next factor's own block is fetched from global memory while the current product runs.
template <int D, int BS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) fold_blk_kernel(const cplx* __restrict__ mats, const int B, const int M, const int d,
                                                              cplx* __restrict__ out) {
    using L = BlkLayout<D, BS>;
    constexpr int NB = L::NB, LPM = L::LPM, MPW = L::MPW, LD = L::LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sWarps = reinterpret_cast<cplx*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g_raw = lane / LPM;
    const bool lane_on = g_raw < MPW;
    const int g = lane_on ? g_raw : MPW - 1;
    const int li = ((lane_on ? (lane - g_raw * LPM) : LPM - 1) + L::rot(g)) % LPM;
    const int bi = li / NB, bj = li - bi * NB;
    const int r0 = bi * BS, c0 = bj * BS;
    const int rc_off = r0 * LD + c0;
    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + L::group_off(g);
    cplx* bufX = gbase + 2 * L::BUF;
    const long long nwu = ((long long)B + MPW - 1) / MPW;
    auto fetch = [&](const cplx* src, cplx (&x)[BS][BS], const bool pred) {
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int c = 0; c < BS; ++c)
                x[a][c] = (pred && r0 + a < d && c0 + c < d) ? src[(r0 + a) * d + c0 + c] : cmake(0.0, 0.0);
    };
    for (long long wu = (long long)blockIdx.x * WARPS + warp; wu < nwu; wu += (long long)gridDim.x * WARPS) {
        const long long b = wu * MPW + g;
        const bool have = lane_on && b < B;
        const cplx* base = mats + (size_t)(have ? b : 0) * M * d * d;
        cplx* P = gbase;
        cplx* T = gbase + L::BUF;
        cplx X[BS][BS], C[BS][BS];
        fetch(base, X, have);
        store_blk<D, BS, LD>(P + rc_off, X, lane_on);
        if (M > 1) fetch(base + (size_t)d * d, X, have);
        __syncwarp();
        for (int m = 1; m < M; ++m) {
            store_blk<D, BS, LD>(bufX + rc_off, X, lane_on);
            __syncwarp();
            if (m + 1 < M) fetch(base + (size_t)(m + 1) * d * d, X, have);      // in flight during the product
            mm_blk<D, BS, LD>(bufX + r0 * LD, P + c0, C);
            store_blk<D, BS, LD>(T + rc_off, C, lane_on);
            cplx* t = P; P = T; T = t;
            __syncwarp();
        }
        if (have) {
#pragma unroll
            for (int a = 0; a < BS; ++a) {
                const int row = r0 + a;
                if (row < d) {
#pragma unroll
                    for (int c = 0; c < BS; ++c)
                        if (c0 + c < d) out[(size_t)b * d * d + row * d + c0 + c] = P[rc_off + a * LD + c];
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace c3b
