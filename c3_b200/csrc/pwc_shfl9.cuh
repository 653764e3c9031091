// d = 9 (and zero-padded d = 7, 8... any d <= 9) fused PWC propagator kernel, third generation of the 3x3-lane block
// layout: operand blocks are EXCHANGED BY WARP SHUFFLES, never published through shared memory.
//
// Same contract as pwc_blk9_taylor_kernel (assemble -> trace-shifted degree-18 Taylor exponential in 5 products -> ordered
// product; replaces c3/libraries/propagation.py:426-440,460-515 and c3/utils/tf_utils.py:120-193).
//
// Why.  ncu on pwc_blk9_taylor_kernel (profiles/r01_prof_blk9_own_final.txt): the shared-memory / shuffle pipe (MIO, 128
// lane-bytes per clock per SM) is 83 % busy while the fp64 pipe is at 59.5 %.  Of the 1 720 wavefronts a warp spends per
// slice-triple, 868 are the operand loads of the six 9x9 products and 470 are STORES whose only purpose is to publish a
// freshly computed block to the other lanes of its group.  A SHFL.32 costs one wavefront, an LDS.128 four
// (profiles/r01_mb3_shfl.log), so moving a 16-byte element by four shuffles costs what the load alone did -- and the
// store, the bank-conflict tables and every __syncwarp disappear.  Per product and lane: 144 SHFL (4 foreign blocks x 9
// elements x 4 words) against 324 DFMA; per slice ~1 200 wavefronts against ~1 170 fp64-pipe clocks: the two pipes are
// balanced instead of MIO-bound.
//
// Mapping.  lane = 9 g + 3 bi + bj (g = 0..2: the warp's three concurrent slices, lanes 27..31 mirror lanes 0..4 and
// never store).  Lane (bi,bj) owns block (bi,bj) of every matrix of its group in registers.  C(bi,bj) = sum_k X(bi,k) Y(k,bj):
//   off-diagonal lane:  X(bi,bi) * Y(bi,bj)[own]  +  X(bi,bj)[own] * Y(bj,bj)  +  X(bi,k2) * Y(k2,bj)
//   diagonal lane:      X(bi,k1) * Y(k1,bi)       +  X(bi,bi)[own] * Y(bi,bi)[own]  +  X(bi,k2) * Y(k2,bi)
// Every lane publishes element (a,kk) of its own X block and (kk,b) of its own Y block; each lane reads two X sources and
// two Y sources.  One instruction stream: on diagonal lanes the received Y block and the own Y block swap roles (selects).
// The only shared memory left is lane-private parking (B3, B2 of the Taylor scheme and the running product) and the model.
#pragma once
#include "c3b_params.cuh"

namespace c3b {

struct Shfl9 {
    static constexpr int PARK = 27;   // lane-private complex slots: B3, B2, running product (9 each)
    __host__ __device__ static size_t smem_bytes(int K, int warps) {
        return (size_t)(K + 1) * 81 * sizeof(cplx) + (size_t)(((K + 1) * 9 + 1) & ~1) * sizeof(double) +
               (size_t)PARK * warps * 32 * sizeof(cplx);
    }
};

struct Shfl9Lane {
    int sx1, sy1, sx2, sy2;   // source lanes of the four foreign operand blocks
    bool diag;
};

__device__ __forceinline__ cplx shfl_cplx(const cplx v, const int src) {
    cplx r;
    r.x = __shfl_sync(0xffffffffu, v.x, src);
    r.y = __shfl_sync(0xffffffffu, v.y, src);
    return r;
}

// C = X * Y for this lane's block; XO / YO: own blocks of X and Y (every lane of the warp calls this together).
__device__ __forceinline__ void mm_shfl9(const cplx (&XO)[3][3], const cplx (&YO)[3][3], const Shfl9Lane& L, cplx (&c)[3][3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) c[a][b] = cmake(0.0, 0.0);
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        cplx x1[3], x2[3], y1[3], y2[3], ya[3], yb[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            x1[a] = shfl_cplx(XO[a][kk], L.sx1);
            x2[a] = shfl_cplx(XO[a][kk], L.sx2);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            y1[b] = shfl_cplx(YO[kk][b], L.sy1);
            y2[b] = shfl_cplx(YO[kk][b], L.sy2);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            ya[b].x = L.diag ? y1[b].x : YO[kk][b].x;
            ya[b].y = L.diag ? y1[b].y : YO[kk][b].y;
            yb[b].x = L.diag ? YO[kk][b].x : y1[b].x;
            yb[b].y = L.diag ? YO[kk][b].y : y1[b].y;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                cfma(c[a][b], x1[a], ya[b]);
                cfma(c[a][b], XO[a][kk], yb[b]);
                cfma(c[a][b], x2[a], y2[b]);
            }
    }
}

// generic product with all operand blocks fetched by shuffle: X block (bi,k) from lane xbase + 3 bi + k, Y block (k,bj)
// from lane ybase + 3 k + bj (fold of the three group products: once per work unit)
__device__ __forceinline__ void mm_shfl9_groups(const cplx (&Xsrc)[3][3], const int xbase, const cplx (&Ysrc)[3][3], const int ybase,
                                                const int bi, const int bj, cplx (&c)[3][3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) c[a][b] = cmake(0.0, 0.0);
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        const int sx = xbase + 3 * bi + k, sy = ybase + 3 * k + bj;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            cplx x[3], y[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) x[a] = shfl_cplx(Xsrc[a][kk], sx);
#pragma unroll
            for (int b = 0; b < 3; ++b) y[b] = shfl_cplx(Ysrc[kk][b], sy);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) cfma(c[a][b], x[a], y[b]);
        }
    }
}

template <int WARPS, int MINB, int GATED>
__global__ void __launch_bounds__(WARPS * 32, MINB) pwc_shfl9_kernel(const RowsParams p, unsigned int* __restrict__ counter) {
    constexpr int D = 9, NT = WARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);                 // [(K+1)][element 0..8][block 0..8], zero padded
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * 81);
    cplx* park = reinterpret_cast<cplx*>(sRS + (((K + 1) * D + 1) & ~1)) + threadIdx.x;   // [slot][thread]: lane-private

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool hmode = p.hlist != nullptr;

    if (!hmode) {
        for (int idx = tid; idx < (K + 1) * D * D; idx += NT) {
            const int k = idx / (D * D);
            const int rem = idx - k * D * D;
            const int r = rem / D, j = rem - r * D;
            cplx v = cmake(0.0, 0.0);
            if (r < d && j < d) v = p.G[(size_t)k * d * d + r * d + j];
            sG[k * 81 + ((r % 3) * 3 + (j % 3)) * 9 + (r / 3) * 3 + j / 3] = v;
        }
        for (int idx = tid; idx < (K + 1) * D; idx += NT) {
            const int k = idx / D, r = idx - k * D;
            sRS[idx] = (r < d) ? p.RS[k * d + r] : 0.0;
        }
    }
    __syncthreads();

    const bool lane_on = lane < 27;
    const int src = lane_on ? lane : lane - 27;
    const int g = src / 9;                     // lane group = the slice chunk this lane works on
    const int li = src - g * 9;                // block owned
    const int bi = li / 3, bj = li - bi * 3;
    const int r0 = bi * 3, c0 = bj * 3;
    Shfl9Lane L;
    L.diag = (bi == bj);
    {
        int kx1, ky1, k2;
        if (!L.diag) { kx1 = bi; ky1 = bj; k2 = 3 - bi - bj; }
        else { kx1 = ky1 = (bi + 1) % 3; k2 = (bi + 2) % 3; }
        L.sx1 = g * 9 + bi * 3 + kx1;
        L.sy1 = g * 9 + ky1 * 3 + bj;
        L.sx2 = g * 9 + bi * 3 + k2;
        L.sy2 = g * 9 + k2 * 3 + bj;
    }
    const bool on_diag = L.diag;
    const cplx hs = cmake(p.hscale_re, p.hscale_im);
    const long long total_units = (long long)p.B * p.S;
    const bool shifted = (p.TR != nullptr) && !hmode;
    const cplx* gown = sG + li;                // own block of the generators: gown[(k * 9 + e) * 9]
    cplx* parkB3 = park;                       // slots 0..8
    cplx* parkB2 = park + 9 * NT;              // slots 9..17
    cplx* parkP = park + 18 * NT;              // slots 18..26

    // Lockstep work distribution: the CTA takes WARPS consecutive units at a time (warp w: base + w), every warp runs the
    // same number of slice iterations (ceil(seg_len / 3); lanes past the end of a short last segment multiply by the
    // identity) and the warps meet at a barrier before every slice.
    __shared__ unsigned int s_base;
    const int warp = tid >> 5;
    const int cl_all = (p.seg_len + 2) / 3;
    const bool late = warp >= WARPS / 2;
    unsigned int rows_known = 0;            // gated launch: batch rows this warp knows to have landed
    for (;;) {
        __syncthreads();
        if (tid == 0) s_base = atomicAdd(counter, (unsigned int)WARPS);
        __syncthreads();
        const long long base_unit = s_base;
        if (base_unit >= total_units) break;
        const long long unit = base_unit + warp;
        bool live = unit < total_units;                      // warps past the end of the work list idle through the barriers
        const int b = live ? (int)(unit / p.S) : 0;
        if constexpr (GATED != 0) {
            // gated launch: wait (all lanes, uniform code) until this unit's batch row has landed; rows arrive in order.
            // A row that never arrives raises gate[1]; the warp then idles through the barriers and writes nothing.
            if (live && !wait_rows_ready(p.gate, b, rows_known)) live = false;
        }
        const int sidx = live ? (int)(unit - (long long)b * p.S) : 0;
        const int n_begin = sidx * p.seg_len;
        const int n_end = live ? min(p.N, n_begin + p.seg_len) : n_begin;
        const int len = n_end - n_begin;
        const int cl = (len + 2) / 3;
        const int my_begin = n_begin + g * cl;
        const int my_end = min(n_end, my_begin + cl);
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        cplx mu_acc = cmake(0.0, 0.0);

#pragma unroll 1
        for (int it = 0; it < cl_all; ++it) {
            __syncthreads();                                 // lockstep: see the comment at the product chain
            if (late) {                                      // ... with the two warps of a scheduler a fraction of a product apart
                const long long t0 = clock64();
                while (clock64() - t0 < p.skew) {}
            }
            const int n = my_begin + it;
            const bool on = lane_on && (it < cl) && (n < my_end);

            // ---- assemble the own block of A_n = G_0 + sum_k c_k[n] G_k, its norm bound and trace shift ------------
            cplx A[3][3];
            cplx mu = cmake(0.0, 0.0);
            double nb = 0.0;
            if (!hmode) {
                double nba[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    nba[a] = on ? sRS[r0 + a] : 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) A[a][c] = on ? gown[(a * 3 + c) * 9] : cmake(0.0, 0.0);
                }
                if (shifted && on) mu = p.TR[0];
                for (int k = 0; k < K; ++k) {
                    const double cs = on ? load_signal<GATED>(sig_b + (size_t)k * p.N + n) : 0.0;
                    const cplx* gk = gown + (k + 1) * 81;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx gv = gk[(a * 3 + c) * 9];
                            A[a][c].x = fma(cs, gv.x, A[a][c].x);
                            A[a][c].y = fma(cs, gv.y, A[a][c].y);
                        }
                        nba[a] = fma(fabs(cs), sRS[(k + 1) * D + r0 + a], nba[a]);
                    }
                    if (shifted) { const cplx t = p.TR[k + 1]; mu.x = fma(cs, t.x, mu.x); mu.y = fma(cs, t.y, mu.y); }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) nb = fmax(nb, nba[a]);
            } else {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int row = r0 + a;
                    double rs = 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        cplx h = cmake(0.0, 0.0);
                        if (on && row < d && c0 + c < d)
                            h = p.hlist[((size_t)b * p.N + n) * d * d + (size_t)row * d + c0 + c];
                        A[a][c] = cmul(hs, h);
                        rs += cabs1(A[a][c]);
                    }
                    // row sum over the three lanes of this block row
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < 3; ++q) tot += __shfl_sync(0xffffffffu, rs, g * 9 + bi * 3 + q);
                    nb = fmax(nb, tot);
                }
            }
            // warp-wide upper bound of the norm estimates in ONE redux.sync: nb >= 0, so the high words order like the
            // doubles; rounding the high word up keeps it an upper bound (relative slack 2^-20)
            nb = __hiloint2double((int)__reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(nb) + 1u), 0);
            mu_acc.x += mu.x; mu_acc.y += mu.y;

            const int s = squarings_for(nb, C3B_THETA18);
            if (s > 0) {
                const double sc = pow2neg(s);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) { A[a][c].x *= sc; A[a][c].y *= sc; }
            }

            // ---- degree-18 Taylor polynomial in 5 products (c3b_common.cuh) -------------------------------------------
            // Straight-line code: every product is its own inlined instance, so no operand block is ever copied between
            // registers -- at the price of a 65 KB loop body, twice the instruction cache.  The CTA's warps therefore
            // run the slice loop in LOCKSTEP (one bar.sync per slice, see below): the SM streams the body once per slice
            // for all of its warps instead of once per warp.
            cplx A2[3][3], A3[3][3], C[3][3];
            mm_shfl9(A, A, L, A2);
            mm_shfl9(A2, A, L, A3);
            mm_shfl9(A3, A3, L, C);                             // A^6
            cplx B1[3][3], B5[3][3], B4[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const cplx x1 = A[a][c], x2 = A2[a][c], x3 = A3[a][c], x6 = C[a][c];
                    const double dg = (on_diag && a == c) ? 1.0 : 0.0;
                    cplx b3, b2;
                    B1[a][c].x = C3B_T18_A11 * x1.x + C3B_T18_A21 * x2.x + C3B_T18_A31 * x3.x;
                    B1[a][c].y = C3B_T18_A11 * x1.y + C3B_T18_A21 * x2.y + C3B_T18_A31 * x3.y;
                    B5[a][c].x = C3B_T18_B24 * x2.x + C3B_T18_B34 * x3.x + C3B_T18_B64 * x6.x;
                    B5[a][c].y = C3B_T18_B24 * x2.y + C3B_T18_B34 * x3.y + C3B_T18_B64 * x6.y;
                    B4[a][c].x = C3B_T18_B03 * dg + C3B_T18_B13 * x1.x + C3B_T18_B23 * x2.x + C3B_T18_B33 * x3.x + C3B_T18_B63 * x6.x;
                    B4[a][c].y = C3B_T18_B13 * x1.y + C3B_T18_B23 * x2.y + C3B_T18_B33 * x3.y + C3B_T18_B63 * x6.y;
                    b3.x = C3B_T18_B02 * dg + C3B_T18_B12 * x1.x + C3B_T18_B22 * x2.x + C3B_T18_B32 * x3.x + C3B_T18_B62 * x6.x;
                    b3.y = C3B_T18_B12 * x1.y + C3B_T18_B22 * x2.y + C3B_T18_B32 * x3.y + C3B_T18_B62 * x6.y;
                    b2.x = C3B_T18_B11 * x1.x + C3B_T18_B21 * x2.x + C3B_T18_B31 * x3.x + C3B_T18_B61 * x6.x;
                    b2.y = C3B_T18_B11 * x1.y + C3B_T18_B21 * x2.y + C3B_T18_B31 * x3.y + C3B_T18_B61 * x6.y;
                    parkB3[(a * 3 + c) * NT] = b3;              // lane-private: re-read after the next product
                    parkB2[(a * 3 + c) * NT] = b2;              // re-read after the last Taylor product
                }
            mm_shfl9(B1, B5, L, C);                             // B1 B5
            cplx A9[3][3], LH[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const cplx b3 = parkB3[(a * 3 + c) * NT];
                    A9[a][c] = cmake(C[a][c].x + B4[a][c].x, C[a][c].y + B4[a][c].y);
                    LH[a][c] = cmake(b3.x + A9[a][c].x, b3.y + A9[a][c].y);
                }
            cplx E[3][3];
            mm_shfl9(LH, A9, L, E);                             // (B3 + A9) A9
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const cplx b2 = parkB2[(a * 3 + c) * NT];
                    E[a][c].x += b2.x; E[a][c].y += b2.y;       // T18 = exp(A_n / 2^s)
                }
#pragma unroll 1
            for (int q = 0; q < s; ++q) {                       // undo the scaling: square s times
                cplx T[3][3];
                mm_shfl9(E, E, L, T);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) E[a][c] = T[a][c];
            }
            if (p.dUs_out != nullptr && on) {
                const cplx ph_n = shifted ? cexp_(mu) : cmake(1.0, 0.0);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int row = r0 + a;
                    if (row < d) {
                        cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)row * d + c0;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            if (c0 + c < d) o[c] = shifted ? cmul(ph_n, E[a][c]) : E[a][c];
                    }
                }
            }
            // ---- running product of this group's slices: P <- dU_n P ---------------------------------------------------
            if (it == 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) parkP[(a * 3 + c) * NT] = E[a][c];
            } else {
                cplx P[3][3], T[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) P[a][c] = parkP[(a * 3 + c) * NT];
                mm_shfl9(E, P, L, T);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) parkP[(a * 3 + c) * NT] = T[a][c];
            }
        }

        // ---- this group's product with its accumulated shift re-applied: P_g <- exp(sum mu) P_g ---------------------
        cplx P[3][3];
        {
            const cplx ph_g = shifted ? cexp_(mu_acc) : cmake(1.0, 0.0);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const cplx v = parkP[(a * 3 + c) * NT];
                    P[a][c] = shifted ? cmul(ph_g, v) : v;
                }
        }
        // ---- fold the group products: P_2 P_1 P_0 (every lane computes block (bi,bj) of it; group 0 writes) ---------
        cplx T1[3][3], T2[3][3];
        mm_shfl9_groups(P, 18, P, 9, bi, bj, T1);               // P_2 P_1
        mm_shfl9_groups(T1, g * 9, P, 0, bi, bj, T2);           // (P_2 P_1) P_0
        if (live && lane_on && g == 0) {
            cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int row = r0 + a;
                if (row < d) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (c0 + c < d) o[row * d + c0 + c] = T2[a][c];
                }
            }
        }
    }
}

}  // namespace c3b
