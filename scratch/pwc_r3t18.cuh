// R-rows-per-lane PWC propagator kernel with the degree-18 Taylor exponential.
//
// Same fused contract as pwc_blk.cuh (assemble -> expm -> ordered product; replaces
// c3/libraries/propagation.py:426-440,460-515 and c3/utils/tf_utils.py:120-193) on the lane mapping of
// pwc_rows3_kernel: a lane owns R rows of every matrix (d=9: R=3 -> 3 lanes per matrix, 10 matrices
// per warp, 30/32 lanes busy).  The LEFT operand of every product is the lane's own rows and stays in
// registers (it is always either the previous product or an element-wise combination the lane just
// formed), the RIGHT operand streams from shared memory with group-broadcast LDS.128: each 16-byte
// operand feeds R = 3 complex MACs (the 3x3-block kernel: 1.5), so the shared-memory wavefront pipe
// (128 lane-bytes/clk/SM) is no longer the limiter: per k step 9 loads = 36 MIO clocks against
// 27 cfma = 54 fp64-pipe clocks.
//
// What made this mapping lose with Pade (profiles/README_r01.md) was the serial Gauss-Jordan sweep at
// one warp per scheduler.  The Taylor scheme T18 (c3b_common.cuh) has none: every phase is a dense
// product or an element-wise combination, and its live set fits FOUR shared buffers per matrix
//   bufA: A -> B3 (own rows)   bufB: A2 -> B2 (own rows)   bufX: A3 -> B5 -> A9 -> squaring operand
//   bufP: running product
// because B1 and (B3 + A9) go straight into the left-operand registers and B4 / B2 pre-load the
// accumulators of the products they are added to.  4 x 1296 B x 10 matrices = 52 KB per warp ->
// 4 warps (one per scheduler) per SM; the 27 independent accumulators per lane cover the DFMA and
// LDS latencies by instruction-level parallelism.
#pragma once
#include "../c3_b200/csrc/c3b_common.cuh"
#include "../c3_b200/csrc/pwc_rows.cuh"   // RowsParams, Rows3Layout, store_rows

namespace c3b {

// keeps ptxas from hoisting every shared load of an unrolled element-wise phase above the first use
// (which costs ~200 live registers on top of the two 108-register operand arrays)
#define C3B_SCHED_FENCE() asm volatile("" ::: "memory")

// c += x * Y  (x: the lane's R rows in registers, Y: D x D in shared memory, broadcast reads)
template <int D, int R>
__device__ __forceinline__ void mm_rows_acc(const cplx (&x)[R][D], const cplx* __restrict__ Y, cplx (&c)[R][D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const cplx y = Y[k * D + j];
#pragma unroll
            for (int a = 0; a < R; ++a) cfma(c[a][j], x[a][k], y);
        }
    }
}

template <int D, int R>
__device__ __forceinline__ void zero_rows(cplx (&c)[R][D]) {
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) c[a][j] = cmake(0.0, 0.0);
}

template <int D, int R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pwc_r3t18_kernel(const RowsParams p, unsigned int* __restrict__ counter) {
    using L = Rows3Layout<D, R>;
    constexpr int LPM = L::LPM, MPW = L::MPW, DP = LPM * R;   // DP: padded row count
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);                   // [(K+1), DP, D] zero padded
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * DP * D); // [(K+1), DP]
    cplx* sWarps = reinterpret_cast<cplx*>(sRS + (((K + 1) * DP + 1) & ~1));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool hmode = p.hlist != nullptr;

    if (!hmode) {
        for (int idx = tid; idx < (K + 1) * DP * D; idx += WARPS * 32) {
            const int k = idx / (DP * D);
            const int rem = idx - k * DP * D;
            const int r = rem / D, j = rem - r * D;
            cplx v = cmake(0.0, 0.0);
            if (r < d && j < d) v = p.G[(size_t)k * d * d + r * d + j];
            sG[idx] = v;
        }
        for (int idx = tid; idx < (K + 1) * DP; idx += WARPS * 32) {
            const int k = idx / DP, r = idx - k * DP;
            sRS[idx] = (r < d) ? p.RS[k * d + r] : 0.0;
        }
    }
    __syncthreads();   // the only CTA-wide barrier; warps are independent from here on

    const int g_raw = lane / LPM;
    const bool lane_on = g_raw < MPW;                 // leftover lanes shadow group 0 and never store
    const int g = lane_on ? g_raw : 0;
    const int l = lane_on ? (lane - g_raw * LPM) : 0; // lane within the group
    const int row0 = l * R;                           // first row owned by this lane

    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + (size_t)g * L::GROUP_ELEMS;
    cplx* bufA = gbase;
    cplx* bufB = gbase + L::BUF;
    cplx* bufX = gbase + 2 * L::BUF;
    cplx* bufP = gbase + 3 * L::BUF;
    cplx* ownA = bufA + row0 * D;                     // this lane's rows inside each buffer
    cplx* ownB = bufB + row0 * D;
    cplx* ownX = bufX + row0 * D;
    cplx* ownP = bufP + row0 * D;
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
        const int cl = (len + MPW - 1) / MPW;          // slices per lane group
        const int my_begin = n_begin + g * cl;
        const int my_end = min(n_end, my_begin + cl);
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        cplx mu_acc = cmake(0.0, 0.0);                 // sum of the trace shifts of this group's slices

#pragma unroll 1
        for (int it = 0; it < cl; ++it) {
            const int n = my_begin + it;
            const bool on = lane_on && (n < my_end);

            // ---- assemble this lane's rows of A_n, the inf-norm bound and the trace shift ----------
            cplx Xop[R][D];
            cplx mu = cmake(0.0, 0.0);
            double nb = 0.0;
            if (!hmode) {
#pragma unroll
                for (int a = 0; a < R; ++a) {
#pragma unroll
                    for (int j = 0; j < D; ++j) Xop[a][j] = on ? sG[(row0 + a) * D + j] : cmake(0.0, 0.0);
                }
                double nba[R];
#pragma unroll
                for (int a = 0; a < R; ++a) nba[a] = on ? sRS[row0 + a] : 0.0;
                if (shifted && on) mu = p.TR[0];
                for (int k = 0; k < K; ++k) {
                    const double c = on ? __ldg(sig_b + (size_t)k * p.N + n) : 0.0;
                    const cplx* gk = sG + (k + 1) * DP * D + row0 * D;
#pragma unroll
                    for (int a = 0; a < R; ++a) {
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            const cplx gv = gk[a * D + j];
                            Xop[a][j].x = fma(c, gv.x, Xop[a][j].x);
                            Xop[a][j].y = fma(c, gv.y, Xop[a][j].y);
                        }
                        nba[a] = fma(fabs(c), sRS[(k + 1) * DP + row0 + a], nba[a]);
                    }
                    if (shifted) { const cplx t = p.TR[k + 1]; mu.x = fma(c, t.x, mu.x); mu.y = fma(c, t.y, mu.y); }
                }
#pragma unroll
                for (int a = 0; a < R; ++a) nb = fmax(nb, nba[a]);
            } else {
#pragma unroll
                for (int a = 0; a < R; ++a) {
                    const int row = row0 + a;
                    const cplx* hrow = p.hlist + ((size_t)b * p.N + (on ? n : 0)) * d * d + (size_t)(row < d ? row : 0) * d;
                    double rs = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        cplx h = cmake(0.0, 0.0);
                        if (on && row < d && j < d) h = hrow[j];
                        Xop[a][j] = cmul(hs, h);
                        rs += cabs1(Xop[a][j]);
                    }
                    nb = fmax(nb, rs);
                }
            }
            nb = warp_max(nb);
            mu_acc.x += mu.x; mu_acc.y += mu.y;

            const int s = squarings_for(nb, C3B_THETA18);
            if (s > 0) {
                const double sc = pow2neg(s);
#pragma unroll
                for (int a = 0; a < R; ++a)
#pragma unroll
                    for (int j = 0; j < D; ++j) { Xop[a][j].x *= sc; Xop[a][j].y *= sc; }
            }
            store_rows<D, R>(bufA, row0, Xop, lane_on);
            __syncwarp();

            // Straight-line phases (no rolled phase loop: the two register arrays P0 / P1 swap the roles of
            // left operand and accumulator from product to product, so no copies and no fixed-register
            // constraint at a loop back-edge):
            //   A2 = A A | A3 = A2 A | A6 = A3 A3 (+ combinations) | A9 = B4 + B1 B5 | T18 = B2 + (B3 + A9) A9
            //   | s squarings | running product
            cplx (&P0)[R][D] = Xop;
            cplx P1[R][D];
            zero_rows<D, R>(P1);
            mm_rows_acc<D, R>(P0, bufA, P1);                    // P1 = A^2 (only ever read by its owner)
            store_rows<D, R>(bufB, row0, P1, lane_on);
            zero_rows<D, R>(P0);
            mm_rows_acc<D, R>(P1, bufA, P0);                    // P0 = A^3 = A^2 A
            store_rows<D, R>(bufX, row0, P0, lane_on);
            __syncwarp();
            zero_rows<D, R>(P1);
            mm_rows_acc<D, R>(P0, bufX, P1);                    // P1 = A^6 = A^3 A^3
            __syncwarp();                                       // bufX (A^3) fully read by the group
#pragma unroll
            for (int a = 0; a < R; ++a) {
                const int row = row0 + a;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const cplx x1 = ownA[a * D + j], x2 = ownB[a * D + j], x3 = P0[a][j], x6 = P1[a][j];
                    const double dg = (j == row) ? 1.0 : 0.0;
                    cplx b1, b5, b4, b3, b2;
                    b1.x = C3B_T18_A11 * x1.x + C3B_T18_A21 * x2.x + C3B_T18_A31 * x3.x;
                    b1.y = C3B_T18_A11 * x1.y + C3B_T18_A21 * x2.y + C3B_T18_A31 * x3.y;
                    b5.x = C3B_T18_B24 * x2.x + C3B_T18_B34 * x3.x + C3B_T18_B64 * x6.x;
                    b5.y = C3B_T18_B24 * x2.y + C3B_T18_B34 * x3.y + C3B_T18_B64 * x6.y;
                    b4.x = C3B_T18_B03 * dg + C3B_T18_B13 * x1.x + C3B_T18_B23 * x2.x + C3B_T18_B33 * x3.x + C3B_T18_B63 * x6.x;
                    b4.y = C3B_T18_B13 * x1.y + C3B_T18_B23 * x2.y + C3B_T18_B33 * x3.y + C3B_T18_B63 * x6.y;
                    b3.x = C3B_T18_B02 * dg + C3B_T18_B12 * x1.x + C3B_T18_B22 * x2.x + C3B_T18_B32 * x3.x + C3B_T18_B62 * x6.x;
                    b3.y = C3B_T18_B12 * x1.y + C3B_T18_B22 * x2.y + C3B_T18_B32 * x3.y + C3B_T18_B62 * x6.y;
                    b2.x = C3B_T18_B11 * x1.x + C3B_T18_B21 * x2.x + C3B_T18_B31 * x3.x + C3B_T18_B61 * x6.x;
                    b2.y = C3B_T18_B11 * x1.y + C3B_T18_B21 * x2.y + C3B_T18_B31 * x3.y + C3B_T18_B61 * x6.y;
                    P0[a][j] = b1;                              // left operand of B1 B5
                    P1[a][j] = b4;                              // accumulator of A9 = B4 + B1 B5
                    if (lane_on && row < D) {
                        ownX[a * D + j] = b5;                   // right operand
                        ownA[a * D + j] = b3;                   // own rows only, re-read after the product
                        ownB[a * D + j] = b2;
                    }
                }
            }
            __syncwarp();
            mm_rows_acc<D, R>(P0, bufX, P1);                    // P1 = A9
            __syncwarp();                                       // bufX (B5) fully read
#pragma unroll
            for (int a = 0; a < R; ++a) {
                const int row = row0 + a;
                const int rr = row < D ? a : 0;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const cplx a9 = P1[a][j];
                    const cplx b3 = ownA[rr * D + j], b2 = ownB[rr * D + j];
                    if (lane_on && row < D) ownX[a * D + j] = a9;   // right operand A9
                    P0[a][j] = cmake(b3.x + a9.x, b3.y + a9.y);     // left operand B3 + A9
                    P1[a][j] = b2;                                  // accumulator of T18 = B2 + (B3 + A9) A9
                }
            }
            __syncwarp();
            mm_rows_acc<D, R>(P0, bufX, P1);                    // P1 = T18(A_n / 2^s)
#pragma unroll 1
            for (int q = 0; q < s; ++q) {                       // P1 <- P1^2
                __syncwarp();                                   // bufX fully read
                store_rows<D, R>(bufX, row0, P1, lane_on);
                __syncwarp();
#pragma unroll
                for (int a = 0; a < R; ++a)
#pragma unroll
                    for (int j = 0; j < D; ++j) { P0[a][j] = P1[a][j]; P1[a][j] = cmake(0.0, 0.0); }
                mm_rows_acc<D, R>(P0, bufX, P1);
            }
            // P1 = dU_n (up to the scalar exp(mu_n))
            if (p.dUs_out != nullptr && on) {
                const cplx ph_n = shifted ? cexp_(mu) : cmake(1.0, 0.0);
#pragma unroll
                for (int a = 0; a < R; ++a) {
                    const int row = row0 + a;
                    if (row < d) {
                        cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)row * d;
#pragma unroll
                        for (int j = 0; j < D; ++j)
                            if (j < d) o[j] = shifted ? cmul(ph_n, P1[a][j]) : P1[a][j];
                    }
                }
            }
            if (it == 0) {
                store_rows<D, R>(bufP, row0, P1, lane_on);
            } else {
                zero_rows<D, R>(P0);
                mm_rows_acc<D, R>(P1, bufP, P0);                // P0 = dU_n * P
                __syncwarp();                                   // everyone has finished reading the old P
                store_rows<D, R>(bufP, row0, P0, lane_on);
            }
        }
        __syncwarp();

        // ---- re-apply this group's accumulated shift: P_g <- exp(sum mu) P_g -------------------------
        if (shifted) {
            const cplx ph_g = cexp_(mu_acc);
#pragma unroll
            for (int a = 0; a < R; ++a) {
                if (lane_on && row0 + a < D) {
#pragma unroll
                    for (int j = 0; j < D; ++j) ownP[a * D + j] = cmul(ph_g, ownP[a * D + j]);
                }
            }
            __syncwarp();
        }

        // ---- fold the MPW group products of this warp (pairwise tree, later chunks on the left) --
        // level with stride st: group g computes M[hi] * M[lo] with lo = g rounded down to a
        // multiple of 2 st, hi = lo + st (if it exists); every group stores into its own bufX/bufA.
        cplx* wbase = sWarps + (size_t)warp * L::WARP_ELEMS;
        int cur = 3;   // buffer index holding each group's current partial product (3 = bufP)
#pragma unroll 1
        for (int st = 1; st < MPW; st <<= 1) {
            const int lo = g & ~(2 * st - 1);
            const int hi = lo + st;
            const int nxtbuf = (cur == 2) ? 0 : 2;
            cplx T[R][D];
            if (hi < MPW) {
                cplx Xh[R][D];
                const cplx* Mh = wbase + (size_t)hi * L::GROUP_ELEMS + cur * L::BUF;
#pragma unroll
                for (int a = 0; a < R; ++a) {
                    const int rr = (row0 + a) < D ? (row0 + a) : 0;
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        Xh[a][j] = (row0 + a) < D ? Mh[rr * D + j] : cmake(0.0, 0.0);
                        T[a][j] = cmake(0.0, 0.0);
                    }
                }
                mm_rows_acc<D, R>(Xh, wbase + (size_t)lo * L::GROUP_ELEMS + cur * L::BUF, T);
            } else {
                const cplx* Ml = wbase + (size_t)lo * L::GROUP_ELEMS + cur * L::BUF;
#pragma unroll
                for (int a = 0; a < R; ++a) {
                    const int rr = (row0 + a) < D ? (row0 + a) : 0;
#pragma unroll
                    for (int j = 0; j < D; ++j) T[a][j] = Ml[rr * D + j];
                }
            }
            store_rows<D, R>(gbase + nxtbuf * L::BUF, row0, T, lane_on);
            __syncwarp();
            cur = nxtbuf;
        }
        if (lane_on && g == 0) {
            const cplx* fin = gbase + cur * L::BUF;
            cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
            for (int a = 0; a < R; ++a) {
                const int row = row0 + a;
                if (row < d) {
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j < d) o[row * d + j] = fin[row * D + j];
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace c3b
