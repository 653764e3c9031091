// Fused gradient kernel for the closed-system d = 9 (and zero-padded d = 7, 8) PWC propagator with anti-Hermitian slice
// generators (Hermitian Hamiltonians: every C3 model) -- SURVEY.md section 8f, row f-1: what tf.GradientTape provides to
// the reference's optimisers (c3/optimizers/optimizer.py:210-215, 277-313) for the propagators of
// c3/libraries/propagation.py:426-440,460-515.
//
//   U = dU_{N-1} ... dU_0,  dU_n = exp(A_n),  A_n = G_0 + sum_k c_k[n] G_k,      dL = Re tr(Ubar^dag dU)
//   dL/dc_k[n] = Re tr( M_n L(A_n, G_k) ),   M_n = F_n Ubar^dag R_n,   F_n = dU_{n-1}..dU_0,   R_n = dU_{N-1}..dU_{n+1}
//
// (grad.cuh).  With  Y_n = M_n dU_n = F_n Ubar^dag U F_n^{-1}  and  tr(Y e^{-A} L(A, G)) = tr(L(A, Y) e^{-A} G):
//
//   dL/dc_k[n] = Re tr( V_n G_k ),   V_n = L(A_n, Y_n) dU_n^dag,      Y_{n+1} = dU_n Y_n dU_n^dag,   Y_0 = Ubar^dag U
//
// for UNITARY dU_n.  Everything about slice n is then a function of A_n and the running Y_n: no stored partial
// propagators, no backward sweep, no second pass over memory -- the same streaming structure as the forward kernel
// (pwc_blk9.cuh), whose lane layout, tables and own-block product this kernel reuses.  A lane group walks a chunk of CL
// consecutive slices; Y at the chunk boundaries comes from the chunk products of one forward launch
// (grad9_prefix_kernel + grad9_ybound_kernel).  The trace shift drops out (|e^mu| = 1).
//
// Per slice: the Frechet derivative of the four-product Taylor scheme (c3b_common.cuh) in direction Y, in lockstep with
// the scheme itself --
//   A2 = A A               dA2 = A Y + Y A
//   P0 = A2 Q0             dP0 = dA2 Q0 + A2 dQ0             Q0 = a1 A2 + a2 A
//   P1 = L1 R1 + b5 P0     dP1 = dL1 R1 + L1 dR1 + b5 dP0    L1 = P0 + b1 A2 + b2 A,  R1 = P0 + b3 A2 + b4 I
//   T  = L2 R2 + E0        dT  = dL2 R2 + L2 dR2 + dE0       L2, R2, E0 as in c3b_common.cuh
// -- then V = dT T^dag, the K contractions, and Y <- T Y T^dag: 15 products of 9 x 9 per slice (+ 3 per squaring).
// Seven matrix buffers per lane group (8 warps x 3 groups x 7 x 1296 B = 214 KB per SM): operands are overwritten as
// soon as they are dead, and the own blocks of A2 / P0 (dA2 / dP0) needed by the later combinations are recovered from
// L1, R1, A (dL1, dR1, Y) instead of being kept:  A2 = (L1 - R1' - b2 A) / (b1 - b3),  P0 = R1' - b3 A2  (R1' = R1 - b4 I).
#pragma once
#include "pwc_blk9.cuh"

namespace c3b {

struct Grad9T {
    static constexpr int BUF = Blk9::BUF;
    static constexpr int NBUF = 7;
    // group bases: the residues (mod 8) of the annealed forward layout (Blk9T<true>: 0, 409, 821).  Measured and dropped:
    // the select-based tables (every lane fetches its own block of Y: 4 wavefronts instead of 6) + 6 %, skipping the
    // operand fetches whose block is still in registers behind run-time flags + 3 %.  A layout searched for THIS kernel's mixed
    // operand fetch (scratch/conflict_blockmajor.py, mode "grad") gets it from 6 to 5 wavefronts at best, 2.8 % of the
    // kernel's shared-memory traffic (the kernel runs the MIO pipe at 92 %): not worth a third table set.
    static constexpr int G1 = 569, G2 = 1141, WARP_ELEMS = 1712;
    static_assert(G1 % 8 == 409 % 8 && G2 % 8 == 821 % 8 && WARP_ELEMS % 8 == 0, "bank residues of the searched layout");
    static_assert(G1 >= NBUF * BUF && G2 >= G1 + NBUF * BUF && WARP_ELEMS >= G2 + NBUF * BUF, "layout");
    __host__ __device__ static constexpr int group_off(int g) { return g == 0 ? 0 : (g == 1 ? G1 : G2); }
    __host__ __device__ static size_t smem_bytes(int K, int warps) {
        size_t model = (size_t)(K + 1) * BUF * sizeof(cplx) + (size_t)(((K + 1) * 9 + 1) & ~1) * sizeof(double);
        return model + (size_t)warps * WARP_ELEMS * sizeof(cplx);
    }
};

// Y at the chunk boundaries, in two kernels.  (1) grad9_prefix_kernel, one warp per batch row: the prefix products
// F_0 = I, F_{q+1} = E_q F_q of the chunk products E_q of the forward launch -- the only sequential part, ONE product per chunk
// with the next E_q prefetched into registers -- and U = F_Q, which replaces the forward pass's own fold.  (2)
// grad9_ybound_kernel, one warp per (row, chunk), fully parallel:  Y_q = F_q (Ubar^dag U) F_q^dag  (unit-modulus phases of the
// trace shift cancel).  Three d x d matrices in shared memory per warp; element (i, j) of a product is one lane's dot product.
__device__ __forceinline__ void warp_mm(cplx* __restrict__ C, const cplx* __restrict__ A, const cplx* __restrict__ B, const int d,
                                        const int lane, const int mode) {
    // mode 0: C = A B;  1: C = A B^dag;  2: C = A^dag B
    for (int e = lane; e < d * d; e += 32) {
        const int i = e / d, j = e - i * d;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k < d; ++k) {
            const cplx a = (mode == 2) ? A[k * d + i] : A[i * d + k];
            const cplx bb = (mode == 1) ? B[j * d + k] : B[k * d + j];
            const double ay = (mode == 2) ? -a.y : a.y, by = (mode == 1) ? -bb.y : bb.y;
            acc.x += a.x * bb.x - ay * by;
            acc.y += a.x * by + ay * bb.x;
        }
        C[e] = acc;
    }
}

__global__ void grad9_prefix_kernel(const cplx* __restrict__ seg, cplx* __restrict__ F, cplx* __restrict__ U, const int B, const int Q,
                                    const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= B) return;
    const int dd = d * d;
    cplx* P = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 3 * dd;
    cplx* E = P + dd;
    cplx* T = E + dd;
    for (int e = lane; e < dd; e += 32) P[e] = cmake((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
    constexpr int kPre = 32;                               // dd <= 1024: at most 32 elements per lane
    cplx nxt[kPre];
    auto prefetch = [&](const int q) {
        const cplx* Eq = seg + ((size_t)b * Q + q) * dd;
#pragma unroll
        for (int u = 0; u < kPre; ++u) { const int e = lane + u * 32; if (e < dd) nxt[u] = Eq[e]; }
    };
    prefetch(0);
    cplx* Fb = F + (size_t)b * Q * dd;
    for (int q = 0; q < Q; ++q) {
        __syncwarp();
        for (int e = lane; e < dd; e += 32) Fb[(size_t)q * dd + e] = P[e];          // F_q
#pragma unroll
        for (int u = 0; u < kPre; ++u) { const int e = lane + u * 32; if (e < dd) E[e] = nxt[u]; }
        __syncwarp();
        if (q + 1 < Q) prefetch(q + 1);
        warp_mm(T, E, P, d, lane, 0);                      // F_{q+1} = E_q F_q
        cplx* t = P; P = T; T = t;
    }
    __syncwarp();
    for (int e = lane; e < dd; e += 32) U[(size_t)b * dd + e] = P[e];
}

__global__ void grad9_ybound_kernel(const cplx* __restrict__ F, const cplx* __restrict__ U, const cplx* __restrict__ Ubar,
                                    cplx* __restrict__ Ybound, const int B, const int Q, const int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (w >= (long long)B * Q) return;
    const int b = (int)(w / Q);
    const int dd = d * d;
    cplx* X = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * 3 * dd;
    cplx* Y = X + dd;
    cplx* Z = Y + dd;
    for (int e = lane; e < dd; e += 32) { X[e] = Ubar[(size_t)b * dd + e]; Y[e] = U[(size_t)b * dd + e]; }
    __syncwarp();
    warp_mm(Z, X, Y, d, lane, 2);                          // C = Ubar^dag U
    __syncwarp();
    for (int e = lane; e < dd; e += 32) X[e] = F[(size_t)w * dd + e];
    __syncwarp();
    warp_mm(Y, X, Z, d, lane, 0);                          // F C
    __syncwarp();
    warp_mm(Z, Y, X, d, lane, 1);                          // (F C) F^dag
    __syncwarp();
    for (int e = lane; e < dd; e += 32) Ybound[(size_t)w * dd + e] = Z[e];
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) grad_blk9_kernel(const Grad9Params p, unsigned int* __restrict__ counter) {
    using LY = Grad9T;
    using TB = Blk9Tab<true>;
    constexpr int D = 9, S = Blk9::S, BUF = Blk9::BUF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);                 // [(K+1)] element-major, zero padded
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * BUF);
    cplx* sWarps = reinterpret_cast<cplx*>(sRS + (((K + 1) * D + 1) & ~1));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    for (int idx = tid; idx < (K + 1) * D * D; idx += WARPS * 32) {
        const int k = idx / (D * D);
        const int rem = idx - k * D * D;
        const int r = rem / D, j = rem - r * D;
        cplx v = cmake(0.0, 0.0);
        if (r < d && j < d) v = p.G[(size_t)k * d * d + r * d + j];
        sG[k * BUF + ((r % 3) * 3 + (j % 3)) * S + TB::slot((r / 3) * 3 + j / 3)] = v;
    }
    for (int idx = tid; idx < (K + 1) * D; idx += WARPS * 32) {
        const int k = idx / D, r = idx - k * D;
        sRS[idx] = (r < d) ? p.RS[k * d + r] : 0.0;
    }
    __syncthreads();

    const bool lane_on = lane < 27;
    const int src = lane_on ? lane : TB::shadow(lane - 27);
    const int g = TB::perm(src) / 9;           // lane group = the chunk this lane works on
    const int li = TB::perm(src) - g * 9;      // block owned
    const int bi = li / 3, bj = li - bi * 3;
    const int r0 = bi * 3, c0 = bj * 3;
    Blk9Lane L;
    L.diag = (bi == bj);
    L.sown = TB::slot(li);
    L.syd = L.sown;
    {
        int kx1, ky1, k2;
        if (!L.diag) { kx1 = bi; ky1 = bj; k2 = 3 - bi - bj; }
        else {
            const int ko = TB::kord(src);
            kx1 = ky1 = (bi + 1 + ko) % 3;
            k2 = (bi + 2 - ko) % 3;
            L.syd = TB::slot(ky1 * 3 + bj); ky1 = bi;      // YO holds Y(k1,bi); the loaded block is the own one
        }
        L.sx1 = TB::slot(bi * 3 + kx1);
        L.sy1 = TB::slot(ky1 * 3 + bj);
        L.sx2 = TB::slot(bi * 3 + k2);
        L.sy2 = TB::slot(k2 * 3 + bj);
    }
    const int sownT = TB::slot(bj * 3 + bi);   // where this lane's block of X lands in X^dag
    const bool on_diag = L.diag;

    // buffers of this lane's group, as BASE pointers (own slot: + L.sown)
    cplx* gbase = sWarps + (size_t)warp * LY::WARP_ELEMS + LY::group_off(g);
    cplx* const bA = gbase;               // A          -> R2           -> T^dag
    cplx* const bY = gbase + BUF;         // Y (running)
    cplx* const bA2 = gbase + 2 * BUF;    // A2         -> dL1          -> dT
    cplx* const bQ = gbase + 3 * BUF;     // Q0         -> L1           -> dL2
    cplx* const bdA2 = gbase + 4 * BUF;   // dA2        -> dR1          -> T
    cplx* const bdQ = gbase + 5 * BUF;    // dQ0        -> R1'          -> dR2
    cplx* const b7 = gbase + 6 * BUF;     // L2         -> T Y
    const long long total = (long long)p.B * p.Q;
    const int CL = p.CL;

    cplx XO[3][3], YO[3][3], C[3][3], R2[3][3];
    auto load_x = [&](const cplx* buf) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) XO[a][c] = buf[(a * 3 + c) * S + L.sown];
    };
    auto load_y = [&](const cplx* buf) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) YO[a][c] = buf[(a * 3 + c) * S + L.syd];
    };
    auto store = [&](cplx* buf, const cplx (&x)[3][3]) { store_own9(buf + L.sown, 0, x, lane_on); };
    // own-slot reads are predicated off on the 5 shadow lanes: a shadow lane shares its slot address with a working lane, and
    // where that lane rewrites the slot without a warp barrier in between (state 6) the read would be a (harmless, but
    // racecheck-visible) hazard
    auto own = [&](const cplx* buf, const int a, const int c) { return lane_on ? buf[(a * 3 + c) * S + L.sown] : cmake(0.0, 0.0); };

    for (;;) {
        unsigned int unit_u = 0;
        if (lane == 0) unit_u = atomicAdd(counter, 1u);
        unit_u = __shfl_sync(0xffffffffu, unit_u, 0);
        if ((long long)unit_u * 3 >= total) break;
        const long long chunk = (long long)unit_u * 3 + g;
        const bool valid = chunk < total;
        const int b = valid ? (int)(chunk / p.Q) : 0;
        const int q = valid ? (int)(chunk - (long long)b * p.Q) : 0;
        const int n0 = q * CL;
        const int n_end = valid ? min(p.N, n0 + CL) : 0;
        const double* sig_b = p.signals + (size_t)b * K * p.N;
        double* grad_b = p.grad + (size_t)b * K * p.N;

        // Y at the head of the chunk
        {
            const cplx* Yb = p.Ybound + ((size_t)b * p.Q + q) * d * d;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    cplx v = cmake(0.0, 0.0);
                    if (valid && lane_on && r0 + a < d && c0 + c < d) v = Yb[(size_t)(r0 + a) * d + c0 + c];
                    C[a][c] = v;
                }
            __syncwarp();
            store(bY, C);
            __syncwarp();
        }

#pragma unroll 1
        for (int it = 0; it < CL; ++it) {
            const int n = n0 + it;
            const bool on = lane_on && (n < n_end);
            if (!__any_sync(0xffffffffu, on)) break;

            // ---- assemble A (own block) and the norm bound ------------------------------------------------------
            double nb = 0.0;
            {
                double nba[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    nba[a] = on ? sRS[r0 + a] : 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) XO[a][c] = on ? sG[(a * 3 + c) * S + L.sown] : cmake(0.0, 0.0);
                }
                for (int k = 0; k < K; ++k) {
                    const double cs = on ? __ldg(sig_b + (size_t)k * p.N + n) : 0.0;
                    const cplx* gk = sG + (k + 1) * BUF + L.sown;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx gv = lane_on ? gk[(a * 3 + c) * S] : cmake(0.0, 0.0);
                            XO[a][c].x = fma(cs, gv.x, XO[a][c].x);
                            XO[a][c].y = fma(cs, gv.y, XO[a][c].y);
                        }
                        nba[a] = fma(fabs(cs), sRS[(k + 1) * D + r0 + a], nba[a]);
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) nb = fmax(nb, nba[a]);
            }
            nb = __hiloint2double((int)__reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(nb) + 1u), 0);
            const int s = squarings_for(nb, C3B_THETA15);
            const double sc = pow2neg(s);
            if (s > 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) { XO[a][c].x *= sc; XO[a][c].y *= sc; }
            }
            store(bA, XO);
            __syncwarp();

            // ---- the 15 (+ 3 s) products as ONE product instance in a phase loop: straight-line code would be ~100 KB, three
            //      times the instruction cache (pwc_shfl9.cuh tells that story).  State st:
            //        0 A A | 1 A Y | 2 + Y A | 3 A2 Q0 | 4 dA2 Q0 | 5 + A2 dQ0 | 6 L1 R1' | 7 dL1 R1' | 8 + L1 dR1 |
            //        9 dE0 + dL2 R2 | 10 + L2 dR2 | 11 E0 + L2 R2 | 12 dT T | 13 + T dT | 14 T T (12-14: s times) |
            //        15 dT T^dag | 16 T Y | 17 (T Y) T^dag
            constexpr double kI = 1.0 / (C3B_T15_B1 - C3B_T15_B3);
            const cplx* xb = bA;
            const cplx* yb = bA;
            bool acc = false;
            int st = 0, sq = 0;
#pragma unroll 1
            for (;;) {
                load_x(xb);
                load_y(yb);
                if (!acc) {
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) C[a][c] = cmake(0.0, 0.0);
                }
                mm_own9<true, true>(xb, yb, L, XO, YO, C);
                if (st == 0) {                                  // C = A2;  XO = own block of A
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            R2[a][c] = cmake(C3B_T15_A1 * C[a][c].x + C3B_T15_A2 * XO[a][c].x, C3B_T15_A1 * C[a][c].y + C3B_T15_A2 * XO[a][c].y);
                    store(bA2, C);
                    store(bQ, R2);
                    xb = bA; yb = bY; acc = false; st = 1;
                } else if (st == 1) {
                    xb = bY; yb = bA; acc = true; st = 2;
                } else if (st == 2) {                           // C = dA2;  XO = own block of Y:  dQ0 = a1 dA2 + a2 Y
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            R2[a][c] = cmake(C3B_T15_A1 * C[a][c].x + C3B_T15_A2 * XO[a][c].x, C3B_T15_A1 * C[a][c].y + C3B_T15_A2 * XO[a][c].y);
                    store(bdA2, C);
                    store(bdQ, R2);
                    __syncwarp();
                    xb = bA2; yb = bQ; acc = false; st = 3;
                } else if (st == 3) {                           // C = P0 (own block kept in R2)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) R2[a][c] = C[a][c];
                    xb = bdA2; yb = bQ; acc = false; st = 4;
                } else if (st == 4) {
                    xb = bA2; yb = bdQ; acc = true; st = 5;
                } else if (st == 5) {                           // C = dP0;  XO = own block of A2
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx x2 = XO[a][c], p0 = R2[a][c], dp0 = C[a][c];
                            const cplx x1 = own(bA, a, c), dx2 = own(bdA2, a, c), y = own(bY, a, c);
                            YO[a][c] = cmake(p0.x + C3B_T15_B1 * x2.x + C3B_T15_B2 * x1.x, p0.y + C3B_T15_B1 * x2.y + C3B_T15_B2 * x1.y);      // L1
                            R2[a][c] = cmake(p0.x + C3B_T15_B3 * x2.x, p0.y + C3B_T15_B3 * x2.y);                                              // R1'
                            XO[a][c] = cmake(dp0.x + C3B_T15_B1 * dx2.x + C3B_T15_B2 * y.x, dp0.y + C3B_T15_B1 * dx2.y + C3B_T15_B2 * y.y);    // dL1
                            C[a][c] = cmake(dp0.x + C3B_T15_B3 * dx2.x, dp0.y + C3B_T15_B3 * dx2.y);                                            // dR1
                        }
                    __syncwarp();                          // every lane has read A2, Q0, dA2, dQ0
                    store(bQ, YO);
                    store(bdQ, R2);
                    store(bA2, XO);
                    store(bdA2, C);
                    __syncwarp();
                    xb = bQ; yb = bdQ; acc = false; st = 6;
                } else if (st == 6) {                           // C = L1 R1';  XO = own block of L1
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx l1 = XO[a][c], r1 = own(bdQ, a, c), x1 = own(bA, a, c);
                            const cplx x2 = cmake((l1.x - r1.x - C3B_T15_B2 * x1.x) * kI, (l1.y - r1.y - C3B_T15_B2 * x1.y) * kI);
                            const cplx p0 = cmake(r1.x - C3B_T15_B3 * x2.x, r1.y - C3B_T15_B3 * x2.y);
                            const cplx p1 = cmake(C[a][c].x + C3B_T15_B4 * l1.x + C3B_T15_B5 * p0.x, C[a][c].y + C3B_T15_B4 * l1.y + C3B_T15_B5 * p0.y);
                            const double dg = (on_diag && a == c) ? 1.0 : 0.0;
                            XO[a][c] = cmake(p1.x + C3B_T15_C1 * x2.x + C3B_T15_C2 * x1.x, p1.y + C3B_T15_C1 * x2.y + C3B_T15_C2 * x1.y);   // L2
                            YO[a][c] = cmake(p1.x + C3B_T15_C3 * p0.x + C3B_T15_C4 * x1.x, p1.y + C3B_T15_C3 * p0.y + C3B_T15_C4 * x1.y);   // R2
                            R2[a][c] = cmake(C3B_T15_C9 * p1.x + C3B_T15_C5 * p0.x + C3B_T15_C6 * x2.x + C3B_T15_C7 * x1.x + C3B_T15_C8 * dg,
                                             C3B_T15_C9 * p1.y + C3B_T15_C5 * p0.y + C3B_T15_C6 * x2.y + C3B_T15_C7 * x1.y);                 // E0
                        }
                    store(b7, XO);                         // b7 is free; bA is read through the own slot only from here on
                    store(bA, YO);
                    xb = bA2; yb = bdQ; acc = false; st = 7;
                } else if (st == 7) {
                    xb = bQ; yb = bdA2; acc = true; st = 8;
                } else if (st == 8) {                           // C = dL1 R1' + L1 dR1
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx dl1 = own(bA2, a, c), dr1 = own(bdA2, a, c), y = own(bY, a, c);
                            const cplx dx2 = cmake((dl1.x - dr1.x - C3B_T15_B2 * y.x) * kI, (dl1.y - dr1.y - C3B_T15_B2 * y.y) * kI);
                            const cplx dp0 = cmake(dr1.x - C3B_T15_B3 * dx2.x, dr1.y - C3B_T15_B3 * dx2.y);
                            const cplx dp1 = cmake(C[a][c].x + C3B_T15_B4 * dl1.x + C3B_T15_B5 * dp0.x, C[a][c].y + C3B_T15_B4 * dl1.y + C3B_T15_B5 * dp0.y);
                            XO[a][c] = cmake(dp1.x + C3B_T15_C1 * dx2.x + C3B_T15_C2 * y.x, dp1.y + C3B_T15_C1 * dx2.y + C3B_T15_C2 * y.y);  // dL2
                            YO[a][c] = cmake(dp1.x + C3B_T15_C3 * dp0.x + C3B_T15_C4 * y.x, dp1.y + C3B_T15_C3 * dp0.y + C3B_T15_C4 * y.y);  // dR2
                            C[a][c] = cmake(C3B_T15_C9 * dp1.x + C3B_T15_C5 * dp0.x + C3B_T15_C6 * dx2.x + C3B_T15_C7 * y.x,
                                            C3B_T15_C9 * dp1.y + C3B_T15_C5 * dp0.y + C3B_T15_C6 * dx2.y + C3B_T15_C7 * y.y);                 // dE0
                        }
                    __syncwarp();                          // every lane has read L1, R1', dL1, dR1
                    store(bQ, XO);
                    store(bdQ, YO);
                    __syncwarp();
                    xb = bQ; yb = bA; acc = true; st = 9;       // C = dE0 + dL2 R2 + ...
                } else if (st == 9) {
                    xb = b7; yb = bdQ; acc = true; st = 10;
                } else if (st == 10) {                          // C = dT of the scaled slice
                    if (s > 0) {                                // its direction was Y / 2^s
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c) { C[a][c].x *= sc; C[a][c].y *= sc; }
                    }
                    store(bA2, C);                              // dT (dL1 is dead)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) C[a][c] = R2[a][c];            // E0 + L2 R2
                    xb = b7; yb = bA; acc = true; st = 11;
                } else if (st == 11 || st == 14) {              // C = T (st 14: squared)
                    if (st == 14) {
                        __syncwarp();                           // dT, T fully read
                        store(bA2, R2);
                    }
                    store(bdA2, C);
                    __syncwarp();
                    if (sq < s) {
                        ++sq;
                        xb = bA2; yb = bdA2; acc = false; st = 12;
                    } else {
                        // T^dag: this lane's block of T, conjugate-transposed, is block (bj, bi)  (bA: R2 is dead)
                        if (lane_on) {
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int c = 0; c < 3; ++c) bA[(c * 3 + a) * S + sownT] = cmake(C[a][c].x, -C[a][c].y);
                        }
                        __syncwarp();
                        xb = bA2; yb = bA; acc = false; st = 15;
                    }
                } else if (st == 12) {
                    xb = bdA2; yb = bA2; acc = true; st = 13;
                } else if (st == 13) {                          // C = dT T + T dT (kept in R2 until T T is done)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) R2[a][c] = C[a][c];
                    xb = bdA2; yb = bdA2; acc = false; st = 14;
                } else if (st == 15) {                          // C = V = dT T^dag: the K contractions
                    for (int k = 0; k < K; ++k) {
                        // Re tr(V (G_k + t_k I)): this lane's block (bi,bj) of V meets block (bj,bi) of G_k, transposed
                        const cplx* gk = sG + (k + 1) * BUF + sownT;
                        double part = 0.0;
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const cplx gv = gk[(c * 3 + a) * S];
                                part += C[a][c].x * gv.x - C[a][c].y * gv.y;
                            }
                        if (on_diag) {
                            const cplx t = p.TR[k + 1];
#pragma unroll
                            for (int a = 0; a < 3; ++a) part += C[a][a].x * t.x - C[a][a].y * t.y;
                        }
                        double tot = 0.0;
#pragma unroll
                        for (int gg = 0; gg < 3; ++gg) {
                            const double v = warp_sum((on && g == gg) ? part : 0.0);
                            if (g == gg) tot = v;
                        }
                        if (on && li == 0) grad_b[(size_t)k * p.N + n] = tot;
                    }
                    xb = bdA2; yb = bY; acc = false; st = 16;
                } else if (st == 16) {                          // C = T Y (L2 is dead)
                    store(b7, C);
                    __syncwarp();
                    xb = b7; yb = bA; acc = false; st = 17;
                } else {                                        // C = T Y T^dag: the next slice's Y
                    __syncwarp();
                    store(bY, C);
                    __syncwarp();
                    break;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace c3b
