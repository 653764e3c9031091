// d = 9 (and zero-padded d = 7) fused PWC propagator kernel, second generation of the 3x3-lane block layout
// (pwc_blk.cuh): "own-block" products in a block-interleaved shared layout.
//
// Same contract as pwc_blk_taylor_kernel (assemble -> trace-shifted degree-15+ Taylor exponential in 4 products ->
// ordered product; replaces c3/libraries/propagation.py:426-440,460-515 and c3/utils/tf_utils.py:120-193).
//
// What changed.  pwc_blk_taylor_kernel streams BOTH operands of every 9x9 product from shared memory: per lane
// 54 LDS.128 for 81 complex MACs, and the shared-memory return path (128 lane-bytes per clock per SM) is what
// bounds it (82 % busy at 53 % fp64 pipe).  Lane (bi,bj) of a 3x3 lane group owns block (bi,bj) of every matrix it
// produces, so two of the six operand blocks of  C(bi,bj) = sum_k X(bi,k) Y(k,bj)  are already in its registers:
//   off-diagonal lane:  X(bi,bj) [own] * Y(bj,bj),   X(bi,bi) * Y(bi,bj) [own],   X(bi,k2) * Y(k2,bj)
//   diagonal lane:      X(bi,bi) [own] * Y(bi,bi) [own],   X(bi,k1) * Y(k1,bi),   X(bi,k2) * Y(k2,bi)
// -> 36 LDS.128 per product for every lane, in ONE instruction stream: the loaded Y block and the own Y block swap
// roles on diagonal lanes (a register select).
//
// Layout.  The row-major layout of pwc_blk.cuh cannot serve these loads without bank conflicts (the 8 lanes of a
// quarter-warp now read 8 different blocks).  Here a matrix buffer is element-major: element e = 3a+c of block blk
// lives at  e * S + slot[blk]  (16-byte units), so all lanes of one LDS/STS read the SAME element of different
// blocks and a conflict-free instruction only needs (slot[blk] + group offset) mod 8 distinct inside each
// quarter-warp.  Slots, group offsets, the lane -> block permutation of every group, the shadow targets of the 5
// spare lanes and the k order of the diagonal lanes were found by simulated annealing
// (scratch/conflict_blockmajor.py): 4.0 wavefronts per load, 5 per store / generator load.
#pragma once
#include "c3b_common.cuh"
#include "c3b_params.cuh"

namespace c3b {

template <bool NOSEL>
struct Blk9T {
    static constexpr int S = 9;                    // element stride: slots 0..8
    static constexpr int BUF = 9 * S;              // one 9x9 matrix
    static constexpr int NBUF = 4;                 // A, B, X, P per lane group
    // group bases (mod 8) as searched with the tables below; warp stride = 0 (mod 8)
    static constexpr int G1 = NOSEL ? 329 : 330;   // = 409 / 410 (mod 8)
    static constexpr int G2 = NOSEL ? 653 : 654;   // = 821 / 822 (mod 8)
    static constexpr int WARP_ELEMS = 984;
    static_assert(G1 >= NBUF * BUF && G2 >= G1 + NBUF * BUF && WARP_ELEMS >= G2 + NBUF * BUF && WARP_ELEMS % 8 == 0, "layout");
    __host__ __device__ static constexpr int group_off(int g) { return g == 0 ? 0 : (g == 1 ? G1 : G2); }
    __host__ __device__ static size_t smem_bytes(int K, int warps) {
        size_t model = (size_t)(K + 1) * BUF * sizeof(cplx) + (size_t)(((K + 1) * 9 + 1) & ~1) * sizeof(double);
        return model + (size_t)warps * WARP_ELEMS * sizeof(cplx);
    }
};
struct Blk9 { static constexpr int S = 9; static constexpr int BUF = 81; };

// block (bi*3+bj) -> slot inside an element row (slots 0 and 8 share a bank)
__constant__ signed char kB9Slot[9] = {4, 0, 2, 1, 8, 6, 3, 7, 5};
// lane (0..26) -> 9 * group + block owned
__constant__ signed char kB9Perm[27] = {5, 1, 6, 2, 8, 0, 3, 7, 4,   11, 15, 12, 9, 10, 16, 17, 14, 13,   26, 24, 22, 21, 23, 25, 19, 18, 20};
// lanes 27..31 shadow these lanes (same addresses, never store)
__constant__ signed char kB9Shadow[5] = {12, 9, 14, 13, 12};
// diagonal lanes: 0 -> (k1, k2) = (bi+1, bi+2), 1 -> (bi+2, bi+1)
__constant__ signed char kB9Kord[27] = {0, 0, 0, 0, 0, 1, 0, 0, 1,   1, 1, 0, 1, 1, 1, 0, 1, 0,   1, 0, 0, 0, 1, 0, 1, 1, 1};

// second table set: layout searched for the select-free product (NOSEL), see mm_own9
__constant__ signed char kB9nSlot[9] = {2, 6, 1, 5, 4, 0, 7, 3, 8};
__constant__ signed char kB9nPerm[27] = {6, 9, 16, 11, 12, 13, 15, 17, 1,   4, 10, 0, 3, 2, 7, 8, 22, 26,   20, 21, 19, 18, 25, 24, 23, 14, 5};
__constant__ signed char kB9nShadow[5] = {0, 8, 25, 8, 25};
__constant__ signed char kB9nKord[27] = {0, 0, 0, 0, 0, 1, 1, 1, 0,   1, 0, 1, 1, 0, 0, 0, 1, 0,   0, 0, 0, 0, 0, 1, 0, 0, 1};
template <bool NOSEL> struct Blk9Tab {
    __device__ static __forceinline__ int slot(int i) { return NOSEL ? kB9nSlot[i] : kB9Slot[i]; }
    __device__ static __forceinline__ int perm(int i) { return NOSEL ? kB9nPerm[i] : kB9Perm[i]; }
    __device__ static __forceinline__ int shadow(int i) { return NOSEL ? kB9nShadow[i] : kB9Shadow[i]; }
    __device__ static __forceinline__ int kord(int i) { return NOSEL ? kB9nKord[i] : kB9Kord[i]; }
};

struct Blk9Lane {
    int sown, sx1, sy1, sx2, sy2;   // slots of the own block and of the four loaded operand blocks
    int syd;                        // NOSEL: slot of Y(k1,bi), the block a diagonal lane keeps in YO
    bool diag;
};

// C = X * Y for this lane's block.  XO / YO: own blocks of X and Y (registers); Xb / Yb: the matrices in shared memory.
template <bool NOSEL, bool ACC = false>
__device__ __forceinline__ void mm_own9(const cplx* __restrict__ Xb, const cplx* __restrict__ Yb, const Blk9Lane& L,
                                        const cplx (&XO)[3][3], const cplx (&YO)[3][3], cplx (&c)[3][3]) {
    constexpr int S = Blk9::S;
    const cplx* x1 = Xb + L.sx1;
    const cplx* y1 = Yb + L.sy1;
    const cplx* x2 = Xb + L.sx2;
    const cplx* y2 = Yb + L.sy2;
    if constexpr (!ACC) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) c[a][b] = cmake(0.0, 0.0);
    }
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        cplx lx[3], ly[3], ya[3], yb[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) ly[b] = y1[(kk * 3 + b) * S];
#pragma unroll
        for (int a = 0; a < 3; ++a) lx[a] = x1[(a * 3 + kk) * S];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if constexpr (NOSEL) { ya[b] = ly[b]; yb[b] = YO[kk][b]; }
            else {
                ya[b].x = L.diag ? YO[kk][b].x : ly[b].x;
                ya[b].y = L.diag ? YO[kk][b].y : ly[b].y;
                yb[b].x = L.diag ? ly[b].x : YO[kk][b].x;
                yb[b].y = L.diag ? ly[b].y : YO[kk][b].y;
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                cfma(c[a][b], XO[a][kk], ya[b]);
                cfma(c[a][b], lx[a], yb[b]);
            }
    }
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
        cplx lx[3], ly[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) ly[b] = y2[(kk * 3 + b) * S];
#pragma unroll
        for (int a = 0; a < 3; ++a) lx[a] = x2[(a * 3 + kk) * S];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) cfma(c[a][b], lx[a], ly[b]);
    }
}

// generic product with all six operand blocks from shared memory (fold of the group products: rare)
template <bool NOSEL>
__device__ __forceinline__ void mm_full9(const cplx* __restrict__ Xb, const cplx* __restrict__ Yb, const int bi, const int bj,
                                         cplx (&c)[3][3]) {
    constexpr int S = Blk9::S;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) c[a][b] = cmake(0.0, 0.0);
#pragma unroll 1
    for (int kb = 0; kb < 3; ++kb) {
        const cplx* x = Xb + Blk9Tab<NOSEL>::slot(bi * 3 + kb);
        const cplx* y = Yb + Blk9Tab<NOSEL>::slot(kb * 3 + bj);
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) cfma(c[a][b], x[(a * 3 + kk) * S], y[(kk * 3 + b) * S]);
    }
}

__device__ __forceinline__ void store_own9(cplx* __restrict__ M, const int sown, const cplx (&x)[3][3], const bool pred) {
    if (pred) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) M[(a * 3 + b) * Blk9::S + sown] = x[a][b];
    }
}

template <int WARPS, int MINB, bool NOSEL, int GATED>
__global__ void __launch_bounds__(WARPS * 32, MINB) pwc_blk9_taylor_kernel(const RowsParams p, unsigned int* __restrict__ counter) {
    using LY = Blk9T<NOSEL>;
    using TB = Blk9Tab<NOSEL>;
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
    const bool hmode = p.hlist != nullptr;

    if (!hmode) {
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
    }
    __syncthreads();

    const bool lane_on = lane < 27;
    const int src = lane_on ? lane : TB::shadow(lane - 27);
    const int g = TB::perm(src) / 9;           // lane group = the slice chunk this lane works on
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
            if (NOSEL) { L.syd = TB::slot(ky1 * 3 + bj); ky1 = bi; }   // YO holds Y(k1,bi); the loaded block is the own one
        }
        L.sx1 = TB::slot(bi * 3 + kx1);
        L.sy1 = TB::slot(ky1 * 3 + bj);
        L.sx2 = TB::slot(bi * 3 + k2);
        L.sy2 = TB::slot(k2 * 3 + bj);
    }
    const bool on_diag = L.diag;
    int rowlane[3] = {0, 0, 0};       // H-list mode: the lanes holding this lane's block row (row sums by shuffle)
    if (hmode) {
#pragma unroll
        for (int q = 0; q < 3; ++q)
            for (int j = 0; j < 27; ++j)
                if (TB::perm(j) == g * 9 + bi * 3 + q) rowlane[q] = j;
    }

    cplx* gbase = sWarps + (size_t)warp * LY::WARP_ELEMS + LY::group_off(g);
    cplx* bufA = gbase + L.sown;              // A, then (own block only) B3          -- all five pre-offset to the own slot
    cplx* bufB = gbase + BUF + L.sown;        // A^2, then B1, then the left operands of the later products
    cplx* bufX = gbase + 2 * BUF + L.sown;    // A^3, then B5, then A9
    cplx* bufP = gbase + 3 * BUF + L.sown;    // running product of this group's slices
    const int unown = -L.sown;                // back to the buffer base for the operand loads
    const cplx hs = cmake(p.hscale_re, p.hscale_im);
    const long long total_units = (long long)p.B * p.S;
    const bool shifted = (p.TR != nullptr) && !hmode;

    for (;;) {
        unsigned int unit_u = 0;
        if (lane == 0) {
            unit_u = atomicAdd(counter, 1u);
            if constexpr (GATED != 0) {
                // gated launch: lane 0 waits until this unit's batch row has landed (rows arrive in order) -- all the polling
                // state lives and dies here, nothing stays in registers across the slice loop (a per-warp cached counter or
                // an all-lanes wait cost 5 registers = spills in this 252-register kernel: 10.18 instead of 9.7 ms).  A row
                // that never arrives ends this warp's work like an exhausted counter and raises gate[1]; the caller pre-fills
                // U with NaN, so the failure is loud and the GPU does not hang.
                if ((long long)unit_u < total_units) {
                    unsigned int known = 0;
                    if (!wait_rows_ready(p.gate, (int)(unit_u / (unsigned int)p.S), known)) unit_u = 0xffffffffu;
                }
            }
        }
        unit_u = __shfl_sync(0xffffffffu, unit_u, 0);
        if constexpr (GATED != 0) __syncwarp();   // lane 0's acquire, then the warp barrier: every lane's loads of the row come after
        const long long unit = unit_u;
        if (unit >= total_units) break;
        const int b = (int)(unit / p.S);
        const int sidx = (int)(unit - (long long)b * p.S);
        const int n_begin = sidx * p.seg_len;
        const int n_end = min(p.N, n_begin + p.seg_len);
        const int len = n_end - n_begin;
        const int cl = (len + 2) / 3;
        const int my_begin = n_begin + g * cl;
        const int my_end = min(n_end, my_begin + cl);
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        cplx mu_acc = cmake(0.0, 0.0);

#pragma unroll 1
        for (int it = 0; it < cl; ++it) {
            const int n = my_begin + it;
            const bool on = lane_on && (n < my_end);

            cplx XO[3][3], YO[3][3], C[3][3], R2[3][3];     // own blocks: X, Y operands, product, A^2 then E0
            cplx mu = cmake(0.0, 0.0);
            double nb = 0.0;
            if (!hmode) {
                double nba[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    nba[a] = on ? sRS[r0 + a] : 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) XO[a][c] = on ? sG[(a * 3 + c) * S + L.sown] : cmake(0.0, 0.0);
                }
                if (shifted && on) mu = p.TR[0];
                for (int k = 0; k < K; ++k) {
                    const double cs = on ? load_signal<GATED>(sig_b + (size_t)k * p.N + n) : 0.0;
                    const cplx* gk = sG + (k + 1) * BUF + L.sown;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx gv = lane_on ? gk[(a * 3 + c) * S] : cmake(0.0, 0.0);   // shadow lanes off: 3 wavefronts, not 4
                            XO[a][c].x = fma(cs, gv.x, XO[a][c].x);
                            XO[a][c].y = fma(cs, gv.y, XO[a][c].y);
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
                        XO[a][c] = cmul(hs, h);
                        rs += cabs1(XO[a][c]);
                    }
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < 3; ++q) tot += __shfl_sync(0xffffffffu, rs, rowlane[q]);
                    nb = fmax(nb, tot);
                }
            }
            // warp-wide upper bound of the norm estimates in ONE redux.sync: nb >= 0, so the high words order like the
            // doubles; rounding the high word up keeps it an upper bound (relative slack 2^-20)
            nb = __hiloint2double((int)__reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(nb) + 1u), 0);
            mu_acc.x += mu.x; mu_acc.y += mu.y;

            const int s = squarings_for(nb, C3B_THETA15);
            if (s > 0) {
                const double sc = pow2neg(s);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) { XO[a][c].x *= sc; XO[a][c].y *= sc; }
            }
            store_own9(bufA, 0, XO, lane_on);
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) YO[a][c] = XO[a][c];
            // phases (degree-15+ scheme in 4 products, c3b_common.cuh):
            //   0: A2 = A A | 1: P0 = A2 (a1 A2 + a2 A) | 2: P1 = L1 R1 + b5 P0 | 3: T = L2 R2 + E0 | s squarings | product
            // own blocks kept in registers: XO / YO (operands), C (product), R2 (A^2, then E0); A stays in bufA
            const int ph_lastsq = 3 + s;
            const int ph_last = ph_lastsq + (it > 0 ? 1 : 0);
            const cplx* Xb = bufA + unown;
            const cplx* Yb = bufA + unown;

#pragma unroll 1
            for (int ph = 0; ph <= ph_last; ++ph) {
                if constexpr (NOSEL) {
                    if (L.diag) {      // diagonal lanes pair their loaded X(bi,k1) with Y(k1,bi): fetch it into YO
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c) YO[a][c] = Yb[(a * 3 + c) * S + L.syd];
                    }
                }
                mm_own9<NOSEL>(Xb, Yb, L, XO, YO, C);
                if (ph == 0) {                                  // C = A^2: operands of P0 = A2 (a1 A2 + a2 A)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx a1v = XO[a][c];                                        // own block of A: still the X operand
                            const cplx q0 = cmake(C3B_T15_A1 * C[a][c].x + C3B_T15_A2 * a1v.x, C3B_T15_A1 * C[a][c].y + C3B_T15_A2 * a1v.y);
                            R2[a][c] = C[a][c];
                            XO[a][c] = C[a][c];
                            YO[a][c] = q0;
                            if (lane_on) {
                                bufB[(a * 3 + c) * S] = C[a][c];  // left operand A^2
                                bufX[(a * 3 + c) * S] = q0;       // right operand
                            }
                        }
                    __syncwarp();
                    Xb = bufB + unown; Yb = bufX + unown;
                } else if (ph == 1) {                           // C = P0: L1 = P0 + b1 A2 + b2 A, R1' = P0 + b3 A2  (R1 = R1' + b4 I)
                    cplx A1[3][3];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) A1[a][c] = bufA[(a * 3 + c) * S];          // loads first
                    __syncwarp();                               // bufB / bufX fully read
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const cplx p0 = C[a][c], x2 = R2[a][c], x1 = A1[a][c];
                            const cplx l1 = cmake(p0.x + C3B_T15_B1 * x2.x + C3B_T15_B2 * x1.x, p0.y + C3B_T15_B1 * x2.y + C3B_T15_B2 * x1.y);
                            const cplx r1 = cmake(p0.x + C3B_T15_B3 * x2.x, p0.y + C3B_T15_B3 * x2.y);
                            XO[a][c] = l1;
                            YO[a][c] = r1;
                            if (lane_on) {
                                bufB[(a * 3 + c) * S] = l1;
                                bufX[(a * 3 + c) * S] = r1;
                            }
                        }
                    __syncwarp();
                } else if (ph == 2) {
                    // C = L1 R1'.  Nothing was parked: L1 is still the X operand, R1' is re-read (a diagonal lane's YO holds
                    // another block), and P0 = R1' - b3 A2, A = (L1 - P0 - b1 A2) / b2 are recovered from them (no cancellation:
                    // |b3 A2| ~ |P0| / 25 ... and L1 is dominated by b2 A).  P1 = C + b4 L1 + b5 P0; then L2, R2 and the epilogue
                    // E0, which takes over the registers of A2.
                    cplx R1[3][3];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) R1[a][c] = bufX[(a * 3 + c) * S];
                    __syncwarp();                               // bufB / bufX fully read
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            constexpr double kIB2 = 1.0 / C3B_T15_B2;
                            const cplx x2 = R2[a][c], l1 = XO[a][c];
                            const cplx p0 = cmake(R1[a][c].x - C3B_T15_B3 * x2.x, R1[a][c].y - C3B_T15_B3 * x2.y);
                            const cplx x1 = cmake((l1.x - p0.x - C3B_T15_B1 * x2.x) * kIB2, (l1.y - p0.y - C3B_T15_B1 * x2.y) * kIB2);
                            const cplx p1 = cmake(C[a][c].x + C3B_T15_B4 * l1.x + C3B_T15_B5 * p0.x, C[a][c].y + C3B_T15_B4 * l1.y + C3B_T15_B5 * p0.y);
                            const double dg = (on_diag && a == c) ? 1.0 : 0.0;
                            const cplx l2 = cmake(p1.x + C3B_T15_C1 * x2.x + C3B_T15_C2 * x1.x, p1.y + C3B_T15_C1 * x2.y + C3B_T15_C2 * x1.y);
                            const cplx r2 = cmake(p1.x + C3B_T15_C3 * p0.x + C3B_T15_C4 * x1.x, p1.y + C3B_T15_C3 * p0.y + C3B_T15_C4 * x1.y);
                            R2[a][c] = cmake(C3B_T15_C9 * p1.x + C3B_T15_C5 * p0.x + C3B_T15_C6 * x2.x + C3B_T15_C7 * x1.x + C3B_T15_C8 * dg,
                                             C3B_T15_C9 * p1.y + C3B_T15_C5 * p0.y + C3B_T15_C6 * x2.y + C3B_T15_C7 * x1.y);
                            XO[a][c] = l2;
                            YO[a][c] = r2;
                            if (lane_on) {
                                bufB[(a * 3 + c) * S] = l2;
                                bufX[(a * 3 + c) * S] = r2;
                            }
                        }
                    __syncwarp();
                } else {
                    if (ph == 3) {                              // C = L2 R2: T = C + E0
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c) { C[a][c].x += R2[a][c].x; C[a][c].y += R2[a][c].y; }
                    }
                    if (ph <= ph_lastsq) {
                        // C = exp(A_n / 2^s)^(2^(ph-3)); publish as the next left operand
                        __syncwarp();
                        store_own9(bufB, 0, C, lane_on);
                        Xb = bufB + unown;
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int c = 0; c < 3; ++c) XO[a][c] = C[a][c];
                        if (ph < ph_lastsq) {
                            Yb = bufB + unown;
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int c = 0; c < 3; ++c) YO[a][c] = C[a][c];
                        } else {
                            Yb = bufP + unown;
                            if (it == 0) store_own9(bufP, 0, C, lane_on);
                            else {
#pragma unroll
                                for (int a = 0; a < 3; ++a)
#pragma unroll
                                    for (int c = 0; c < 3; ++c) YO[a][c] = bufP[(a * 3 + c) * S];
                            }
                            if (GATED == 0 && p.dUs_out != nullptr && on) {
                                const cplx ph_n = shifted ? cexp_(mu) : cmake(1.0, 0.0);
#pragma unroll
                                for (int a = 0; a < 3; ++a) {
                                    const int row = r0 + a;
                                    if (row < d) {
                                        cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)row * d + c0;
#pragma unroll
                                        for (int c = 0; c < 3; ++c)
                                            if (c0 + c < d) o[c] = shifted ? cmul(ph_n, C[a][c]) : C[a][c];
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    } else {                                    // C = dU_n * P
                        __syncwarp();
                        store_own9(bufP, 0, C, lane_on);
                    }
                }
            }
        }
        __syncwarp();

        // ---- re-apply this group's accumulated shift: P_g <- exp(sum mu) P_g -------------------------
        if (shifted) {
            const cplx ph_g = cexp_(mu_acc);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const cplx v = bufP[(a * 3 + c) * S];
                    if (lane_on) bufP[(a * 3 + c) * S] = cmul(ph_g, v);
                }
            __syncwarp();
        }

        // ---- fold the group products: P_2 P_1 P_0 (every group computes it; group 0 writes) ----------
        cplx* wbase = sWarps + (size_t)warp * LY::WARP_ELEMS;
        const cplx* cur = wbase + LY::group_off(2) + 3 * BUF;
        int flip = 0;
#pragma unroll 1
        for (int gg = 1; gg >= 0; --gg) {
            cplx T[3][3];
            mm_full9<NOSEL>(cur, wbase + LY::group_off(gg) + 3 * BUF, bi, bj, T);
            cplx* dst = flip ? bufA : bufX;
            store_own9(dst, 0, T, lane_on);
            __syncwarp();
            cur = dst + unown;
            flip ^= 1;
        }
        if (lane_on && g == 0) {
            cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int row = r0 + a;
                if (row < d) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (c0 + c < d) o[row * d + c0 + c] = cur[(a * 3 + c) * S + L.sown];
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace c3b
