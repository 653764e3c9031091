// CTA-cooperative PWC propagator kernel for larger dimensions (closed d > 12, Lindblad D = d^2 up to
// 81 and beyond), built on fp64 tensor-core tiles.
//
// Per slice: A_n (trace-shifted generators, scaled by 2^-s from the row-sum bound) -> degree-15+ Taylor polynomial in FOUR
// complex products (c3b_common.cuh), every linear combination formed in the epilogue of the product before it -> s squarings
// -> running ordered product.  Every O(D^3) step is ONE routine, cta_zgemm: C = A B on zero-padded DP x DP complex
// matrices (DP = D rounded up to 8), each warp owning macro tiles of m8n8 blocks and issuing mma.sync.m8n8k4.f64 (DMMA):
// a complex tile step is 3 real DMMAs (the 3M product: Ar Br, Ai Bi, (Ar + Ai)(Br + Bi)) on fragments loaded straight from
// the interleaved complex storage (one 16-byte load = re and im of a fragment element).  One operand load feeds 8 complex
// MACs per lane (vs 1-1.5 in the register kernels), so the D^2 x D^2 Lindblad superoperator is a genuinely dense
// contraction on the fp64 pipe, as BASELINE.json's north_star asks.  Matrices live in shared memory (DP <= 32, XOR
// swizzled) or in a per-CTA global workspace that stays in L1/L2 (D = 81: 6 x 124 KB).  Replaces
// c3/libraries/propagation.py:426-440,551-585 and c3/utils/tf_utils.py:120-193 for these shapes; no linear solve, no pivoting.
#pragma once
#include "c3b_common.cuh"
#include "pwc_cta.cuh"   // CtaParams, kCtaThreads

namespace c3b {


// XOR swizzle of the shared-memory matrices of the DP = 32 kernels (leading dimension exactly 32, 16-byte columns):
// element (i, j) lives at i * 32 + (j ^ swz_of_row(i)).  With f(i) = 4 (i & 1) | 2 ((i >> 1) & 1) the 8 lanes of a quarter-warp
// hit 8 distinct 16-byte bank slots for BOTH fragment patterns of mma.m8n8k4 (A: 2 rows x 4 consecutive k, B: 4 consecutive k
// x 2 columns); the padded layout LD = 36 left the B-fragment loads 2-way conflicted (ncu: 32 % of all shared wavefronts,
// the MIO pipe at 80 % with the tensor pipe at 49 %, profiles/r02_prof_d27_3m.txt).
__device__ __forceinline__ int swz_of_row(const int i) { return ((i & 1) << 2) | (i & 2); }
template <bool SWZ>
__device__ __forceinline__ int mat_idx(const int i, const int j, const int LD) {
    if constexpr (SWZ) return i * 32 + (j ^ swz_of_row(i));
    else return i * LD + j;
}

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Fused epilogues of cta_zgemm: what happens to the two adjacent results c0 = (A B)[row, col], c1 = (A B)[row, col + 1] a lane
// holds (idx = their storage index; every element is read and written by the lane that owns it, so an output may alias
// any epilogue INPUT -- never an operand of the running product).
//   0: C = A B                          1: C = A B + E1, E2 += C                    2: C = A B + E1
// degree-15+ scheme (c3b_common.cuh), so that no separate element-wise pass forms the combinations:
//   3: C = A B (= A^2),  E2 = a1 C + a2 E1                                        (E1 = A; E2 receives a1 A^2 + a2 A)
//   4: c = A B (= P0);  C = c + b3 E1 (= R1 - b4 I),  E2 = c + b1 E1 + b2 E3 (= L1)          (E1 = A^2, E3 = A)
//   5: p = A B (= L1 (R1 - b4 I));  with P0 = E1 - b3 E3 recovered (E1 = R1 - b4 I is an operand: read only) and
//      P1 = p + b4 L1 + b5 P0 = p + (b4 + b5) P0 + b4 b1 A^2 + b4 b2 A:
//      E2 <- P1 + c1 E3 + c2 E2 (= L2),  E3 <- P1 + c3 P0 + c4 E2 (= R2),  C <- c9 P1 + c5 P0 + c6 E3 + c7 E2 + c8 I (= E0)
//      (on entry E2 = A, E3 = A^2).  The identity part of R1 is applied as b4 L1 so that P0 can be recovered from the
//      stored operand without cancellation against a diagonal of 3 (|b3 A^2| <= 0.05: the recovery is exact to 1e-17).
//      c8 I also lands on the padding diagonal: the padding block of T is then the identity, which products keep
//      block-diagonal and no output reads.
template <int EPI>
__device__ __forceinline__ void zgemm_epilogue(cplx c0, cplx c1, const int idx, const int row, const int col, cplx* C, const cplx* E1,
                                               cplx* E2, cplx* E3) {
    if constexpr (EPI == 1 || EPI == 2) {
        const cplx e0 = E1[idx], e1 = E1[idx + 1];
        c0.x += e0.x; c0.y += e0.y; c1.x += e1.x; c1.y += e1.y;
    }
    if constexpr (EPI == 1) {
        const cplx f0 = E2[idx], f1 = E2[idx + 1];
        E2[idx] = cmake(f0.x + c0.x, f0.y + c0.y);
        E2[idx + 1] = cmake(f1.x + c1.x, f1.y + c1.y);
    }
    if constexpr (EPI == 3) {
        const cplx a0 = E1[idx], a1 = E1[idx + 1];
        E2[idx] = cmake(C3B_T15_A1 * c0.x + C3B_T15_A2 * a0.x, C3B_T15_A1 * c0.y + C3B_T15_A2 * a0.y);
        E2[idx + 1] = cmake(C3B_T15_A1 * c1.x + C3B_T15_A2 * a1.x, C3B_T15_A1 * c1.y + C3B_T15_A2 * a1.y);
    }
    if constexpr (EPI == 4) {
        const cplx s0 = E1[idx], s1 = E1[idx + 1], a0 = E3[idx], a1 = E3[idx + 1];
        E2[idx] = cmake(c0.x + C3B_T15_B1 * s0.x + C3B_T15_B2 * a0.x, c0.y + C3B_T15_B1 * s0.y + C3B_T15_B2 * a0.y);
        E2[idx + 1] = cmake(c1.x + C3B_T15_B1 * s1.x + C3B_T15_B2 * a1.x, c1.y + C3B_T15_B1 * s1.y + C3B_T15_B2 * a1.y);
        c0 = cmake(c0.x + C3B_T15_B3 * s0.x, c0.y + C3B_T15_B3 * s0.y);
        c1 = cmake(c1.x + C3B_T15_B3 * s1.x, c1.y + C3B_T15_B3 * s1.y);
    }
    if constexpr (EPI == 5) {
        constexpr double kP0 = C3B_T15_B4 + C3B_T15_B5, kS = C3B_T15_B4 * C3B_T15_B1, kA = C3B_T15_B4 * C3B_T15_B2;
        const cplx r0 = E1[idx], r1 = E1[idx + 1], a0 = E2[idx], a1 = E2[idx + 1], s0 = E3[idx], s1 = E3[idx + 1];
        const cplx q0 = cmake(r0.x - C3B_T15_B3 * s0.x, r0.y - C3B_T15_B3 * s0.y);
        const cplx q1 = cmake(r1.x - C3B_T15_B3 * s1.x, r1.y - C3B_T15_B3 * s1.y);
        const cplx p0 = cmake(c0.x + kP0 * q0.x + kS * s0.x + kA * a0.x, c0.y + kP0 * q0.y + kS * s0.y + kA * a0.y);
        const cplx p1 = cmake(c1.x + kP0 * q1.x + kS * s1.x + kA * a1.x, c1.y + kP0 * q1.y + kS * s1.y + kA * a1.y);
        E2[idx] = cmake(p0.x + C3B_T15_C1 * s0.x + C3B_T15_C2 * a0.x, p0.y + C3B_T15_C1 * s0.y + C3B_T15_C2 * a0.y);
        E2[idx + 1] = cmake(p1.x + C3B_T15_C1 * s1.x + C3B_T15_C2 * a1.x, p1.y + C3B_T15_C1 * s1.y + C3B_T15_C2 * a1.y);
        E3[idx] = cmake(p0.x + C3B_T15_C3 * q0.x + C3B_T15_C4 * a0.x, p0.y + C3B_T15_C3 * q0.y + C3B_T15_C4 * a0.y);
        E3[idx + 1] = cmake(p1.x + C3B_T15_C3 * q1.x + C3B_T15_C4 * a1.x, p1.y + C3B_T15_C3 * q1.y + C3B_T15_C4 * a1.y);
        c0 = cmake(C3B_T15_C9 * p0.x + C3B_T15_C5 * q0.x + C3B_T15_C6 * s0.x + C3B_T15_C7 * a0.x + (row == col ? C3B_T15_C8 : 0.0),
                   C3B_T15_C9 * p0.y + C3B_T15_C5 * q0.y + C3B_T15_C6 * s0.y + C3B_T15_C7 * a0.y);
        c1 = cmake(C3B_T15_C9 * p1.x + C3B_T15_C5 * q1.x + C3B_T15_C6 * s1.x + C3B_T15_C7 * a1.x + (row == col + 1 ? C3B_T15_C8 : 0.0),
                   C3B_T15_C9 * p1.y + C3B_T15_C5 * q1.y + C3B_T15_C6 * s1.y + C3B_T15_C7 * a1.y);
    }
    C[idx] = c0;
    C[idx + 1] = c1;
}

// One macro tile of TI x TJ m8n8 blocks with run-time leading dimension (the global-workspace path): accumulate over k, then the
// (optionally fused) epilogue of cta_zgemm.  Instantiated for every block count a partial macro tile at the matrix edge can
// have, so that those tiles issue DMMAs for the blocks they own only.
template <int TI, int TJ, int EPI>
__device__ __forceinline__ void zgemm_tile(cplx* C, const cplx* A, const cplx* B, const int bi0, const int bj0, const int LD, const int KP,
                                           const cplx* E1, cplx* E2, cplx* E3, const int fr, const int fc) {
    double p1[TI][TJ][2], p2[TI][TJ][2], p3[TI][TJ][2];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) { p1[i][j][0] = p1[i][j][1] = p2[i][j][0] = p2[i][j][1] = p3[i][j][0] = p3[i][j][1] = 0.0; }
    int arow[TI], bcol[TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i) arow[i] = ((bi0 + i) * 8 + fr) * LD + fc;
#pragma unroll
    for (int j = 0; j < TJ; ++j) bcol[j] = fc * LD + (bj0 + j) * 8 + fr;
#pragma unroll 2
    for (int k0 = 0; k0 < KP; k0 += 4) {
        cplx a[TI], b[TJ];
#pragma unroll
        for (int i = 0; i < TI; ++i) a[i] = A[arow[i] + k0];
#pragma unroll
        for (int j = 0; j < TJ; ++j) b[j] = B[bcol[j] + k0 * LD];
        double as[TI], bs[TJ];
#pragma unroll
        for (int i = 0; i < TI; ++i) as[i] = a[i].x + a[i].y;
#pragma unroll
        for (int j = 0; j < TJ; ++j) bs[j] = b[j].x + b[j].y;
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                dmma8x8x4(p1[i][j][0], p1[i][j][1], a[i].x, b[j].x);
                dmma8x8x4(p2[i][j][0], p2[i][j][1], a[i].y, b[j].y);
                dmma8x8x4(p3[i][j][0], p3[i][j][1], as[i], bs[j]);
            }
    }
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const int row = (bi0 + i) * 8 + fr, col = (bj0 + j) * 8 + 2 * fc;
            const cplx c0 = cmake(p1[i][j][0] - p2[i][j][0], p3[i][j][0] - p1[i][j][0] - p2[i][j][0]);
            const cplx c1 = cmake(p1[i][j][1] - p2[i][j][1], p3[i][j][1] - p1[i][j][1] - p2[i][j][1]);
            zgemm_epilogue<EPI>(c0, c1, row * LD + col, row, col, C, E1, E2, E3);
        }
}

// dispatch on the (warp-uniform) block counts of a partial macro tile
template <int TM, int TN, int EPI>
__device__ __forceinline__ void zgemm_edge_tile(cplx* C, const cplx* A, const cplx* B, const int bi0, const int bj0, const int ti, const int tj,
                                                const int LD, const int KP, const cplx* E1, cplx* E2, cplx* E3, const int fr, const int fc) {
    if constexpr (TM >= 3) {
        if (ti == 2 && tj == TN) return zgemm_tile<2, TN, EPI>(C, A, B, bi0, bj0, LD, KP, E1, E2, E3, fr, fc);
        if constexpr (TN >= 2) { if (ti == 2 && tj == 1) return zgemm_tile<2, 1, EPI>(C, A, B, bi0, bj0, LD, KP, E1, E2, E3, fr, fc); }
    }
    if constexpr (TM >= 2) {
        if (ti == 1 && tj == TN) return zgemm_tile<1, TN, EPI>(C, A, B, bi0, bj0, LD, KP, E1, E2, E3, fr, fc);
    }
    if constexpr (TN >= 2) {
        if (ti == TM && tj == 1) return zgemm_tile<TM, 1, EPI>(C, A, B, bi0, bj0, LD, KP, E1, E2, E3, fr, fc);
    }
    if (ti == 1 && tj == 1) return zgemm_tile<1, 1, EPI>(C, A, B, bi0, bj0, LD, KP, E1, E2, E3, fr, fc);
    // block counts without an instance (TM, TN > 3): one block at a time
    for (int i = 0; i < ti; ++i)
        for (int j = 0; j < tj; ++j) zgemm_tile<1, 1, EPI>(C, A, B, bi0 + i, bj0 + j, LD, KP, E1, E2, E3, fr, fc);
}

// ---- cp.async (Ampere-style asynchronous global -> shared copies, 16 bytes per thread) ----------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gptr) {
    const unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(__cvta_generic_to_global(gptr)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// C = A B with the k-panels of BOTH operands staged through shared memory (global-workspace matrices, 64 < DP <= 8 NBMAX):
// the per-warp macro tiles of cta_zgemm read every operand fragment straight from the L2-resident workspace, and at D = 81
// the product waited on those loads (long_scoreboard 14.8 stalls per issue, tensor pipe 51 %).  Here the CTA streams panels
// of 8 k-values -- A[:, k0:k0+8] and B[k0:k0+8, :] -- into a 3-stage ring with cp.async, one barrier per panel, and a warp
// owns HALF a block-row of C: one A fragment and 6 B fragments per k-step feed 18 DMMAs (3M product), its accumulators stay
// in registers across the panel loop, every operand byte leaves L2 once per product.
// Panel layouts: A rows padded to 12 elements, B rows to DP + 2: both fragment patterns of mma.m8n8k4 are then conflict-free
// per quarter-warp.  Needs DP == 8 NBMAX, 2 NBMAX <= NT / 32 warps and 3 (12 DP + 8 (DP + 2)) 16 bytes of dynamic shared memory at offset 0.
// __noinline__: a real call gives the panel loop its own register allocation (inlined into the kernel, whose ~40 live
// values surround every product, the accumulators spilled inside the loop: 0.6 local-memory accesses per DMMA).
template <int NBMAX, int NT, int EPI>
__device__ __noinline__ void cta_zgemm_rows(cplx* C, const cplx* A, const cplx* B, const int LD, const int KP,
                                            const cplx* E1, cplx* E2, cplx* E3) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KB = 8, LDA = 12, NSTG = 3, DP = 8 * NBMAX, LDB = DP + 2;
    constexpr int PER = (DP * KB + NT - 1) / NT;           // panel elements per thread and operand (1 at NT = 704, 2 at 384)
    constexpr bool HALF = (NT / 32 >= 2 * NBMAX);          // enough warps to split every block-row in two
    const int npan = (KP + KB - 1) / KB;
    constexpr int stage_elems = DP * LDA + KB * LDB;
    cplx* stage = reinterpret_cast<cplx*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fc = lane & 3;
    // this thread's elements of every A panel (row e / 8, k e % 8) and of every B panel (k e / DP, column e % DP)
    const cplx* a_src[PER];
    const cplx* b_src[PER];
    int a_dst[PER], b_dst[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int e = min(tid + q * NT, DP * KB - 1);      // (a clamped duplicate copies the same element twice: harmless)
        a_src[q] = A + (size_t)(e >> 3) * LD + (e & 7);
        b_src[q] = B + (size_t)(e / DP) * LD + (e % DP);
        a_dst[q] = (e >> 3) * LDA + (e & 7);
        b_dst[q] = DP * LDA + (e / DP) * LDB + (e % DP);
    }
    auto issue = [&](const int pan) {
        cplx* st = stage + (pan % NSTG) * stage_elems;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            cp_async16(st + a_dst[q], a_src[q] + pan * KB);
            cp_async16(st + b_dst[q], b_src[q] + (size_t)pan * KB * LD);
        }
        cp_async_commit();
    };
    // Work split.  HALF (>= 2 NBMAX warps): warp w owns block-row w / 2, column half w % 2 (blocks 0..NJ-1 or NBMAX-NJ..NBMAX-1;
    // with NBMAX odd the middle block is computed twice and stored by the first half -- branch-free).  Otherwise warp w owns
    // R = ceil(NBMAX / warps) whole block-rows w R .. w R + R - 1 (a row past the end is a clamped duplicate that is not stored).
    constexpr int NW = NT / 32;
    constexpr int R = HALF ? 1 : (NBMAX + NW - 1) / NW;
    constexpr int NJ = HALF ? (NBMAX + 1) / 2 : NBMAX;
    const int brow0 = HALF ? (warp >> 1) : warp * R, half = HALF ? (warp & 1) : 0, jbase = half ? NBMAX - NJ : 0;
    double p1[R][NJ][2], p2[R][NJ][2], p3[R][NJ][2];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { p1[r][j][0] = p1[r][j][1] = p2[r][j][0] = p2[r][j][1] = p3[r][j][0] = p3[r][j][1] = 0.0; }
    int a_off[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a_off[r] = (min(brow0 + r, NBMAX - 1) * 8 + fr) * LDA + fc;
    issue(0);
    if (npan > 1) issue(1);
    for (int pan = 0; pan < npan; ++pan) {
        if (pan + 1 < npan) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();                                   // panel pan has landed for everyone; stage (pan + 2) % 3 is free
        if (pan + 2 < npan) issue(pan + 2);
        if (brow0 < NBMAX) {
            const cplx* As = stage + (pan % NSTG) * stage_elems;
            const cplx* Bs = stage + (pan % NSTG) * stage_elems + DP * LDA + fc * LDB + jbase * 8 + fr;
            auto kstep = [&](const int ks) {
                cplx a[R];
                double as[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { a[r] = As[a_off[r] + ks * 4]; as[r] = a[r].x + a[r].y; }
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    // at most two B fragments in flight: hoisting all of them spilled the accumulators inside the loop
                    if ((j & 1) == 0) asm volatile("" ::: "memory");
                    const cplx b = Bs[ks * 4 * LDB + j * 8];
                    const double bs = b.x + b.y;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        dmma8x8x4(p1[r][j][0], p1[r][j][1], a[r].x, b.x);
                        dmma8x8x4(p2[r][j][0], p2[r][j][1], a[r].y, b.y);
                        dmma8x8x4(p3[r][j][0], p3[r][j][1], as[r], bs);
                    }
                }
            };
            kstep(0);
            if (KP - pan * KB > 4) kstep(1);               // the last panel may hold one k-step only
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (brow0 + r < NBMAX) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (half == 0 || 2 * NJ == NBMAX || j > 0) {   // the duplicated middle block belongs to the first half
                    const int row = (brow0 + r) * 8 + fr, col = (jbase + j) * 8 + 2 * fc;
                    const cplx c0 = cmake(p1[r][j][0] - p2[r][j][0], p3[r][j][0] - p1[r][j][0] - p2[r][j][0]);
                    const cplx c1 = cmake(p1[r][j][1] - p2[r][j][1], p3[r][j][1] - p1[r][j][1] - p2[r][j][1]);
                    zgemm_epilogue<EPI>(c0, c1, row * LD + col, row, col, C, E1, E2, E3);
                }
            }
        }
    }
}

// C = A * B for zero-padded DP x DP complex matrices (row-major, leading dimension DP, DP % 8 == 0).
// Warp w owns macro tiles of TM x TN m8n8 blocks, assigned round-robin.  No __restrict__: operands
// may be global-workspace buffers written earlier by this CTA.
// LD is the leading dimension (>= DP); KP = D rounded up to 4 bounds the k loop (columns/rows beyond D
// are zero).  In shared memory LD = DP + 4 (= 4 mod 8 in 16-byte units) makes the A-fragment loads
// bank-conflict free and the B-fragment loads 2-way (with LD = DP = 32 they were 8-way / 4-way).
// EPI fuses the element-wise step that follows two of the scheme's products into the epilogue (the accumulators are still in
// registers), see zgemm_epilogue.  Padding rows / columns are zero in every operand, so they stay zero (mode 5 puts the
// identity on the padding diagonal: block-diagonal, never read).
template <int TM, int TN, int DPT, int KST, int NT, int EPI>
__device__ __forceinline__ void cta_zgemm_tiles(cplx* C, const cplx* A, const cplx* B, const int DP_, const int LD_, const int KP,
                                                const cplx* E1, cplx* E2, const unsigned char* sched, cplx* E3) {
    // DPT > 0: tile extent and leading dimension are compile-time (DPT, DPT + 4): the k loop unrolls fully and
    // every fragment address is base + immediate
    constexpr bool SWZ = (DPT == 32);
    const int DP = DPT > 0 ? DPT : DP_;
    const int LD = DPT > 0 ? (SWZ ? DPT : DPT + 4) : LD_;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    const int nb = DP >> 3;                       // m8n8 blocks per dimension
    const int mt_r = (nb + TM - 1) / TM, mt_c = (nb + TN - 1) / TN;
    const int fr = lane >> 2, fc = lane & 3;      // fragment coordinates
    // macro tiles of this warp: round-robin, or (sched) the balanced list built by gemm_build_schedule -- partial macro tiles
    // at the matrix edge carry fewer blocks, and their DMMAs are skipped
    const int rounds = (mt_r * mt_c + NW - 1) / NW;
    for (int rd = 0; rd < rounds; ++rd) {
        const int mt = sched ? (int)sched[warp * rounds + rd] : warp + rd * NW;
        if (mt >= mt_r * mt_c) continue;
        const int bi0 = (mt / mt_c) * TM, bj0 = (mt % mt_c) * TN;
        if constexpr (DPT == 0 && (TM > 1 || TN > 1)) {
            // run-time extents: a macro tile at the matrix edge is computed by the instance with exactly its block counts
            const int ti = min(TM, nb - bi0), tj = min(TN, nb - bj0);
            if (ti != TM || tj != TN) {
                zgemm_edge_tile<TM, TN, EPI>(C, A, B, bi0, bj0, ti, tj, LD, KP, E1, E2, E3, fr, fc);
                continue;
            }
        }
        // 3M complex product: P1 = Ar Br, P2 = Ai Bi, P3 = (Ar + Ai)(Br + Bi); Cr = P1 - P2, Ci = P3 - P1 - P2 -- three real
        // DMMAs per tile step instead of four (the operand sums cost one DADD per loaded fragment element)
        double p1[TM][TN][2], p2[TM][TN][2], p3[TM][TN][2];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) { p1[i][j][0] = p1[i][j][1] = p2[i][j][0] = p2[i][j][1] = p3[i][j][0] = p3[i][j][1] = 0.0; }
        // clamp block indices of partial macro tiles (results of clamped duplicates are not stored)
        // arow: address of A(row, fc) for even k steps, arow_odd for odd ones (swizzled: (4 kk + fc) ^ f(row) =
        // 4 kk + (fc ^ (f & 2)) +- (f & 4), the sign alternating with the parity of kk); bcol: address of B(fc, col)
        int arow[TM], arow_odd[TM], bcol[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int row = min(bi0 + i, nb - 1) * 8 + fr;
            if constexpr (SWZ) {
                const int f = swz_of_row(row);
                arow[i] = row * 32 + (fc ^ (f & 2)) + (f & 4);
                arow_odd[i] = row * 32 + (fc ^ (f & 2)) - (f & 4);
            } else {
                arow[i] = row * LD + fc;
                arow_odd[i] = arow[i];
            }
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if constexpr (SWZ) bcol[j] = fc * 32 + min(bj0 + j, nb - 1) * 8 + (fr ^ swz_of_row(fc));
            else bcol[j] = fc * LD + min(bj0 + j, nb - 1) * 8 + fr;
        }
        auto kstep = [&](const int k0) {
            cplx a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = A[arow[i] + k0];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = B[bcol[j] + k0 * LD];
            double as[TM], bs[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) as[i] = a[i].x + a[i].y;
#pragma unroll
            for (int j = 0; j < TN; ++j) bs[j] = b[j].x + b[j].y;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    dmma8x8x4(p1[i][j][0], p1[i][j][1], a[i].x, b[j].x);
                    dmma8x8x4(p2[i][j][0], p2[i][j][1], a[i].y, b[j].y);
                    dmma8x8x4(p3[i][j][0], p3[i][j][1], as[i], bs[j]);
                }
        };
        if (DPT > 0) {
            // software pipeline by hand: the fragments of k step kk+1 are requested before the DMMAs of step kk are
            // issued.  ptxas otherwise sinks every LDS right in front of its first use (minimal registers) and both
            // warps of a scheduler eat the shared-memory latency at every k step; the __syncwarp() is a scheduling
            // fence that loads may not sink below.
            constexpr int KS = KST > 0 ? KST : (DPT > 0 ? DPT / 4 : 1);   // k steps: ceil(D / 4), compile-time
            cplx af[KS][TM], bf[KS][TN];
            auto fetch = [&](const int kk) {
#pragma unroll
                for (int i = 0; i < TM; ++i) af[kk][i] = A[((kk & 1) ? arow_odd[i] : arow[i]) + kk * 4];
#pragma unroll
                for (int j = 0; j < TN; ++j) bf[kk][j] = B[bcol[j] + kk * 4 * LD];
            };
            fetch(0);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                if (kk + 1 < KS) fetch(kk + 1);
                __syncwarp();
                double as[TM], bs[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) as[i] = af[kk][i].x + af[kk][i].y;
#pragma unroll
                for (int j = 0; j < TN; ++j) bs[j] = bf[kk][j].x + bf[kk][j].y;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        dmma8x8x4(p1[i][j][0], p1[i][j][1], af[kk][i].x, bf[kk][j].x);
                        dmma8x8x4(p2[i][j][0], p2[i][j][1], af[kk][i].y, bf[kk][j].y);
                        dmma8x8x4(p3[i][j][0], p3[i][j][1], as[i], bs[j]);
                    }
            }
        } else {
            // fragments of the next k step are requested before the DMMAs of this one are issued (operands come from L1 / L2:
            // long_scoreboard was the top stall at 7.6 per issue)
            cplx a[TM], b[TN], an[TM], bn[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = A[arow[i]];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = B[bcol[j]];
#pragma unroll 1
            for (int k0 = 0; k0 < KP; k0 += 4) {
                const int kn = (k0 + 4 < KP) ? k0 + 4 : k0;
#pragma unroll
                for (int i = 0; i < TM; ++i) an[i] = A[arow[i] + kn];
#pragma unroll
                for (int j = 0; j < TN; ++j) bn[j] = B[bcol[j] + kn * LD];
                double as[TM], bs[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) as[i] = a[i].x + a[i].y;
#pragma unroll
                for (int j = 0; j < TN; ++j) bs[j] = b[j].x + b[j].y;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        dmma8x8x4(p1[i][j][0], p1[i][j][1], a[i].x, b[j].x);
                        dmma8x8x4(p2[i][j][0], p2[i][j][1], a[i].y, b[j].y);
                        dmma8x8x4(p3[i][j][0], p3[i][j][1], as[i], bs[j]);
                    }
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = an[i];
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = bn[j];
            }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                if (bi0 + i < nb && bj0 + j < nb) {
                    const int row = (bi0 + i) * 8 + fr, col = (bj0 + j) * 8 + 2 * fc;
                    const cplx c0 = cmake(p1[i][j][0] - p2[i][j][0], p3[i][j][0] - p1[i][j][0] - p2[i][j][0]);
                    const cplx c1 = cmake(p1[i][j][1] - p2[i][j][1], p3[i][j][1] - p1[i][j][1] - p2[i][j][1]);
                    zgemm_epilogue<EPI>(c0, c1, mat_idx<SWZ>(row, col, LD), row, col, C, E1, E2, E3);   // idx + 1: the next column (f is even)
                }
            }
    }
}

// TM > 0: per-warp macro tiles (cta_zgemm_tiles); TM == 0: staged block-row product with at most TN blocks per row
template <int TM, int TN, int DPT = 0, int KST = 0, int NT = kCtaThreads, int EPI = 0>
__device__ __forceinline__ void cta_zgemm(cplx* C, const cplx* A, const cplx* B, const int DP_, const int LD_, const int KP,
                                          const cplx* E1 = nullptr, cplx* E2 = nullptr, const unsigned char* sched = nullptr,
                                          cplx* E3 = nullptr) {
    if constexpr (TM == 0) cta_zgemm_rows<TN, NT, EPI>(C, A, B, LD_, KP, E1, E2, E3);
    else cta_zgemm_tiles<TM, TN, DPT, KST, NT, EPI>(C, A, B, DP_, LD_, KP, E1, E2, sched, E3);
}

// Balanced static assignment of macro tiles to warps (longest-processing-time first): a macro tile at the matrix edge holds
// fewer m8n8 blocks than TM x TN, and round-robin leaves some warps with full tiles only (D = 81, 3 x 2 macro tiles on 11 x 11
// blocks, 8 warps: 18 blocks on the busiest warp against 15.1 on average; balanced: 16).  sched[w * rounds + r] = macro tile
// index or 255.  Called by one thread; at most 64 macro tiles.
template <int TM, int TN>
__device__ inline void gemm_build_schedule(unsigned char* sched, int* load, const int nb, const int nw) {
    const int mt_r = (nb + TM - 1) / TM, mt_c = (nb + TN - 1) / TN, nt = mt_r * mt_c;
    const int rounds = (nt + nw - 1) / nw;
    for (int w = 0; w < nw; ++w) {
        load[w] = 0;
        load[nw + w] = 0;     // tiles taken
        for (int r = 0; r < rounds; ++r) sched[w * rounds + r] = 255;
    }
    for (int size = TM * TN; size >= 1; --size)
        for (int mt = 0; mt < nt; ++mt) {
            const int bi0 = (mt / mt_c) * TM, bj0 = (mt % mt_c) * TN;
            const int blocks = min(TM, nb - bi0) * min(TN, nb - bj0);
            if (blocks != size) continue;
            int best = -1;
            for (int w = 0; w < nw; ++w)
                if (load[nw + w] < rounds && (best < 0 || load[w] < load[best])) best = w;
            sched[best * rounds + load[nw + best]] = (unsigned char)mt;
            load[best] += blocks;
            load[nw + best] += 1;
        }
}

// inf-norm (largest row sum of |a_ij|) over the D x D part of an LD-strided matrix, all threads busy:
// 8 threads per row, shuffle-reduced.  Any subordinate norm bounds the Taylor truncation error the same way.
template <int NT>
__device__ __forceinline__ double cta_norm_inf_ld(const cplx* A, const int D, const int LD, double* red, const int ncols) {
    const int part = threadIdx.x & 7;
    double best = 0.0;
    for (int r = threadIdx.x >> 3; r < ((D + 31) & ~31); r += NT / 8) {
        double s = 0.0;
        if (r < D)
            for (int j = part; j < ncols; j += 8) s += cabs1(A[r * LD + j]);   // swizzled rows: all LD slots (padding is zero)
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        best = fmax(best, s);
    }
    best = warp_max(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    double v = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) v = fmax(v, red[w]);
    __syncthreads();
    return v;
}


// Slot plan (6 matrices of DP x LD), degree-15+ Taylor scheme in FOUR products (c3b_common.cuh); every combination is
// formed in the epilogue of the product before it (zgemm_epilogue), so a slice is 4 + s + 1 products and as many barriers,
// with no element-wise pass:
//   product 1 (EPI 3)  S1 = A A = A2,                 S2 = a1 A2 + a2 A = Q0              (A = S0)
//   product 2 (EPI 4)  S3 = A2 Q0 + b3 A2 = R1',      S4 = A2 Q0 + b1 A2 + b2 A = L1
//   product 3 (EPI 5)  p  = L1 R1';   S0 <- L2,  S1 <- R2,  S2 <- E0                     (A, A2 consumed in place)
//   product 4 (EPI 2)  S3 = L2 R2 + E0 = T                                                (R1' is dead)
//   squarings ping-pong S3 <-> S4;  the fold X P goes to S2 (E0 is dead) and S2 / P swap roles (pointers, no copy);
//   the first slice of a segment swaps X and P instead of copying.
template <int TM, int TN, int DPT = 0, int KST = 0, int NT = kCtaThreads>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) pwc_taylor_cta_kernel(const GemmParams gp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    const CtaParams& p = gp.c;
    const int D = p.D, K = p.K;
    constexpr bool SWZ = (DPT == 32);
    const int DP = DPT > 0 ? DPT : gp.DP;
    const int LD = DPT > 0 ? (SWZ ? DPT : DPT + 4) : gp.LD;
    const int KP = (D + 3) & ~3;
    const int PP = DP * LD;
    const int RL = D * LD;                          // rows >= D are padding: zeroed once, never written non-zero
    const int tid = threadIdx.x;

    // DPT > 0 is only launched with shared-memory matrices: keeping the pointer provably shared lets ptxas emit
    // LDS/STS instead of generic loads
    cplx* mats = (DPT > 0) ? reinterpret_cast<cplx*>(smem_raw)
                           : (p.use_smem ? reinterpret_cast<cplx*>(smem_raw) : p.ws + (size_t)blockIdx.x * kGemmSlots * PP);
    cplx* const S0 = mats;
    cplx* const S1 = mats + (size_t)1 * PP;
    cplx* S2 = mats + (size_t)2 * PP;
    cplx* S3 = mats + (size_t)3 * PP;
    cplx* S4 = mats + (size_t)4 * PP;
    cplx* P = mats + (size_t)5 * PP;
    const bool shifted = gp.TR != nullptr && p.hlist == nullptr;

    __shared__ unsigned char s_sched[64];
    __shared__ int s_load[2 * (NT / 32)];
    const unsigned char* sched = nullptr;
    if constexpr (TM > 0) {
        if (DPT == 0 && ((DP >> 3) + TM - 1) / TM * (((DP >> 3) + TN - 1) / TN) <= 64 - NT / 32) {
            if (tid == 0) gemm_build_schedule<TM, TN>(s_sched, s_load, DP >> 3, NT / 32);
            sched = s_sched;
        }
    }
    // zero everything once: the padding rows/columns stay zero through every product
    for (int e = tid; e < kGemmSlots * PP; e += NT) mats[e] = cmake(0.0, 0.0);
    const cplx* Gs = nullptr;                       // generators staged in shared memory (shared model only)
    if (gp.g_in_smem) {
        cplx* g = reinterpret_cast<cplx*>(smem_raw) + (size_t)kGemmSlots * PP;
        for (int e = tid; e < (K + 1) * D * D; e += NT) g[e] = p.G[e];
        Gs = g;
    }
    __syncthreads();

    const long long units = (long long)p.B * p.S;
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int b = (int)(unit / p.S);
        const int sidx = (int)(unit - (long long)b * p.S);
        const int n_begin = sidx * p.seg_len;
        const int n_end = min(p.N, n_begin + p.seg_len);
        const cplx* Gb = Gs ? Gs : (p.G ? p.G + (size_t)b * p.model_stride : nullptr);
        const cplx* TRb = shifted ? gp.TR + (size_t)b * (p.model_stride ? (K + 1) : 0) : nullptr;
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        const cplx hs = cmake(p.hscale_re, p.hscale_im);
        cplx mu_acc = cmake(0.0, 0.0);

        const double* RSb = (gp.RS != nullptr && p.hlist == nullptr) ? gp.RS + (size_t)b * (p.model_stride ? (size_t)(K + 1) * D : 0) : nullptr;
        for (int n = n_begin; n < n_end; ++n) {
            cplx* A = S0;
            // ---- scaling from the row-sum bound  ||A_n||_inf <= max_r ( rs_0[r] + sum_k |c_k[n]| rs_k[r] ):  known BEFORE the
            //      slice is assembled (every warp computes it, no barrier), so the 2^-s factor is folded into the assembly and
            //      the norm pass over the slice (two barriers, D^2 square roots) disappears.  The bound is exact when the
            //      generators have disjoint supports (diagonal drift + off-diagonal drives), never below the true norm.
            int s_pre = -1;
            double asc = 1.0;
            if (RSb != nullptr) {
                double v = 0.0;
                for (int r = tid & 31; r < D; r += 32) {
                    double t = RSb[r];
                    for (int k = 0; k < K; ++k) t = fma(fabs(__ldg(sig_b + (size_t)k * p.N + n)), RSb[(size_t)(k + 1) * D + r], t);
                    v = fmax(v, t);
                }
                s_pre = squarings_for(warp_max(v), C3B_THETA15);
                asc = pow2neg(s_pre);
            }
            // ---- assemble (D x D part; padding stays zero) ---------------------------------------
            if (p.hlist == nullptr) {
                auto element = [&](const int i, const int j) {
                    const int e = i * D + j;
                    cplx v = Gb[e];
                    for (int k = 0; k < K; ++k) {
                        const double c = __ldg(sig_b + (size_t)k * p.N + n);
                        const cplx gk = Gb[(size_t)(k + 1) * D * D + e];
                        v.x = fma(c, gk.x, v.x);
                        v.y = fma(c, gk.y, v.y);
                    }
                    A[mat_idx<SWZ>(i, j, LD)] = cmake(v.x * asc, v.y * asc);
                };
                if (DPT > 0) {                                   // DPT lanes walk one row: no integer division
                    constexpr int W = DPT > 0 ? DPT : 1;
                    const int j = tid % W;
                    if (j < D)
                        for (int i = tid / W; i < D; i += NT / W) element(i, j);
                } else {
                    if (K <= 3) {                                  // batches of 4 elements, loads first (operands in L2)
                        constexpr int AU = 4;
                        double cs[3] = {0.0, 0.0, 0.0};
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (k < K) cs[k] = __ldg(sig_b + (size_t)k * p.N + n);
                        const int DD = D * D;
                        for (int e0 = tid; e0 < DD; e0 += AU * NT) {
                            cplx g0[AU], g1[AU], g2[AU], g3[AU];
#pragma unroll
                            for (int u = 0; u < AU; ++u) {
                                const int e = min(e0 + u * NT, DD - 1);
                                g0[u] = Gb[e];
                                g1[u] = K > 0 ? Gb[(size_t)DD + e] : cmake(0.0, 0.0);
                                g2[u] = K > 1 ? Gb[(size_t)2 * DD + e] : cmake(0.0, 0.0);
                                g3[u] = K > 2 ? Gb[(size_t)3 * DD + e] : cmake(0.0, 0.0);
                            }
#pragma unroll
                            for (int u = 0; u < AU; ++u) {
                                const int e = e0 + u * NT;
                                if (e < DD) {
                                    cplx v = g0[u];
                                    v.x = fma(cs[0], g1[u].x, v.x); v.y = fma(cs[0], g1[u].y, v.y);
                                    v.x = fma(cs[1], g2[u].x, v.x); v.y = fma(cs[1], g2[u].y, v.y);
                                    v.x = fma(cs[2], g3[u].x, v.x); v.y = fma(cs[2], g3[u].y, v.y);
                                    A[(e / D) * LD + (e % D)] = cmake(v.x * asc, v.y * asc);
                                }
                            }
                        }
                    } else {
                        for (int e = tid; e < D * D; e += NT) element(e / D, e % D);
                    }
                }
                if (shifted) {
                    cplx mu = TRb[0];
                    for (int k = 0; k < K; ++k) {
                        const double c = __ldg(sig_b + (size_t)k * p.N + n);
                        mu.x = fma(c, TRb[k + 1].x, mu.x);
                        mu.y = fma(c, TRb[k + 1].y, mu.y);
                    }
                    mu_acc.x += mu.x; mu_acc.y += mu.y;
                }
            } else {
                const cplx* H = p.hlist + ((size_t)b * p.N + n) * D * D;
                for (int e = tid; e < D * D; e += NT) {
                    const int i = e / D, j = e - i * D;
                    A[mat_idx<SWZ>(i, j, LD)] = cmul(hs, H[e]);
                }
            }
            __syncthreads();
            int s = s_pre;
            if (s_pre < 0) {                                   // explicit slices (H list): exact inf-norm of the slice
                const double nrm = cta_norm_inf_ld<NT>(A, D, LD, red, SWZ ? LD : D);
                s = squarings_for(nrm, C3B_THETA15);
                if (s > 0) {
                    const double sc = pow2neg(s);
                    for (int e = tid; e < RL; e += NT) { A[e].x *= sc; A[e].y *= sc; }
                    __syncthreads();
                }
            }
            // ---- degree-15+ Taylor polynomial in four products (slot plan above) ---------------------------------------
            cta_zgemm<TM, TN, DPT, KST, NT, 3>(S1, A, A, DP, LD, KP, A, S2, sched);
            __syncthreads();
            cta_zgemm<TM, TN, DPT, KST, NT, 4>(S3, S1, S2, DP, LD, KP, S1, S4, sched, S0);
            __syncthreads();
            cta_zgemm<TM, TN, DPT, KST, NT, 5>(S2, S4, S3, DP, LD, KP, S3, S0, sched, S1);
            __syncthreads();
            cta_zgemm<TM, TN, DPT, KST, NT, 2>(S3, S0, S1, DP, LD, KP, S2, nullptr, sched);
            __syncthreads();
            cplx* X = S3;
            cplx* Y = S4;
            for (int i = 0; i < s; ++i) {                          // undo the scaling
                cta_zgemm<TM, TN, DPT, KST, NT>(Y, X, X, DP, LD, KP, nullptr, nullptr, sched);
                __syncthreads();
                cplx* t = X; X = Y; Y = t;
            }
            if (p.dUs_out != nullptr) {
                cplx* o = p.dUs_out + ((size_t)b * p.N + n) * D * D;
                cplx phn = cmake(1.0, 0.0);
                if (shifted) {
                    cplx mu = TRb[0];
                    for (int k = 0; k < K; ++k) {
                        const double c = __ldg(sig_b + (size_t)k * p.N + n);
                        mu.x = fma(c, TRb[k + 1].x, mu.x);
                        mu.y = fma(c, TRb[k + 1].y, mu.y);
                    }
                    phn = cexp_(mu);
                }
                for (int e = tid; e < D * D; e += NT) {
                    const int i = e / D, j = e - i * D;
                    o[e] = cmul(phn, X[mat_idx<SWZ>(i, j, LD)]);
                }
            }
            if (n == n_begin) {                                    // P <- dU_n: swap the roles of the two slots
                S3 = P; P = X; S4 = Y;
            } else {
                cta_zgemm<TM, TN, DPT, KST, NT>(S2, X, P, DP, LD, KP, nullptr, nullptr, sched);                // E0 is dead: its slot takes dU_n P
                __syncthreads();
                cplx* t = P; P = S2; S2 = t;
                S3 = X; S4 = Y;
            }
        }
        cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * D * D) : (p.seg_out + ((size_t)b * p.S + sidx) * D * D);
        if (n_end > n_begin) {
            const cplx phu = shifted ? cexp_(mu_acc) : cmake(1.0, 0.0);
            for (int e = tid; e < D * D; e += NT) {
                const int i = e / D, j = e - i * D;
                o[e] = cmul(phu, P[mat_idx<SWZ>(i, j, LD)]);
            }
        } else {
            for (int e = tid; e < D * D; e += NT) o[e] = cmake((e / D) == (e % D) ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
    }
}

}  // namespace c3b
