// d <= 9 fused PWC propagator kernel on the fp64 tensor-core instruction: ONE WARP PER SLICE CHAIN, every 9 x 9 complex
// product as an 8 x 8 DMMA core plus a 1-wide edge.
//
// Same contract as pwc_blk9_taylor_kernel (assemble -> trace-shifted degree-18 Taylor exponential in 5 products -> ordered
// product; replaces c3/libraries/propagation.py:426-440,460-515 and c3/utils/tf_utils.py:120-193).
//
// Why.  The lane-group kernels (pwc_blk9.cuh, pwc_shfl9.cuh) spend 540 issued instructions per 3 matrix products (324
// DFMA + operand traffic) and hold ~250 registers, i.e. two warps per scheduler: issue slots, the shared-memory pipe and
// the fp64 pipe are all within 1.4x of each other and 5 of 32 lanes idle (profiles/README_r02.md) -- they saturate near
// half the fp64 peak.  mma.sync.m8n8k4.f64 retires 256 FMAs per issued instruction from two operand registers per lane.
// A 9 x 9 product does not tile by 8, but 9 = 8 + 1 does not need padding to 16:
//   C[0:8,0:8]  = sum_k X[0:8,k] Y[k,0:8], k = 0..8: three k steps of m8n8k4 (the third carries the single k = 8 term) with the
//                 3M complex product (Ar Br, Ai Bi, (Ar + Ai)(Br + Bi)) = 9 DMMAs;
//   C[0:8,8], C[8,0:8], C[8,8]: 17 dot products of length 9.  Lane (r,q) already holds X[r,4j+q] (its A fragments) and
//                 lane (k,c) holds Y[4j+k,c] (its B fragments): three complex MACs per lane and a 4-lane shuffle reduction per edge.
// Per product: 9 DMMA + 36 DFMA + 12 LDS.128 + 24 SHFL instead of 180 instructions per matrix, ~60 registers, so 16 warps per
// SM hide the DMMA and shared-memory latencies by thread-level parallelism.  Matrices live in per-warp shared memory
// (9 rows x 12 columns, zero-padded columns: the A-fragment loads are bank-conflict free, the B-fragment loads 2-way), six
// slots with the in-place plan of pwc_gemm.cuh.
#pragma once
#include "c3b_params.cuh"

namespace c3b {

struct Dmma9 {
    static constexpr int LD = 12;                 // leading dimension (16-byte units): rows r, r+1 fall in opposite bank halves
    static constexpr int MAT = 9 * LD;            // one matrix slot
    static constexpr int SLOTS = 6;               // S0..S4 scratch + running product
    __host__ __device__ static size_t smem_bytes(int K, int warps) {
        return (size_t)(K + 1) * MAT * sizeof(cplx) + (size_t)(((K + 1) * 9 + 1) & ~1) * sizeof(double) +
               (size_t)warps * SLOTS * MAT * sizeof(cplx);
    }
};

__device__ __forceinline__ void dmma9_mma(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ cplx quad_sum(cplx v) {            // sum over the 4 lanes of a quad (lanes 4i .. 4i+3)
    v.x += __shfl_xor_sync(0xffffffffu, v.x, 1); v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
    v.x += __shfl_xor_sync(0xffffffffu, v.x, 2); v.y += __shfl_xor_sync(0xffffffffu, v.y, 2);
    return v;
}

// C = X Y (+ E1; EPI == 1 also E2 += C) for 9 x 9 complex matrices in shared memory (leading dimension 12), one warp.
// EPI: 0 plain, 1: C = X Y + E1, E2 += C, 2: C = X Y + E1.  C may alias E1 (every element is read and written by the same lane)
// but not X or Y.  The caller separates products by __syncwarp().
template <int EPI>
__device__ __forceinline__ void warp_zgemm9(cplx* C, const cplx* X, const cplx* Y, const int lane, const cplx* E1 = nullptr, cplx* E2 = nullptr) {
    constexpr int LD = Dmma9::LD;
    const int r = lane >> 2, q = lane & 3;        // A fragment: row r, k = 4j + q.  B fragment: k = 4j + q, column r.  C: row r, columns 2q, 2q+1
    const bool first = q == 0;
    cplx a[3], b[3], xr8[3], yc8[3];
    const cplx zero = cmake(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool ok = j < 2 || first;           // k = 4j + q <= 8
        const int k = ok ? 4 * j + q : 0;
        a[j] = ok ? X[r * LD + k] : zero;         // X[r, k]
        b[j] = ok ? Y[k * LD + r] : zero;         // Y[k, c = r]
        xr8[j] = ok ? X[8 * LD + k] : zero;       // X[8, k]  (row-8 edge and corner)
        yc8[j] = ok ? Y[k * LD + 8] : zero;       // Y[k, 8]  (column-8 edge and corner)
    }
    double p1[2] = {0.0, 0.0}, p2[2] = {0.0, 0.0}, p3[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        dmma9_mma(p1[0], p1[1], a[j].x, b[j].x);
        dmma9_mma(p2[0], p2[1], a[j].y, b[j].y);
        dmma9_mma(p3[0], p3[1], a[j].x + a[j].y, b[j].x + b[j].y);
    }
    // edges: column 8 (rows 0..7: partial over this lane's k, then the quad), row 8 (columns 0..7), corner
    cplx ec = zero, er = zero, cc = zero;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        cfma(ec, a[j], yc8[j]);                   // X[r, k] Y[k, 8]
        cfma(er, xr8[j], b[j]);                   // X[8, k] Y[k, c]
        cfma(cc, xr8[j], yc8[j]);                 // X[8, k] Y[k, 8]
    }
    ec = quad_sum(ec);
    er = quad_sum(er);
    cc = quad_sum(cc);
    cplx c0 = cmake(p1[0] - p2[0], p3[0] - p1[0] - p2[0]);
    cplx c1 = cmake(p1[1] - p2[1], p3[1] - p1[1] - p2[1]);
    const int i0 = r * LD + 2 * q;
    // the lane with q == 0 of quad r also owns C[r, 8] and C[8, r]; lane 0 owns the corner
    const int ie = r * LD + 8, ir = 8 * LD + r, ic = 8 * LD + 8;
    if constexpr (EPI != 0) {
        const cplx e0 = E1[i0], e1 = E1[i0 + 1];
        c0.x += e0.x; c0.y += e0.y; c1.x += e1.x; c1.y += e1.y;
        if (first) {
            const cplx f0 = E1[ie], f1 = E1[ir];
            ec.x += f0.x; ec.y += f0.y; er.x += f1.x; er.y += f1.y;
            if (lane == 0) { const cplx f2 = E1[ic]; cc.x += f2.x; cc.y += f2.y; }
        }
    }
    if constexpr (EPI == 1) {
        const cplx g0 = E2[i0], g1 = E2[i0 + 1];
        E2[i0] = cmake(g0.x + c0.x, g0.y + c0.y);
        E2[i0 + 1] = cmake(g1.x + c1.x, g1.y + c1.y);
        if (first) {
            const cplx h0 = E2[ie], h1 = E2[ir];
            E2[ie] = cmake(h0.x + ec.x, h0.y + ec.y);
            E2[ir] = cmake(h1.x + er.x, h1.y + er.y);
            if (lane == 0) { const cplx h2 = E2[ic]; E2[ic] = cmake(h2.x + cc.x, h2.y + cc.y); }
        }
    }
    C[i0] = c0;
    C[i0 + 1] = c1;
    if (first) {
        C[ie] = ec;
        C[ir] = er;
        if (lane == 0) C[ic] = cc;
    }
}

template <int WARPS, int MINB, int GATED>
__global__ void __launch_bounds__(WARPS * 32, MINB) pwc_dmma9_kernel(const RowsParams p, unsigned int* __restrict__ counter) {
    constexpr int D = 9, LD = Dmma9::LD, MAT = Dmma9::MAT, NT = WARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);                 // [(K+1)][9][12] zero padded
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * MAT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cplx* mats = reinterpret_cast<cplx*>(sRS + (((K + 1) * D + 1) & ~1)) + (size_t)warp * Dmma9::SLOTS * MAT;
    const bool hmode = p.hlist != nullptr;

    for (int idx = tid; idx < (K + 1) * MAT; idx += NT) {
        const int k = idx / MAT, rem = idx - k * MAT, i = rem / LD, j = rem - i * LD;
        cplx v = cmake(0.0, 0.0);
        if (!hmode && i < d && j < d) v = p.G[(size_t)k * d * d + i * d + j];
        sG[idx] = v;
    }
    for (int idx = tid; idx < (K + 1) * D; idx += NT) {
        const int k = idx / D, r = idx - k * D;
        sRS[idx] = (!hmode && r < d) ? p.RS[k * d + r] : 0.0;
    }
    for (int e = lane; e < Dmma9::SLOTS * MAT; e += 32) mats[e] = cmake(0.0, 0.0);     // padding columns stay zero
    __syncthreads();

    cplx* const S0 = mats;
    cplx* const S1 = mats + MAT;
    cplx* const S2 = mats + 2 * MAT;
    cplx* const S3 = mats + 3 * MAT;
    cplx* S4 = mats + 4 * MAT;
    cplx* P = mats + 5 * MAT;
    const cplx hs = cmake(p.hscale_re, p.hscale_im);
    const long long total_units = (long long)p.B * p.S;
    const bool shifted = (p.TR != nullptr) && !hmode;
    // element ownership of the element-wise passes: e = lane, lane + 32, lane + 64 over the 81 entries
    int eidx[3];
    bool ediag[3], evalid[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int e = lane + 32 * u;
        evalid[u] = e < 81;
        const int i = evalid[u] ? e / 9 : 0, j = evalid[u] ? e - 9 * (e / 9) : 0;
        eidx[u] = i * LD + j;
        ediag[u] = evalid[u] && i == j;
    }

    for (;;) {
        unsigned int unit_u = 0;
        if (lane == 0) {
            unit_u = atomicAdd(counter, 1u);
            if constexpr (GATED != 0) {
                // gated launch: lane 0 waits until this unit's batch row has landed (rows arrive in order); a row that never
                // arrives ends this warp's work like an exhausted counter and raises gate[1]
                if ((long long)unit_u < total_units) {
                    unsigned int known = 0;
                    if (!wait_rows_ready(p.gate, (int)(unit_u / (unsigned int)p.S), known)) unit_u = 0xffffffffu;
                }
            }
        }
        unit_u = __shfl_sync(0xffffffffu, unit_u, 0);
        if constexpr (GATED != 0) __syncwarp();
        const long long unit = unit_u;
        if (unit >= total_units) break;
        const int b = (int)(unit / p.S);
        const int sidx = (int)(unit - (long long)b * p.S);
        const int n_begin = sidx * p.seg_len;
        const int n_end = min(p.N, n_begin + p.seg_len);
        const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
        cplx mu_acc = cmake(0.0, 0.0);

#pragma unroll 1
        for (int n = n_begin; n < n_end; ++n) {
            // ---- norm bound -> squarings, trace shift ------------------------------------------------------------------
            cplx mu = cmake(0.0, 0.0);
            double nb = 0.0;
            if (!hmode) {
                if (lane < D) {
                    nb = sRS[lane];
                    for (int k = 0; k < K; ++k) nb = fma(fabs(load_signal<GATED>(sig_b + (size_t)k * p.N + n)), sRS[(k + 1) * D + lane], nb);
                }
                if (shifted) {
                    mu = p.TR[0];
                    for (int k = 0; k < K; ++k) {
                        const double cs = load_signal<GATED>(sig_b + (size_t)k * p.N + n);
                        const cplx t = p.TR[k + 1];
                        mu.x = fma(cs, t.x, mu.x); mu.y = fma(cs, t.y, mu.y);
                    }
                }
            } else {
                // explicit slice: A = hscale * H; inf-norm from the row sums (lane r sums row r)
                const cplx* H = p.hlist + ((size_t)b * p.N + n) * d * d;
                if (lane < d)
                    for (int j = 0; j < d; ++j) nb += cabs1(cmul(hs, H[lane * d + j]));
            }
            nb = __hiloint2double((int)__reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(nb) + 1u), 0);
            mu_acc.x += mu.x; mu_acc.y += mu.y;
            const int s = squarings_for(nb, C3B_THETA18);
            const double sc = pow2neg(s);
            // ---- assemble A_n / 2^s into S0 ------------------------------------------------------------------------------
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (!evalid[u]) continue;
                cplx v;
                if (!hmode) {
                    v = sG[eidx[u]];
                    for (int k = 0; k < K; ++k) {
                        const double cs = load_signal<GATED>(sig_b + (size_t)k * p.N + n);
                        const cplx gk = sG[(k + 1) * MAT + eidx[u]];
                        v.x = fma(cs, gk.x, v.x); v.y = fma(cs, gk.y, v.y);
                    }
                } else {
                    const int e = lane + 32 * u, i = e / 9, j = e - 9 * i;
                    v = (i < d && j < d) ? cmul(hs, p.hlist[((size_t)b * p.N + n) * d * d + (size_t)i * d + j]) : cmake(0.0, 0.0);
                }
                S0[eidx[u]] = cmake(v.x * sc, v.y * sc);
            }
            __syncwarp();
            // ---- T18: A2 = S1, A3 = S2, A6 = S3 ------------------------------------------------------------------------------
            warp_zgemm9<0>(S1, S0, S0, lane);
            __syncwarp();
            warp_zgemm9<0>(S2, S1, S0, lane);
            __syncwarp();
            warp_zgemm9<0>(S3, S2, S2, lane);
            __syncwarp();
            // combinations in place: B1 -> S0, B5 -> S1, B4 -> S2, B3 -> S3, B2 -> S4 (every entry by its owning lane)
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (!evalid[u]) continue;
                const int e = eidx[u];
                const cplx x1 = S0[e], x2 = S1[e], x3 = S2[e], x6 = S3[e];
                const double dg = ediag[u] ? 1.0 : 0.0;
                S0[e] = cmake(C3B_T18_A11 * x1.x + C3B_T18_A21 * x2.x + C3B_T18_A31 * x3.x,
                              C3B_T18_A11 * x1.y + C3B_T18_A21 * x2.y + C3B_T18_A31 * x3.y);
                S1[e] = cmake(C3B_T18_B24 * x2.x + C3B_T18_B34 * x3.x + C3B_T18_B64 * x6.x,
                              C3B_T18_B24 * x2.y + C3B_T18_B34 * x3.y + C3B_T18_B64 * x6.y);
                S2[e] = cmake(C3B_T18_B03 * dg + C3B_T18_B13 * x1.x + C3B_T18_B23 * x2.x + C3B_T18_B33 * x3.x + C3B_T18_B63 * x6.x,
                              C3B_T18_B13 * x1.y + C3B_T18_B23 * x2.y + C3B_T18_B33 * x3.y + C3B_T18_B63 * x6.y);
                S3[e] = cmake(C3B_T18_B02 * dg + C3B_T18_B12 * x1.x + C3B_T18_B22 * x2.x + C3B_T18_B32 * x3.x + C3B_T18_B62 * x6.x,
                              C3B_T18_B12 * x1.y + C3B_T18_B22 * x2.y + C3B_T18_B32 * x3.y + C3B_T18_B62 * x6.y);
                S4[e] = cmake(C3B_T18_B11 * x1.x + C3B_T18_B21 * x2.x + C3B_T18_B31 * x3.x + C3B_T18_B61 * x6.x,
                              C3B_T18_B11 * x1.y + C3B_T18_B21 * x2.y + C3B_T18_B31 * x3.y + C3B_T18_B61 * x6.y);
            }
            __syncwarp();
            warp_zgemm9<1>(S2, S0, S1, lane, S2, S3);          // A9 = B4 + B1 B5 -> S2 (in place); B3 + A9 -> S3
            __syncwarp();
            warp_zgemm9<2>(S0, S3, S2, lane, S4);              // T18 = B2 + (B3 + A9) A9 -> S0
            __syncwarp();
            cplx* X = S0;
            for (int i = 0; i < s; ++i) {                      // undo the scaling
                cplx* nxt = (X == S0) ? S1 : S0;
                warp_zgemm9<0>(nxt, X, X, lane);
                __syncwarp();
                X = nxt;
            }
            if (p.dUs_out != nullptr) {
                const cplx ph_n = shifted ? cexp_(mu) : cmake(1.0, 0.0);
                cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d;
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int e = lane + 32 * u, i = e / 9, j = e - 9 * i;
                    if (e < 81 && i < d && j < d) o[i * d + j] = cmul(ph_n, X[eidx[u]]);
                }
            }
            // ---- running product: P <- dU_n P ----------------------------------------------------------------------------------
            if (n == n_begin) {
#pragma unroll
                for (int u = 0; u < 3; ++u)
                    if (evalid[u]) P[eidx[u]] = X[eidx[u]];
                __syncwarp();
            } else {
                warp_zgemm9<0>(S4, X, P, lane);                // B2 is dead: its slot takes dU_n P
                __syncwarp();
                cplx* t = P; P = S4; S4 = t;
            }
        }
        // ---- write the segment product with its accumulated shift ---------------------------------------------------------------
        const cplx ph_u = shifted ? cexp_(mu_acc) : cmake(1.0, 0.0);
        cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = lane + 32 * u, i = e / 9, j = e - 9 * i;
            if (e < 81 && i < d && j < d) o[i * d + j] = cmul(ph_u, P[eidx[u]]);
        }
        __syncwarp();
    }
}

}  // namespace c3b
