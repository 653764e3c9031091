// Register-resident PWC propagator kernel for small Hilbert dimensions (D <= 12).
//
// Replaces, for one whole batch of control signals, the reference chain
//   tf_batch_propagate -> tf_propagation_vectorized -> tf.linalg.expm   (c3/libraries/propagation.py:460-515, 426-440)
//   tf_matmul_n / tf_matmul_left                                        (c3/utils/tf_utils.py:120-193)
// with ONE fused kernel: assemble A_n = G0 + sum_k c_k[n] G_k, exponentiate by
// scaling-and-squaring Pade, and fold the ordered product U = dU_{N-1} ... dU_0 on chip.
//
// Mapping.  A "group" of D lanes owns one time slice at a time; lane r holds ROW r of every
// matrix in registers (D complex = 4D 32-bit registers per matrix row).  floor(32/D) groups
// share a warp and walk consecutive sub-chunks of the warp's time segment.  A product
// C = X * Y takes X's row from registers and streams Y from a per-group shared-memory
// buffer with broadcast LDS.128 (all lanes of a group read the same address), so each
// 16-byte shared load feeds 4 DFMAs and no shuffles are needed.  Warps never synchronise
// with each other (only __syncwarp), so the fp64 pipe is kept busy by 8-12 independent
// warps per SM.
//
// Exponential.  Pade order m in {3,5,7,9} and squarings s are chosen per warp from an
// inf-norm bound  ||A_n|| <= rs0[r] + sum_k |c_k| rs_k[r]  (no square roots on the hot
// path).  s is taken so that the scaled norm is below 2 ln 2, which makes V-U strictly
// diagonally dominant and lets the linear solve be an un-pivoted Gauss-Jordan sweep with
// no divergence (see c3b_common.cuh).  For every BASELINE config this selects exactly
// Higham's minimal (m, s); the accuracy is that of Higham 2005 for any input.
#pragma once
#include "c3b_common.cuh"

namespace c3b {

struct RowsParams {
    const cplx* G;          // [(Bm), K+1, d, d]  pre-scaled generators: A = G0 + sum_k c_k G_k
    const double* RS;       // [(Bm), K+1, d]     row sums of |G_k| (inf-norm bound pieces)
    const cplx* TR;         // [(Bm), K+1] trace shifts t_k = tr(G_k)/d already SUBTRACTED from G_k's diagonal, or null
                            //   (exp(A) = exp(mu) exp(A - mu I), mu_n = t_0 + sum_k c_k[n] t_k; only kernels that
                            //    re-apply exp(mu) accept shifted generators)
    const double* signals;  // [B, K, N] real control fields, contiguous in N (may be null if K == 0)
    const cplx* hlist;      // [B, N, d, d] explicit Hamiltonians (H-list mode) or null
    double hscale_re, hscale_im;  // H-list mode: A = hscale * H   (-i dt for the closed system)
    long long model_stride;       // elements between consecutive batch models in G (0 = shared)
    int B, K, N, d;               // d = actual dimension (<= template D)
    int S;                        // segments per batch element
    int seg_len;                  // slices per segment
    cplx* U_out;                  // [B, d, d]      (used when S == 1)
    cplx* seg_out;                // [B, S, d, d]   (used when S > 1)
    cplx* dUs_out;                // [B, N, d, d] or null
    const unsigned int* rows_ready;   // gated launch (host-resident signals): batch rows [0, *rows_ready) have landed in `signals`; or null
};

// Gated launch: wait until batch row b of the control fields has arrived (warp-uniform: every lane polls).  The host enqueues the
// chunked host->device copies of `signals` on a copy stream, each followed by a 4-byte copy that raises *rows_ready, and
// launches ONE persistent kernel; warps pull units in batch order and spin here only if they overtake the copy engine.
// Returns false after ~4 s without progress.
__device__ __forceinline__ bool wait_rows_ready(const unsigned int* rows_ready, const int b) {
    unsigned int spins = 0, v;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(rows_ready) : "memory");
        if (v > (unsigned int)b) return true;
        __nanosleep(256);
        if (++spins > (1u << 24)) return false;
    }
}
// control-field sample.  Also in a gated launch the read-only path is safe: no thread touches a row's addresses before
// the acquire load of *rows_ready has seen it (the asm's memory clobber keeps the compiler from hoisting), and L1 holds
// no lines of this buffer from before the launch.
__device__ __forceinline__ double load_signal(const double* ptr, const bool gated) { (void)gated; return __ldg(ptr); }

// C(row) = X(row) * Y, Y a D x D matrix in shared memory (row-major, broadcast reads).
template <int D>
__device__ __forceinline__ void mm_row(const cplx (&x)[D], const cplx* __restrict__ Y, cplx (&c)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) c[j] = cmake(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const cplx xk = x[k];
#pragma unroll
        for (int j = 0; j < D; ++j) cfma(c[j], xk, Y[k * D + j]);
    }
}

template <int D>
__device__ __forceinline__ void store_row(cplx* __restrict__ M, int r, const cplx (&x)[D], bool pred) {
    if (pred) {
#pragma unroll
        for (int j = 0; j < D; ++j) M[r * D + j] = x[j];
    }
}

template <int D, int WARPS>
struct RowsLayout {
    static constexpr int G = 32 / D;                   // groups per warp
    static constexpr int GROUP_ELEMS = 3 * D * D + 4 * D;  // bufA, bufA2, bufP, piv[2][2D]  (cplx units)
    static constexpr int WARP_ELEMS = G * GROUP_ELEMS;
    __host__ __device__ static size_t smem_bytes(int K) {
        size_t model = (size_t)(K + 1) * D * D * sizeof(cplx) + (size_t)(((K + 1) * D + 1) & ~1) * sizeof(double);
        return model + (size_t)WARPS * WARP_ELEMS * sizeof(cplx);
    }
};

template <int D, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) pwc_rows_kernel(const RowsParams p) {
    using L = RowsLayout<D, WARPS>;
    constexpr int G = L::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);                    // [(K+1), D, D] zero padded
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * D * D);   // [(K+1), D]
    cplx* sWarps = reinterpret_cast<cplx*>(sRS + (((K + 1) * D + 1) & ~1));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool hmode = p.hlist != nullptr;

    // ---- stage the (shared) model once per CTA -------------------------------------------
    // TODO(per-batch models): model_stride != 0 is routed to the CTA kernel by the host.
    if (!hmode) {
        for (int idx = tid; idx < (K + 1) * D * D; idx += WARPS * 32) {
            const int k = idx / (D * D);
            const int rem = idx - k * D * D;
            const int r = rem / D, j = rem - r * D;
            cplx v = cmake(0.0, 0.0);
            if (r < d && j < d) v = p.G[(size_t)k * d * d + r * d + j];
            sG[idx] = v;
        }
        for (int idx = tid; idx < (K + 1) * D; idx += WARPS * 32) {
            const int k = idx / D, r = idx - k * D;
            sRS[idx] = (r < d) ? p.RS[k * d + r] : 0.0;
        }
    }
    __syncthreads();

    const long long unit = (long long)blockIdx.x * WARPS + warp;
    if (unit >= (long long)p.B * p.S) return;  // whole warp leaves; no CTA barrier after this point
    const int b = (int)(unit / p.S);
    const int sidx = (int)(unit - (long long)b * p.S);

    const int g_raw = lane / D;
    const bool lane_on = g_raw < G;           // lanes beyond G*D idle (they shadow group 0)
    const int g = lane_on ? g_raw : 0;
    const int r = lane_on ? (lane - g_raw * D) : 0;

    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + (size_t)g * L::GROUP_ELEMS;
    cplx* bufA = gbase;
    cplx* bufA2 = gbase + D * D;
    cplx* bufP = gbase + 2 * D * D;
    cplx* piv = gbase + 3 * D * D;

    const int n_begin = sidx * p.seg_len;
    const int n_end = min(p.N, n_begin + p.seg_len);
    const int len = n_end - n_begin;
    const int cl = (len + G - 1) / G;  // slices per group (last group may get fewer)
    const int my_begin = n_begin + g * cl;
    const int my_end = min(n_end, my_begin + cl);

    const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
    const cplx hs = cmake(p.hscale_re, p.hscale_im);

    for (int it = 0; it < cl; ++it) {
        const int n = my_begin + it;
        const bool on = lane_on && (n < my_end);

        // ---- assemble row r of A_n and an inf-norm bound ---------------------------------
        cplx A[D];
        double nb = 0.0;
        if (!hmode) {
#pragma unroll
            for (int j = 0; j < D; ++j) A[j] = on ? sG[r * D + j] : cmake(0.0, 0.0);
            nb = on ? sRS[r] : 0.0;
            for (int k = 0; k < K; ++k) {
                const double c = on ? __ldg(sig_b + (size_t)k * p.N + n) : 0.0;
                const cplx* gk = sG + (k + 1) * D * D + r * D;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const cplx gv = gk[j];
                    A[j].x = fma(c, gv.x, A[j].x);
                    A[j].y = fma(c, gv.y, A[j].y);
                }
                nb = fma(fabs(c), sRS[(k + 1) * D + r], nb);
            }
        } else {
            const cplx* hrow = p.hlist + ((size_t)b * p.N + (on ? n : 0)) * d * d + (size_t)r * d;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                cplx h = cmake(0.0, 0.0);
                if (on && r < d && j < d) h = hrow[j];
                A[j] = cmul(hs, h);
                nb += cabs1(A[j]);
            }
        }
        nb = warp_max(nb);  // warp-uniform => no divergence in what follows

        const int s = squarings_for(nb, C3B_NOPIVOT_LIMIT);
        const double ns = nb * pow2neg(s);
        const int mi = ns < C3B_THETA3 ? 0 : (ns < C3B_THETA5 ? 1 : (ns < C3B_THETA7 ? 2 : 3));  // m = 2 mi + 3
        if (s > 0) {
            const double sc = pow2neg(s);
#pragma unroll
            for (int j = 0; j < D; ++j) { A[j].x *= sc; A[j].y *= sc; }
        }

        // ---- Pade numerator parts: U = A * W (odd), V (even) -----------------------------
        store_row<D>(bufA, r, A, lane_on);
        __syncwarp();
        cplx X[D], W[D], V[D];
        mm_row<D>(A, bufA, X);  // X = A^2 (row r)
        {
            const double c1 = kPade[mi][1], c3 = kPade[mi][3], c0 = kPade[mi][0], c2 = kPade[mi][2];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                W[j] = cmake(c3 * X[j].x, c3 * X[j].y);
                V[j] = cmake(c2 * X[j].x, c2 * X[j].y);
            }
            // identity contributions on the diagonal element of this row
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j == r) { W[j].x += c1; V[j].x += c0; }
        }
        if (mi > 0) {
            store_row<D>(bufA2, r, X, lane_on);
            __syncwarp();
            for (int i = 0; i < mi; ++i) {
                cplx Xn[D];
                mm_row<D>(X, bufA2, Xn);  // A^(2i+4)
                const double cw = kPade[mi][2 * i + 5], cv = kPade[mi][2 * i + 4];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    W[j].x = fma(cw, Xn[j].x, W[j].x);
                    W[j].y = fma(cw, Xn[j].y, W[j].y);
                    V[j].x = fma(cv, Xn[j].x, V[j].x);
                    V[j].y = fma(cv, Xn[j].y, V[j].y);
                    X[j] = Xn[j];
                }
            }
        }
        cplx R[D];
        mm_row<D>(W, bufA, R);  // U = W * A  (polynomials in A commute)
        cplx Q[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            Q[j] = cmake(V[j].x - R[j].x, V[j].y - R[j].y);
            R[j] = cmake(V[j].x + R[j].x, V[j].y + R[j].y);
        }

        // ---- R <- Q^{-1} R by un-pivoted Gauss-Jordan; pivot rows broadcast through smem ---
#pragma unroll
        for (int k = 0; k < D; ++k) {
            cplx* pv = piv + (k & 1) * 2 * D;
            if (lane_on && r == k) {
#pragma unroll
                for (int j = k; j < D; ++j) pv[j] = Q[j];
#pragma unroll
                for (int j = 0; j < D; ++j) pv[D + j] = R[j];
            }
            __syncwarp();
            const cplx inv = crcp(pv[k]);
            cplx f = cmul(Q[k], inv);
            if (r == k) f = cmake(1.0 - inv.x, -inv.y);  // owner: row <- row * inv == row - (1-inv) row
#pragma unroll
            for (int j = k + 1; j < D; ++j) cfms(Q[j], f, pv[j]);
#pragma unroll
            for (int j = 0; j < D; ++j) cfms(R[j], f, pv[D + j]);
        }

        // ---- undo the scaling: s squarings -------------------------------------------------
        for (int i = 0; i < s; ++i) {
            store_row<D>(bufA, r, R, lane_on);
            __syncwarp();
            cplx T[D];
            mm_row<D>(R, bufA, T);
#pragma unroll
            for (int j = 0; j < D; ++j) R[j] = T[j];
            __syncwarp();
        }

        // ---- optional partial propagators ---------------------------------------------------
        if (p.dUs_out != nullptr && on && r < d) {
            cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)r * d;
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j < d) o[j] = R[j];
        }

        // ---- running ordered product P <- dU_n * P -------------------------------------------
        if (it == 0) {
            store_row<D>(bufP, r, R, lane_on);
        } else {
            cplx T[D];
            mm_row<D>(R, bufP, T);
            __syncwarp();
            store_row<D>(bufP, r, T, lane_on);
        }
    }
    __syncwarp();

    // ---- fold the G group products of this warp: P_{G-1} ... P_1 P_0 ------------------------
    // (every group computes the same thing redundantly; lanes of group 0 write it out)
    cplx* wbase = sWarps + (size_t)warp * L::WARP_ELEMS;
    cplx T[D];
    {
        const cplx* last = wbase + (size_t)(G - 1) * L::GROUP_ELEMS + 2 * D * D;
#pragma unroll
        for (int j = 0; j < D; ++j) T[j] = last[r * D + j];
    }
#pragma unroll 1
    for (int gg = G - 2; gg >= 0; --gg) {
        cplx T2[D];
        mm_row<D>(T, wbase + (size_t)gg * L::GROUP_ELEMS + 2 * D * D, T2);
#pragma unroll
        for (int j = 0; j < D; ++j) T[j] = T2[j];
    }
    if (lane_on && g == 0 && r < d) {
        cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
        for (int j = 0; j < D; ++j)
            if (j < d) o[r * d + j] = T[j];
    }
}

// =============================================================================================
// v2: the same algorithm as pwc_rows_kernel, restructured as a small state machine around ONE
// instance of the row-times-matrix product so that the slice loop fits the instruction cache
// and the live register set is {left operand row, product row, W row (+ V row)}.
//   phase 0..mi      : C = A^(2ph+2); W += c_odd C, V += c_even C
//   phase mi+1       : C = U = W A;   Q = V - U, R = V + U;  R <- Q^{-1} R  (Gauss-Jordan)
//   next s phases    : C = R R  (squarings)
//   last phase       : C = dU P  (running ordered product; skipped for the first slice)
// VSMEM keeps the V row in shared memory (own row only) instead of registers.
// =============================================================================================
template <int D, int WARPS, bool VSMEM>
struct Rows2Layout {
    static constexpr int G = 32 / D;
    static constexpr int GROUP_ELEMS = (VSMEM ? 4 : 3) * D * D + 4;  // bufA, bufA2, bufP, [bufV] (+4: bank offset between groups)
    static constexpr int WARP_ELEMS = G * GROUP_ELEMS;
    __host__ __device__ static size_t smem_bytes(int K) {
        size_t model = (size_t)(K + 1) * D * D * sizeof(cplx) + (size_t)(((K + 1) * D + 1) & ~1) * sizeof(double);
        return model + (size_t)WARPS * WARP_ELEMS * sizeof(cplx);
    }
};

template <int D, int WARPS, int MINB, bool VSMEM>
__global__ void __launch_bounds__(WARPS * 32, MINB) pwc_rows2_kernel(const RowsParams p) {
    using L = Rows2Layout<D, WARPS, VSMEM>;
    constexpr int G = L::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K;
    const int d = p.d;
    cplx* sG = reinterpret_cast<cplx*>(smem_raw);
    double* sRS = reinterpret_cast<double*>(sG + (K + 1) * D * D);
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
            sG[idx] = v;
        }
        for (int idx = tid; idx < (K + 1) * D; idx += WARPS * 32) {
            const int k = idx / D, r = idx - k * D;
            sRS[idx] = (r < d) ? p.RS[k * d + r] : 0.0;
        }
    }
    __syncthreads();

    const long long unit = (long long)blockIdx.x * WARPS + warp;
    if (unit >= (long long)p.B * p.S) return;
    const int b = (int)(unit / p.S);
    const int sidx = (int)(unit - (long long)b * p.S);

    const int g_raw = lane / D;
    const bool lane_on = g_raw < G;
    const int g = lane_on ? g_raw : 0;
    const int r = lane_on ? (lane - g_raw * D) : 0;

    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + (size_t)g * L::GROUP_ELEMS;
    cplx* bufA = gbase;
    cplx* bufA2 = gbase + D * D;
    cplx* bufP = gbase + 2 * D * D;
    cplx* bufV = gbase + 3 * D * D + r * D;              // own row (VSMEM only)
    const int gbase_lane = g * D;                        // first lane of this group

    const int n_begin = sidx * p.seg_len;
    const int n_end = min(p.N, n_begin + p.seg_len);
    const int len = n_end - n_begin;
    const int cl = (len + G - 1) / G;
    const int my_begin = n_begin + g * cl;
    const int my_end = min(n_end, my_begin + cl);

    const double* sig_b = p.signals ? p.signals + (size_t)b * K * p.N : nullptr;
    const cplx hs = cmake(p.hscale_re, p.hscale_im);

#pragma unroll 1
    for (int it = 0; it < cl; ++it) {
        const int n = my_begin + it;
        const bool on = lane_on && (n < my_end);

        cplx Xop[D];
        double nb = 0.0;
        if (!hmode) {
#pragma unroll
            for (int j = 0; j < D; ++j) Xop[j] = on ? sG[r * D + j] : cmake(0.0, 0.0);
            nb = on ? sRS[r] : 0.0;
            for (int k = 0; k < K; ++k) {
                const double c = on ? __ldg(sig_b + (size_t)k * p.N + n) : 0.0;
                const cplx* gk = sG + (k + 1) * D * D + r * D;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const cplx gv = gk[j];
                    Xop[j].x = fma(c, gv.x, Xop[j].x);
                    Xop[j].y = fma(c, gv.y, Xop[j].y);
                }
                nb = fma(fabs(c), sRS[(k + 1) * D + r], nb);
            }
        } else {
            const cplx* hrow = p.hlist + ((size_t)b * p.N + (on ? n : 0)) * d * d + (size_t)r * d;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                cplx h = cmake(0.0, 0.0);
                if (on && r < d && j < d) h = hrow[j];
                Xop[j] = cmul(hs, h);
                nb += cabs1(Xop[j]);
            }
        }
        nb = warp_max(nb);

        const int s = squarings_for(nb, C3B_NOPIVOT_LIMIT);
        const double ns = nb * pow2neg(s);
        const int mi = ns < C3B_THETA3 ? 0 : (ns < C3B_THETA5 ? 1 : (ns < C3B_THETA7 ? 2 : 3));
        if (s > 0) {
            const double sc = pow2neg(s);
#pragma unroll
            for (int j = 0; j < D; ++j) { Xop[j].x *= sc; Xop[j].y *= sc; }
        }
        store_row<D>(bufA, r, Xop, lane_on);
        __syncwarp();

        const cplx* Y = bufA;
        cplx W[D];
        cplx V[VSMEM ? 1 : D];
        const int ph_solve = mi + 1;
        const int ph_lastsq = mi + 1 + s;
        const int ph_last = ph_lastsq + (it > 0 ? 1 : 0);
        const double* cf = kPade[mi];

#pragma unroll 1
        for (int ph = 0; ph <= ph_last; ++ph) {
            cplx C[D];
            mm_row<D>(Xop, Y, C);
            if (ph < ph_solve) {
                const double cw = cf[2 * ph + 3], cv = cf[2 * ph + 2];
                if (ph == 0) {
                    const double c1 = cf[1], c0 = cf[0];
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        W[j] = cmake(cw * C[j].x + (j == r ? c1 : 0.0), cw * C[j].y);
                        const cplx v = cmake(cv * C[j].x + (j == r ? c0 : 0.0), cv * C[j].y);
                        if (VSMEM) { if (lane_on) bufV[j] = v; } else V[VSMEM ? 0 : j] = v;
                    }
                    if (mi > 0) {
                        store_row<D>(bufA2, r, C, lane_on);
                        __syncwarp();
                        Y = bufA2;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        W[j].x = fma(cw, C[j].x, W[j].x);
                        W[j].y = fma(cw, C[j].y, W[j].y);
                        if (VSMEM) {
                            cplx v = bufV[j];
                            v.x = fma(cv, C[j].x, v.x);
                            v.y = fma(cv, C[j].y, v.y);
                            if (lane_on) bufV[j] = v;
                        } else {
                            V[VSMEM ? 0 : j].x = fma(cv, C[j].x, V[VSMEM ? 0 : j].x);
                            V[VSMEM ? 0 : j].y = fma(cv, C[j].y, V[VSMEM ? 0 : j].y);
                        }
                    }
                }
                if (ph < mi) {
#pragma unroll
                    for (int j = 0; j < D; ++j) Xop[j] = C[j];
                } else {
#pragma unroll
                    for (int j = 0; j < D; ++j) Xop[j] = W[j];
                    Y = bufA;
                }
            } else if (ph == ph_solve) {
                // C = U.  W <- Q = V - U,  Xop <- R = V + U, then R <- Q^{-1} R
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const cplx v = VSMEM ? bufV[j] : V[VSMEM ? 0 : j];
                    W[j] = cmake(v.x - C[j].x, v.y - C[j].y);
                    Xop[j] = cmake(v.x + C[j].x, v.y + C[j].y);
                }
                // Pivot row k lives in the registers of lane (group base + k): broadcast it with
                // warp shuffles (4 SHFL.32 per complex = 4 MIO wavefronts) instead of a
                // predicated STS + broadcast LDS round trip (~7.5 wavefronts per complex).
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int src = gbase_lane + k;
                    cplx pk;
                    pk.x = __shfl_sync(0xffffffffu, W[k].x, src);
                    pk.y = __shfl_sync(0xffffffffu, W[k].y, src);
                    const cplx inv = crcp(pk);
                    cplx f = cmul(W[k], inv);
                    if (r == k) f = cmake(1.0 - inv.x, -inv.y);
#pragma unroll
                    for (int j = k + 1; j < D; ++j) {
                        cplx pj;
                        pj.x = __shfl_sync(0xffffffffu, W[j].x, src);
                        pj.y = __shfl_sync(0xffffffffu, W[j].y, src);
                        cfms(W[j], f, pj);
                    }
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        cplx pj;
                        pj.x = __shfl_sync(0xffffffffu, Xop[j].x, src);
                        pj.y = __shfl_sync(0xffffffffu, Xop[j].y, src);
                        cfms(Xop[j], f, pj);
                    }
                }
                if (s > 0) {
                    store_row<D>(bufA, r, Xop, lane_on);
                    __syncwarp();
                    Y = bufA;
                } else {
                    Y = bufP;
                }
            } else if (ph <= ph_lastsq) {
#pragma unroll
                for (int j = 0; j < D; ++j) Xop[j] = C[j];
                if (ph < ph_lastsq) {
                    __syncwarp();
                    store_row<D>(bufA, r, Xop, lane_on);
                    __syncwarp();
                    Y = bufA;
                } else {
                    Y = bufP;
                }
            } else {
                // C = dU_n * P
                __syncwarp();
                store_row<D>(bufP, r, C, lane_on);
            }
            if (ph == ph_lastsq) {
                // Xop holds dU_n
                if (p.dUs_out != nullptr && on && r < d) {
                    cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)r * d;
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j < d) o[j] = Xop[j];
                }
                if (it == 0) store_row<D>(bufP, r, Xop, lane_on);
            }
        }
    }
    __syncwarp();

    cplx* wbase = sWarps + (size_t)warp * L::WARP_ELEMS;
    cplx T[D];
    {
        const cplx* last = wbase + (size_t)(G - 1) * L::GROUP_ELEMS + 2 * D * D;
#pragma unroll
        for (int j = 0; j < D; ++j) T[j] = last[r * D + j];
    }
#pragma unroll 1
    for (int gg = G - 2; gg >= 0; --gg) {
        cplx T2[D];
        mm_row<D>(T, wbase + (size_t)gg * L::GROUP_ELEMS + 2 * D * D, T2);
#pragma unroll
        for (int j = 0; j < D; ++j) T[j] = T2[j];
    }
    if (lane_on && g == 0 && r < d) {
        cplx* o = (p.S == 1) ? (p.U_out + (size_t)b * d * d) : (p.seg_out + ((size_t)b * p.S + sidx) * d * d);
#pragma unroll
        for (int j = 0; j < D; ++j)
            if (j < d) o[r * d + j] = T[j];
    }
}

// =============================================================================================
// v3: R rows per lane.  The shared-memory/shuffle data path (MIO) moves 128 lane-bytes per
// clock per SM while the fp64 pipe retires 64 lane-DFMA per clock, i.e. one complex
// multiply-accumulate (4 DFMA) per lane costs half of what one 16-byte operand load costs.
// With one row per lane (v1/v2) every loaded element of Y feeds ONE cfma and the kernel is
// MIO-bound at ~43% fp64-pipe utilisation (ncu: l1tex lsu wavefronts 96% of peak).  Here a
// lane owns R rows of every matrix (d=9: R=3 -> 3 lanes per matrix, 10 matrices per warp), so
// each loaded element feeds R cfma.  Register budget: left operand (R*D cplx) + accumulators
// (R*D cplx); the Pade numerator parts W and V are therefore NOT accumulated in registers but
// formed once, after the last power, from the highest power (registers) and the lower powers
// re-read from the lane's own rows in shared memory.  Orders m in {3,5,7} (+ squarings); for
// 0.95 <= ||A|| < 2 ln 2 this is m=7 with one squaring: the same 5 products as Higham's m=9.
// Warps are persistent and pull (batch, segment) units from an atomic counter.
// =============================================================================================
template <int D, int R>
struct Rows3Layout {
    static constexpr int LPM = (D + R - 1) / R;          // lanes per matrix
    static constexpr int MPW = 32 / LPM;                 // matrices (lane groups) per warp
    static constexpr int BUF = D * D;
    static constexpr int GROUP_ELEMS = 4 * BUF + 1;      // bufA, bufA2, bufT, bufP (+1: odd 16-byte stride)
    static constexpr int WARP_ELEMS = MPW * GROUP_ELEMS;
    __host__ __device__ static size_t smem_bytes(int K, int warps) {
        size_t model = (size_t)(K + 1) * LPM * R * D * sizeof(cplx) + (size_t)(((K + 1) * LPM * R + 1) & ~1) * sizeof(double);
        return model + (size_t)warps * WARP_ELEMS * sizeof(cplx);
    }
};

template <int D, int R>
__device__ __forceinline__ void mm_rows(const cplx (&x)[R][D], const cplx* __restrict__ Y, cplx (&c)[R][D]) {
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) c[a][j] = cmake(0.0, 0.0);
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

// store the lane's R rows (rows >= D are padding and never stored)
template <int D, int R>
__device__ __forceinline__ void store_rows(cplx* __restrict__ M, int row0, const cplx (&x)[R][D], bool pred) {
#pragma unroll
    for (int a = 0; a < R; ++a) {
        if (pred && row0 + a < D) {
#pragma unroll
            for (int j = 0; j < D; ++j) M[(row0 + a) * D + j] = x[a][j];
        }
    }
}

template <int D, int R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pwc_rows3_kernel(const RowsParams p, unsigned int* __restrict__ counter) {
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
    const int gbase_lane = g * LPM;

    cplx* gbase = sWarps + (size_t)warp * L::WARP_ELEMS + (size_t)g * L::GROUP_ELEMS;
    cplx* bufA = gbase;
    cplx* bufA2 = gbase + L::BUF;
    cplx* bufT = gbase + 2 * L::BUF;
    cplx* bufP = gbase + 3 * L::BUF;
    const cplx hs = cmake(p.hscale_re, p.hscale_im);
    const long long total_units = (long long)p.B * p.S;

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

#pragma unroll 1
        for (int it = 0; it < cl; ++it) {
            const int n = my_begin + it;
            const bool on = lane_on && (n < my_end);

            // ---- assemble this lane's rows of A_n and the inf-norm bound ---------------------
            cplx Xop[R][D];
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

            // order m = 2 mi + 3 in {3,5,7}; s squarings so that the scaled norm is < theta_7 and < 2 ln 2
            const int s = squarings_for(nb, C3B_THETA7);
            const double ns = nb * pow2neg(s);
            const int mi = ns < C3B_THETA3 ? 0 : (ns < C3B_THETA5 ? 1 : 2);
            if (s > 0) {
                const double sc = pow2neg(s);
#pragma unroll
                for (int a = 0; a < R; ++a)
#pragma unroll
                    for (int j = 0; j < D; ++j) { Xop[a][j].x *= sc; Xop[a][j].y *= sc; }
            }
            store_rows<D, R>(bufA, row0, Xop, lane_on);
            __syncwarp();

            const cplx* Y = bufA;
            const double* cf = kPade[mi];
            const int ph_solve = mi + 1;
            const int ph_lastsq = mi + 1 + s;
            const int ph_last = ph_lastsq + (it > 0 ? 1 : 0);

#pragma unroll 1
            for (int ph = 0; ph <= ph_last; ++ph) {
                cplx C[R][D];
                mm_rows<D, R>(Xop, Y, C);
                if (ph < ph_solve) {
                    if (ph == 0 && mi > 0) {           // A^2: operand of the following powers
                        store_rows<D, R>(bufA2, row0, C, lane_on);
                        __syncwarp();
                        Y = bufA2;
                    }
                    if (ph < mi) {
                        if (ph == 1) store_rows<D, R>(bufT, row0, C, lane_on);   // A^4 (m = 7): own rows, re-read below
#pragma unroll
                        for (int a = 0; a < R; ++a)
#pragma unroll
                            for (int j = 0; j < D; ++j) Xop[a][j] = C[a][j];
                    } else {
                        // last power: W (odd coefficients) -> Xop, V (even coefficients) -> bufT (own rows)
                        const double cwh = cf[2 * mi + 3], cvh = cf[2 * mi + 2];
                        const double c3 = cf[3], c2 = cf[2], c5 = cf[5], c4 = cf[4], c1 = cf[1], c0 = cf[0];
#pragma unroll
                        for (int a = 0; a < R; ++a) {
                            const int row = row0 + a;
                            const int rr = row < D ? row : 0;
#pragma unroll
                            for (int j = 0; j < D; ++j) {
                                const cplx pw = C[a][j];
                                cplx w = cmake(cwh * pw.x, cwh * pw.y);
                                cplx v = cmake(cvh * pw.x, cvh * pw.y);
                                if (mi >= 1) {
                                    const cplx a2 = bufA2[rr * D + j];
                                    w.x = fma(c3, a2.x, w.x); w.y = fma(c3, a2.y, w.y);
                                    v.x = fma(c2, a2.x, v.x); v.y = fma(c2, a2.y, v.y);
                                }
                                if (mi == 2) {
                                    const cplx a4 = bufT[rr * D + j];
                                    w.x = fma(c5, a4.x, w.x); w.y = fma(c5, a4.y, w.y);
                                    v.x = fma(c4, a4.x, v.x); v.y = fma(c4, a4.y, v.y);
                                }
                                if (j == row) { w.x += c1; v.x += c0; }
                                Xop[a][j] = w;
                                if (lane_on && row < D) bufT[row * D + j] = v;
                            }
                        }
                        Y = bufA;
                    }
                } else if (ph == ph_solve) {
                    // C = U.  Xop <- Q = V - U,  C <- R = V + U, then C <- Q^{-1} C (Gauss-Jordan,
                    // pivot rows broadcast from the owning lane's registers by warp shuffles)
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        const int rr = (row0 + a) < D ? (row0 + a) : 0;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            cplx v = bufT[rr * D + j];
                            if (row0 + a >= D) v = cmake(0.0, 0.0);
                            const cplx u = C[a][j];
                            Xop[a][j] = cmake(v.x - u.x, v.y - u.y);
                            C[a][j] = cmake(v.x + u.x, v.y + u.y);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        constexpr int dummy = 0; (void)dummy;
                        const int owner = k / R;          // compile-time after unrolling
                        const int a0 = k % R;
                        const int src = gbase_lane + owner;
                        cplx pk;
                        pk.x = __shfl_sync(0xffffffffu, Xop[a0][k].x, src);
                        pk.y = __shfl_sync(0xffffffffu, Xop[a0][k].y, src);
                        const cplx inv = crcp(pk);
                        cplx f[R];
#pragma unroll
                        for (int a = 0; a < R; ++a) f[a] = cmul(Xop[a][k], inv);
                        if (l == owner) f[a0] = cmake(1.0 - inv.x, -inv.y);
#pragma unroll
                        for (int j = k + 1; j < D; ++j) {
                            cplx pj;
                            pj.x = __shfl_sync(0xffffffffu, Xop[a0][j].x, src);
                            pj.y = __shfl_sync(0xffffffffu, Xop[a0][j].y, src);
#pragma unroll
                            for (int a = 0; a < R; ++a) cfms(Xop[a][j], f[a], pj);
                        }
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            cplx pj;
                            pj.x = __shfl_sync(0xffffffffu, C[a0][j].x, src);
                            pj.y = __shfl_sync(0xffffffffu, C[a0][j].y, src);
#pragma unroll
                            for (int a = 0; a < R; ++a) cfms(C[a][j], f[a], pj);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < R; ++a)
#pragma unroll
                        for (int j = 0; j < D; ++j) Xop[a][j] = C[a][j];
                    if (s > 0) {
                        store_rows<D, R>(bufT, row0, Xop, lane_on);   // V is dead: bufT becomes the squaring operand
                        __syncwarp();
                        Y = bufT;
                    } else {
                        Y = bufP;
                    }
                } else if (ph <= ph_lastsq) {
#pragma unroll
                    for (int a = 0; a < R; ++a)
#pragma unroll
                        for (int j = 0; j < D; ++j) Xop[a][j] = C[a][j];
                    if (ph < ph_lastsq) {
                        // ping-pong between bufT and bufA (A is dead after U)
                        cplx* nxt = (Y == bufT) ? bufA : bufT;
                        store_rows<D, R>(nxt, row0, Xop, lane_on);
                        __syncwarp();
                        Y = nxt;
                    } else {
                        Y = bufP;
                    }
                } else {
                    __syncwarp();                       // everyone has finished reading the old P
                    store_rows<D, R>(bufP, row0, C, lane_on);
                }
                if (ph == ph_lastsq) {                  // Xop holds dU_n
                    if (p.dUs_out != nullptr && on) {
#pragma unroll
                        for (int a = 0; a < R; ++a) {
                            const int row = row0 + a;
                            if (row < d) {
                                cplx* o = p.dUs_out + ((size_t)b * p.N + n) * d * d + (size_t)row * d;
#pragma unroll
                                for (int j = 0; j < D; ++j)
                                    if (j < d) o[j] = Xop[a][j];
                            }
                        }
                    }
                    if (it == 0) store_rows<D, R>(bufP, row0, Xop, lane_on);
                }
            }
        }
        __syncwarp();

        // ---- fold the MPW group products of this warp (pairwise tree, later chunks on the left) --
        // level with stride st: group g computes M[hi] * M[lo] with lo = g rounded down to a
        // multiple of 2 st, hi = lo + st (if it exists); every group stores into its own bufT/bufA.
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
                    for (int j = 0; j < D; ++j) Xh[a][j] = (row0 + a) < D ? Mh[rr * D + j] : cmake(0.0, 0.0);
                }
                mm_rows<D, R>(Xh, wbase + (size_t)lo * L::GROUP_ELEMS + cur * L::BUF, T);
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
