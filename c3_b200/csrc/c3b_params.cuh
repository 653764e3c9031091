// Parameter blocks of every PWC / product kernel, and the gated-launch helpers shared by the small-dimension (lane-group) PWC kernels
// (pwc_blk.cuh, pwc_blk9.cuh, pwc_shfl9.cuh).  Contract of those kernels: for one whole batch of control signals,
//   tf_batch_propagate -> tf_propagation_vectorized -> tf.linalg.expm   (c3/libraries/propagation.py:460-515, 426-440)
//   tf_matmul_n / tf_matmul_left                                        (c3/utils/tf_utils.py:120-193)
// as ONE fused kernel: assemble A_n = G0 + sum_k c_k[n] G_k, exponentiate, fold the ordered product on chip.
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
    // gated launch (host-resident signals): gate[0] = number of batch rows [0, gate[0]) that have landed in `signals`
    // (raised by the host's copy stream), gate[1] = set to 1 by the kernel if a row did not arrive in time; or null
    unsigned int* gate;
    int skew;                     // shuffle kernel: clocks by which the late half of a CTA's warps trails the early half
};

struct CtaParams {
    const cplx* G;          // [(Bm), K+1, D, D] pre-scaled generators
    const double* signals;  // [B, K, N]
    const cplx* hlist;      // [B, N, D, D] or null
    double hscale_re, hscale_im;
    long long model_stride;  // elements between batch models in G (0 = shared)
    int B, K, N, D;
    int S, seg_len;
    cplx* U_out;    // [B, D, D]
    cplx* seg_out;  // [B, S, D, D]
    cplx* dUs_out;  // [B, N, D, D] or null
    cplx* ws;       // global workspace, gridDim.x * kCtaSlots * D * D (only when matrices do not fit smem)
    int use_smem;   // 1: matrices in dynamic shared memory
    const cplx* unshift;  // literal Higham kernel only: [(Bm), K+1] trace shifts to put back on the diagonal, or null
};

constexpr int kCtaThreads = 256;
constexpr int kCtaSlots = 9;  // M0..M7 scratch + P

constexpr int kGemmSlots = 6;    // S0..S4 scratch + P (see the slot plan in pwc_taylor_cta_kernel)

struct GemmParams {
    CtaParams c;       // same fields as the Pade CTA kernel (G, signals, hlist, sizes, outputs, ws, use_smem)
    const cplx* TR;    // [(Bm), K+1] trace shifts already subtracted from G's diagonals, or null
    const double* RS;  // [(Bm), K+1, D] row sums of |G_k| (after the shift): inf-norm bound without a pass over the slice, or null
    int DP;            // D rounded up to a multiple of 8 (tile extent)
    int LD;            // leading dimension of every workspace matrix (DP, or DP + 4 in shared memory)
    int g_in_smem;     // the (shared) generators are staged in shared memory after the matrix slots
};

struct ProductParams {
    const cplx* mats;    // [B, M, D, D]  or the gate table [Gn, D, D] when idx != null
    const int* idx;      // [B, M] gate indices or null
    const int* lens;     // [B] valid length per batch row or null (= M)
    int B, M, D;
    int S, seg_len;      // segments per batch row
    cplx* out;           // [B, S, D, D]
    cplx* ws;            // gridDim.x * 2 * D * D when !use_smem
    int use_smem;
};

// Gated launch: wait until batch row b of the control fields has arrived (warp-uniform: every lane polls).  The host
// enqueues the chunked host->device copies of `signals` on a copy stream, each followed by a 4-byte copy that raises
// gate[0], and launches ONE persistent kernel; warps pull units in batch order and spin here only if they overtake the
// copy engine.  The patience counter restarts whenever the copy engine makes progress; after ~4 s WITHOUT progress the
// warp gives up, raises gate[1] (the host reads it at its next synchronisation point and reports the failure) and
// returns false.
// `known` caches the last counter value this warp has seen: rows below it need no further look at the counter (the acquire
// that saw them orders every later load), so once the copy engine is done the kernel runs without touching the gate.
__device__ __forceinline__ bool wait_rows_ready(unsigned int* gate, const int b, unsigned int& known) {
    if ((unsigned int)b < known) return true;
    unsigned int spins = 0, v, last = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gate) : "memory");
        if (v > (unsigned int)b) { known = v; return true; }
        if (v != last) { last = v; spins = 0; }
        __nanosleep(256);
        if (++spins > (1u << 24)) {
            atomicExch(gate + 1, 1u);
            return false;
        }
    }
}

// Control-field sample.  GATED = 0: device-resident signals, read-only for the kernel's lifetime: ld.global.nc.
// In a gated launch the copy engine is still writing the buffer while the kernel runs, so the non-coherent path is out of
// contract.  GATED = 1: the loads bypass L1 (ld.global.cg) after the acquire on gate[0] -- needed when batch rows share cache
// lines (K*N*8 not a multiple of 128: the line holding the tail of row b also holds the head of row b+1, which may not have
// landed when the line is first fetched).  GATED = 2: every row starts on its own 128-byte line, so a line is first touched
// only after the acquire load has seen its row; ordinary cached loads (ld.global.ca) are then coherent by construction and
// keep the L1 hit rate of the ungated kernel (the .cg variant measured 4 % slower at the headline shape).
template <int GATED>
__device__ __forceinline__ double load_signal(const double* ptr) {
    if constexpr (GATED == 1) return __ldcg(ptr);
    else if constexpr (GATED == 2) return __ldca(ptr);
    else return __ldg(ptr);
}

// fused d = 9 gradient kernel (grad_blk9.cuh)
struct Grad9Params {
    const cplx* G;          // [(K+1), d, d] trace-shifted generators (forward model blob)
    const double* RS;       // [(K+1), d] row sums
    const cplx* TR;         // [K+1] trace shifts
    const double* signals;  // [B, K, N]
    const cplx* Ybound;     // [B, Q, d, d]: Y at the first slice of every chunk
    double* grad;           // [B, K, N]
    int B, K, N, d, Q, CL;
};


// fused unitary-recurrence gradient kernel on the DMMA product (grad_ucta.cuh)
struct GradUParams {
    const cplx* G;          // [(K+1), D, D] trace-shifted generators
    const double* RS;       // [(K+1), D]
    const cplx* TR;         // [K+1]
    const double* signals;  // [B, K, N]
    const cplx* Ybound;     // [B, Q, D, D]
    double* grad;           // [B, K, N]
    int B, K, N, D, DP, LD, Q, CL;
};


}  // namespace c3b
