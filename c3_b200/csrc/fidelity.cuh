// Goal-function reductions that consume the propagators.
//
// Every fidelity the reference computes from a gate propagator reduces to ONE gathered overlap
//   t[b] = sum_{I,J} M[b, sel[I], sel[J]] * conj(T[I,J])
// between the computational-subspace block of the propagator and the ideal gate:
//   unitary_infid              1 - |tr(P^T U P G^dag) / c|^2              = 1 - |t|^2 / c^2
//       (c3/libraries/fidelities.py:152-183, c3/utils/tf_utils.py:326-364, :430-438)
//   average_infid              1 - |(chi_00 / c + 1) / (c + 1)|, chi_00 = |tr(A^dag G)|^2 = |t|^2
//       (fidelities.py:288-311; tf_utils.py:380-413: the (0,0) chi element of Lambda (x) Lambda^* in the
//        unnormalised Pauli basis is |tr Lambda|^2 -- checked against the op-for-op oracle)
//   lindbladian_unitary_infid  1 - |sqrt(tr(S_comp (G (x) G^*)^dag)) / c|^2 = 1 - |t| / c^2, sel = pairs
//       (fidelities.py:221-249, tf_utils.py:367-376)
//   lindbladian_average_infid  1 - |conj(t) / c + 1| / (c + 1)           (fidelities.py:377-399)
// so the all-gather after a batched propagation carries 8 bytes per batch element instead of 16 d^2.
// The ORBIT goal (fidelities.py:753-790) needs only U_seq |0>: a chain of matrix-VECTOR products
// (seq_state_kernel), 1/d of the work of the sequence propagators evaluate_sequences builds.
#pragma once
#include "c3b_common.cuh"

namespace c3b {

__device__ __forceinline__ double infid_from_overlap(const cplx t, const int C, const int mode) {
    const double a2 = fma(t.x, t.x, t.y * t.y);
    if (mode == 0) return 1.0 - a2 / ((double)C * C);                       // unitary_infid, lvls = C
    if (mode == 1) return 1.0 - (a2 / C + 1.0) / (C + 1.0);                 // average_infid, d = C
    const double c = sqrt((double)C);                                       // superoperator block: C = c^2
    if (mode == 2) return 1.0 - sqrt(a2) / (c * c);                         // lindbladian_unitary_infid
    const double re = t.x / c + 1.0, im = -t.y / c;                         // lindbladian_average_infid
    return 1.0 - sqrt(fma(re, re, im * im)) / (c + 1.0);
}

// one warp per batch element
__global__ void __launch_bounds__(128) gate_overlap_kernel(const cplx* __restrict__ M, const cplx* __restrict__ T,
                                                           const int* __restrict__ sel, int B, int D, int C, int mode,
                                                           double* __restrict__ infid_out, cplx* __restrict__ overlap_out) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const cplx* Mb = M + (size_t)b * D * D;
    cplx t = cmake(0.0, 0.0);
    for (int e = lane; e < C * C; e += 32) {
        const int I = e / C, J = e - I * C;
        const cplx m = Mb[(size_t)sel[I] * D + sel[J]];
        const cplx g = T[e];
        t.x += fma(m.x, g.x, m.y * g.y);      // m * conj(g)
        t.y += fma(m.y, g.x, -m.x * g.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    }
    if (lane == 0) {
        if (overlap_out) overlap_out[b] = t;
        if (infid_out) infid_out[b] = infid_from_overlap(t, C, mode);
    }
}

// Cotangent of U for L = sum_b gbar[b] * infid[b] (modes 0 and 1; torch convention dL = Re tr(Ubar^dag dU)):
//   Ubar[b, sel[I], sel[J]] = coef * gbar[b] * t[b] * T[I,J],  coef = -2/C^2 (mode 0), -2/(C (C+1)) (mode 1);
// all other entries are zero (the caller clears Ubar first).
__global__ void gate_overlap_grad_kernel(const cplx* __restrict__ overlap, const cplx* __restrict__ T,
                                         const int* __restrict__ sel, const double* __restrict__ gbar, int B, int D,
                                         int C, int mode, cplx* __restrict__ Ubar) {
    const long long total = (long long)B * C * C;
    const double coef = mode == 0 ? -2.0 / ((double)C * C) : -2.0 / ((double)C * (C + 1.0));
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / (C * C));
        const int r = (int)(e - (long long)b * C * C);
        const int I = r / C, J = r - I * C;
        const cplx t = overlap[b];
        const double s = coef * (gbar ? gbar[b] : 1.0);
        const cplx g = T[r];
        Ubar[(size_t)b * D * D + (size_t)sel[I] * D + sel[J]] = cmake(s * fma(t.x, g.x, -t.y * g.y), s * fma(t.x, g.y, t.y * g.x));
    }
}

// psi_s = G[idx[s, len_s - 1]] ... G[idx[s, 0]] psi0, one warp per sequence; populations
//   closed system (ld == 0):  pops[s, i] = |psi_i|^2, i < D                     (c3/experiment.py:622-624)
//   Lindblad (ld = d, D = d^2): pops[s, i] = Re psi[i d + i], i < d  (diag of tf_vec_to_dm, :617-621)
// replaces the gate loop of Experiment.evaluate_legacy (c3/experiment.py:273-302) and the
// evaluate_sequences + matmul + populations chain of orbit_infid (c3/libraries/fidelities.py:762-770).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) seq_state_kernel(const cplx* __restrict__ gates, const int* __restrict__ idx,
                                                               const int* __restrict__ lens, const cplx* __restrict__ psi0,
                                                               int S, int Lmax, int D, int ld, double* __restrict__ pops,
                                                               cplx* __restrict__ psi_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);           // [WARPS][2][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * WARPS + warp;
    if (s >= S) return;
    cplx* cur = sm + (size_t)warp * 2 * D;
    cplx* nxt = cur + D;
    for (int i = lane; i < D; i += 32) cur[i] = psi0 ? psi0[i] : cmake(i == 0 ? 1.0 : 0.0, 0.0);
    __syncwarp();
    const int len = lens[s];
    for (int q = 0; q < len; ++q) {
        const cplx* G = gates + (size_t)idx[(size_t)s * Lmax + q] * D * D;
        for (int r = lane; r < D; r += 32) {
            cplx acc = cmake(0.0, 0.0);
            const cplx* Gr = G + (size_t)r * D;
            for (int j = 0; j < D; ++j) cfma(acc, __ldg(Gr + j), cur[j]);
            nxt[r] = acc;
        }
        __syncwarp();
        cplx* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (psi_out)
        for (int i = lane; i < D; i += 32) psi_out[(size_t)s * D + i] = cur[i];
    if (pops) {
        if (ld == 0) {
            for (int i = lane; i < D; i += 32) pops[(size_t)s * D + i] = fma(cur[i].x, cur[i].x, cur[i].y * cur[i].y);
        } else {
            for (int i = lane; i < ld; i += 32) pops[(size_t)s * ld + i] = cur[i * ld + i].x;
        }
    }
}

}  // namespace c3b
