// Batched dressing of model samples: for every drift Hamiltonian of a batch,
//   e, v = eigh(drift);  reorder the eigenvectors by their overlap with the bare states;  T = v * signed_reorder^T;
//   dressed X = T^dag X T  for the drift, the control Hamiltonians and the collapse operators
// (Model.update_drift_eigen / reorder_frame / update_dressed, c3/model.py:453-534), which the reference runs once per
// model update on the host.  When an optimiser samples MODEL parameters (model learning, robust control) this runs
// once per sample: here one warp per sample diagonalises the d x d Hermitian drift by cyclic Jacobi rotations in
// shared memory (d <= 32) and applies the transform to that sample's operators.
//
// Conventions pinned by the reference:
//   * eigenvalues ascending (tf.linalg.eigh) -- the order matters only for `ordered == 0`;
//   * ordered: state s takes the eigenvector with |v[s,m]|^2 > 0.5; if some eigenvector has no such component
//     ("overly dressed") the assignment is greedy on the largest remaining |v|^2 (first flat index on ties);
//   * T[:, s] = v[:, m(s)] * sign(Re v[s, m(s)]).  Eigenvectors are defined up to a phase: for the real-symmetric
//     drifts of the reference's chip models the Jacobi vectors are real and T equals the reference's; for complex
//     Hermitian drifts the phase is fixed by making v[s, m(s)] real positive (a valid dressing, not LAPACK's phase).
#pragma once
#include "c3b_common.cuh"

namespace c3b {

struct DressParams {
    const cplx* drift;     // [B, d, d] Hermitian
    const cplx* ops;       // [B or 1, M, d, d] operators to dress (control Hamiltonians, collapse operators) or null
    int ops_batched;
    int B, M, d, ordered;
    double* eigenframe;    // [B, d]
    cplx* transform;       // [B, d, d]
    cplx* dressed_drift;   // [B, d, d] or null
    cplx* dressed_ops;     // [B, M, d, d] or null
    int* info;             // [B] number of Jacobi sweeps used (>= 100: not converged), or null
};

// out = T^dag X T with one warp; tmp [d*d] scratch
__device__ __forceinline__ void warp_dress(cplx* out, const cplx* X, const cplx* T, cplx* tmp, const int d, const int lane) {
    for (int e = lane; e < d * d; e += 32) {           // tmp = X T
        const int i = e / d, j = e - i * d;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k < d; ++k) cfma(acc, X[i * d + k], T[k * d + j]);
        tmp[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < d * d; e += 32) {           // out = T^dag tmp
        const int i = e / d, j = e - i * d;
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k < d; ++k) cfma(acc, cmake(T[k * d + i].x, -T[k * d + i].y), tmp[k * d + j]);
        out[e] = acc;
    }
    __syncwarp();
}

// one warp per batch element; dynamic smem = warps * (4 d*d complex + 4 d doubles)
__global__ void dress_kernel(const DressParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int b = blockIdx.x * wpb + warp;
    if (b >= p.B) return;
    const int d = p.d, dd = d * d;
    const size_t per_warp = (size_t)4 * dd * sizeof(cplx) + (size_t)4 * d * sizeof(double);
    unsigned char* base = smem_raw + (size_t)warp * per_warp;
    cplx* A = reinterpret_cast<cplx*>(base);           // working copy, diagonalised in place
    cplx* V = A + dd;                                  // accumulated rotations (columns = eigenvectors)
    cplx* T = V + dd;                                  // transform
    cplx* W = T + dd;                                  // scratch
    double* ev = reinterpret_cast<double*>(W + dd);    // [d] eigenvalues
    int* perm = reinterpret_cast<int*>(ev + d);        // [d] ascending order
    int* assign = perm + d;                            // [d] eigenvector index taken by state s
    double* vs = ev + 3 * d;                           // [d] scratch

    const cplx* H = p.drift + (size_t)b * dd;
    double fro = 0.0;
    for (int e = lane; e < dd; e += 32) {
        const cplx h = H[e];
        A[e] = h;
        V[e] = cmake((e / d) == (e % d) ? 1.0 : 0.0, 0.0);
        fro = fma(h.x, h.x, fma(h.y, h.y, fro));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, o);
    __syncwarp();

    // ---- cyclic Jacobi: A <- J^dag A J, V <- V J,  J = [[c, s e^{i phi}], [-s e^{-i phi}, c]] on (p, q) -------------
    int sweeps = 0;
    for (; sweeps < 100; ++sweeps) {
        double off = 0.0;
        for (int e = lane; e < dd; e += 32) {
            const int i = e / d, j = e - i * d;
            if (i < j) off = fma(A[e].x, A[e].x, fma(A[e].y, A[e].y, off));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
        if (off <= 1e-30 * fro) break;
        for (int pi = 0; pi < d - 1; ++pi) {
            for (int qi = pi + 1; qi < d; ++qi) {
                const cplx apq = A[pi * d + qi];
                const double mag = sqrt(fma(apq.x, apq.x, apq.y * apq.y));
                if (mag < 1e-300) continue;            // uniform: every lane reads the same element
                const double app = A[pi * d + pi].x, aqq = A[qi * d + qi].x;
                const double tau = (aqq - app) / (2.0 * mag);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                const cplx ph = cmake(apq.x / mag, apq.y / mag);          // e^{i phi}
                const cplx sp = cmake(s * ph.x, s * ph.y);                // s e^{i phi}
                const cplx sm = cmake(s * ph.x, -s * ph.y);               // s e^{-i phi}
                __syncwarp();
                // columns p, q of A and V:  col_p' = c col_p - s e^{-i phi} col_q,  col_q' = s e^{i phi} col_p + c col_q
                for (int k = lane; k < d; k += 32) {
                    const cplx ap = A[k * d + pi], aq = A[k * d + qi];
                    A[k * d + pi] = cmake(c * ap.x - (sm.x * aq.x - sm.y * aq.y), c * ap.y - (sm.x * aq.y + sm.y * aq.x));
                    A[k * d + qi] = cmake((sp.x * ap.x - sp.y * ap.y) + c * aq.x, (sp.x * ap.y + sp.y * ap.x) + c * aq.y);
                    const cplx vp = V[k * d + pi], vq = V[k * d + qi];
                    V[k * d + pi] = cmake(c * vp.x - (sm.x * vq.x - sm.y * vq.y), c * vp.y - (sm.x * vq.y + sm.y * vq.x));
                    V[k * d + qi] = cmake((sp.x * vp.x - sp.y * vp.y) + c * vq.x, (sp.x * vp.y + sp.y * vp.x) + c * vq.y);
                }
                __syncwarp();
                // rows p, q:  row_p' = c row_p - s e^{i phi} row_q,  row_q' = s e^{-i phi} row_p + c row_q
                for (int k = lane; k < d; k += 32) {
                    const cplx ap = A[pi * d + k], aq = A[qi * d + k];
                    A[pi * d + k] = cmake(c * ap.x - (sp.x * aq.x - sp.y * aq.y), c * ap.y - (sp.x * aq.y + sp.y * aq.x));
                    A[qi * d + k] = cmake((sm.x * ap.x - sm.y * ap.y) + c * aq.x, (sm.x * ap.y + sm.y * ap.x) + c * aq.y);
                }
                __syncwarp();
                if (lane == 0) {
                    A[pi * d + qi] = cmake(0.0, 0.0);
                    A[qi * d + pi] = cmake(0.0, 0.0);
                    A[pi * d + pi].y = 0.0;
                    A[qi * d + qi].y = 0.0;
                }
                __syncwarp();
            }
        }
    }
    if (lane == 0 && p.info) p.info[b] = sweeps;

    // ---- ascending eigenvalue order (as tf.linalg.eigh returns them) --------------------------------------------------
    for (int m = lane; m < d; m += 32) { ev[m] = A[m * d + m].x; }
    __syncwarp();
    for (int m = lane; m < d; m += 32) {               // rank of ev[m]; ties by index
        int r = 0;
        for (int j = 0; j < d; ++j) r += (ev[j] < ev[m]) || (ev[j] == ev[m] && j < m);
        perm[r] = m;
    }
    __syncwarp();

    // ---- assignment of eigenvectors (sorted index m) to bare states s ---------------------------------------------------
    // W[s * d + m] = |v[s, perm[m]]|^2 as a real number in .x; greedy on the largest remaining entry reproduces the
    // "> 0.5" rule whenever that rule applies (an entry above 0.5 is the maximum of its row and of its column)
    if (p.ordered) {
        for (int e = lane; e < dd; e += 32) {
            const int sidx = e / d, m = e - sidx * d;
            const cplx v = V[sidx * d + perm[m]];
            W[e] = cmake(fma(v.x, v.x, v.y * v.y), 0.0);
        }
        __syncwarp();
        for (int it = 0; it < d; ++it) {
            double best = -1.0;
            int bidx = dd;
            for (int e = lane; e < dd; e += 32)
                if (W[e].x > best) { best = W[e].x; bidx = e; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            const int sidx = bidx / d, m = bidx - sidx * d;
            if (lane == 0) assign[sidx] = m;
            __syncwarp();
            for (int k = lane; k < d; k += 32) { W[sidx * d + k].x = -1.0; W[k * d + m].x = -1.0; }
            __syncwarp();
        }
    } else {
        for (int m = lane; m < d; m += 32) assign[m] = m;
        __syncwarp();
    }

    // ---- eigenframe and transform ------------------------------------------------------------------------------------
    for (int sidx = lane; sidx < d; sidx += 32) {
        const int col = perm[assign[sidx]];
        p.eigenframe[(size_t)b * d + sidx] = ev[col];
        // phase factor conj(v[s, m(s)]) / |v[s, m(s)]|: the reference's sign(Re v) for real vectors
        const cplx v = V[sidx * d + col];
        const double mag = sqrt(fma(v.x, v.x, v.y * v.y));
        vs[sidx] = 0.0;
        cplx f = cmake(1.0, 0.0);
        if (p.ordered && mag > 0.0) f = cmake(v.x / mag, -v.y / mag);
        W[sidx] = f;                                   // W[0..d) reused as the per-state phase factors
    }
    __syncwarp();
    for (int e = lane; e < dd; e += 32) {
        const int k = e / d, sidx = e - k * d;
        const cplx v = V[k * d + perm[assign[sidx]]];
        const cplx t = cmul(v, W[sidx]);
        T[e] = t;
        p.transform[(size_t)b * dd + e] = t;
    }
    __syncwarp();

    // ---- dressed operators ---------------------------------------------------------------------------------------------
    if (p.dressed_drift) {
        for (int e = lane; e < dd; e += 32) A[e] = H[e];
        __syncwarp();
        warp_dress(V, A, T, W, d, lane);               // V is free now
        for (int e = lane; e < dd; e += 32) p.dressed_drift[(size_t)b * dd + e] = V[e];
        __syncwarp();
    }
    if (p.dressed_ops && p.ops) {
        for (int m = 0; m < p.M; ++m) {
            const cplx* X = p.ops + ((size_t)(p.ops_batched ? b : 0) * p.M + m) * dd;
            for (int e = lane; e < dd; e += 32) A[e] = X[e];
            __syncwarp();
            warp_dress(V, A, T, W, d, lane);
            cplx* o = p.dressed_ops + ((size_t)b * p.M + m) * dd;
            for (int e = lane; e < dd; e += 32) o[e] = V[e];
            __syncwarp();
        }
    }
}

}  // namespace c3b
