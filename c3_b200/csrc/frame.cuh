// Frame rotation and dephasing channel applied to a batch of propagators on the device.
//
// Reference: Experiment.compute_propagators post-multiplies every gate by FR (closed) / SFR = super(FR) (Lindblad) and by the
// dephasing channel (c3/experiment.py:482-522), with
//   FR  = expm( sum_line 1j a_q^dag a_q (freq_line t_final + framechange_line) )                (c3/model.py:536-578)
//   deph = prod_line ( (1 - p_line) Id + p_line super(expm(1j pi a_q^dag a_q)) ),  p = t_final amp strength   (:597-639)
// a_q^dag a_q is the BARE number operator of the qubit the line drives: diagonal in the product basis, so FR, SFR = FR (x) FR^*
// and the dephasing channel are all diagonal and "matrix times U" is a scaling of the ROWS of U:
//   closed:    U[r, :]      *= exp(1j sum_l occ[l, r] phi[b, l])
//   Lindblad:  S[(i,j), :]  *= exp(1j sum_l (occ[l,i] - occ[l,j]) phi[b,l]) * prod_l ((1 - p[b,l]) + p[b,l] (-1)^(occ[l,i] - occ[l,j]))
// with occ[l, s] the occupation number of line l's qubit in product state s.  One launch for the whole batch (every gate of a
// gate set, every parameter sample), phases per batch row: nothing is built on the host, nothing is multiplied as a matrix.
#pragma once
#include "c3b_common.cuh"

namespace c3b {

__global__ void frame_dephase_kernel(cplx* __restrict__ U, const int* __restrict__ occ, const double* __restrict__ phi,
                                     const double* __restrict__ prob, const int B, const int D, const int d, const int L,
                                     const int lindblad) {
    const long long rows = (long long)B * D;
    for (long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; w < rows; w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int b = (int)(w / D), r = (int)(w - (long long)b * D);
        const int i = lindblad ? r / d : r, j = lindblad ? r - i * d : 0;
        double ang = 0.0, scale = 1.0;
        for (int l = 0; l < L; ++l) {
            const int dn = lindblad ? occ[l * d + i] - occ[l * d + j] : occ[l * d + i];
            if (phi != nullptr) ang = fma((double)dn, phi[(size_t)b * L + l], ang);
            if (lindblad && prob != nullptr) {
                const double p = prob[(size_t)b * L + l];
                scale *= (1.0 - p) + ((dn & 1) ? -p : p);
            }
        }
        double sn, cs;
        sincos(ang, &sn, &cs);
        const cplx f = cmake(scale * cs, scale * sn);
        cplx* row = U + ((size_t)b * D + r) * D;
        for (int c = threadIdx.x & 31; c < D; c += 32) row[c] = cmul(f, row[c]);
    }
}

}  // namespace c3b
