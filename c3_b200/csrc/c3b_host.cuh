// Host-side glue shared by the translation units of libc3b200.so: error reporting, per-thread tuning, launch counter and
// the prototypes of the kernel launchers (each kernel family is compiled in its own .cu so that the library builds in
// parallel and a kernel edit recompiles one file).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/c3b200.h"
#include "c3b_params.cuh"

namespace c3b {

// ---- errors ----------------------------------------------------------------------------------------------------------
// Sets the calling thread's c3b_last_error() message and returns `code`.
int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return ::c3b::fail(C3B_ECUDA, "C3:ERROR: %s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// one more kernel launched by this library (bench.py's gpu_launches)
void count_launch();

// SM count of the current device (148 on B200; cached per thread and device)
int num_sms();

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- tuning ----------------------------------------------------------------------------------------------------------
// PER CALLING THREAD (thread_local): one thread per GPU is the intended use, and a thread that changes a knob (tests,
// A/B measurements) cannot disturb the launches of another.  c3b_set_tuning edits the calling thread's copy.
struct Tuning {
    long long target_units = 0;      // lane-group kernels: aim for this many warp work units per launch; 0: 4096 up to 768 batch rows,
                                     // 12288 above (scratch/seg_sweep.py: fewer, longer units win until the tail of the last wave shows)
    long long min_chunk = 8;         // lane-group kernels: minimum slices per lane group
    long long d9_variant = 1;        // d = 9: 0 generic 3x3-block kernel, 1 own-block shared-memory kernel, 2 shuffle-exchange kernel
    long long d9_skew = 120;         // shuffle kernel: clocks between the early and the late half of a CTA's warps
    long long force_cta = 0;         // route everything to the CTA kernels (testing)
    long long cta_variant = 1;       // 0: literal Higham (Pade + pivoted Gauss-Jordan) cross-check, 1: four-product Taylor scheme on DMMA tiles
    long long cta_threads = 256;     // DMMA CTA kernel, DP = 32: 256 threads (two CTAs per SM) or 512 (one)
    long long gemm_big = 0;          // DMMA CTA kernel, DP = 88 (D = 81): macro-tile shape (0: 3 x 2, 1: 2 x 2)
    long long norm_bound = 1;        // DMMA CTA kernel: scaling from the row-sum bound (1) or the exact inf-norm of every slice (0)
    long long seq_variant = 1;       // evaluate_sequences: 1 lane-group kernel for small d, 0 CTA-per-sequence product kernel
    long long grad_variant = 1;      // 1: best available (fused lockstep kernel at closed d = 7..9, Frechet kernels elsewhere), 0: augmented
                                     // exponential, 2: the stored-propagator Frechet kernels everywhere (cross-check)
    long long grad_chunk = 0;        // fused gradient kernels: slices per chunk (0: chosen from the batch size)
    long long grad_unitary = -1;     // closed-system Hamiltonians Hermitian? 1 yes, 0 no, -1 check on the device (4-byte read-back)
    long long profile = 0;           // bracket the main PWC kernel of each call by events (c3b_last_kernel_ms)
};
Tuning& tuning();

// ---- dimension tables ------------------------------------------------------------------------------------------------
// lane-group kernel instantiations: padded dimension for d <= 12, 0 otherwise
inline int blk_template_dim(int d) {
    const int dims[] = {2, 3, 4, 6, 8, 9, 12};
    for (int t : dims)
        if (d <= t) return t;
    return 0;
}
// matrices per warp of the lane-group kernel that serves dimension d
inline int blk_groups_per_warp(int d) {
    switch (blk_template_dim(d)) {
        case 2: case 3: return 32;
        case 4: case 6: return 8;
        case 8: case 12: return 2;
        default: return 3;
    }
}
inline int round8(int D) { return (D + 7) & ~7; }
inline size_t cta_smem_bytes(int D) { return (((size_t)D * sizeof(int) + 15) & ~(size_t)15) + (size_t)kCtaSlots * D * D * sizeof(cplx); }
// leading dimension of the shared-memory matrices: DP + 4 (bank padding), except DP = 32 (XOR-swizzled, LD = 32)
inline int gemm_ld_smem(int D) { return round8(D) == 32 ? 32 : round8(D) + 4; }
inline size_t gemm_mats_bytes(int D) { return (size_t)kGemmSlots * round8(D) * gemm_ld_smem(D) * sizeof(cplx); }
// CTAs per SM of the shared-memory DMMA kernel (matrix slots only; 4 at most)
inline int gemm_ctas_per_sm(int D) {
    int n = (int)((size_t)220 * 1024 / (gemm_mats_bytes(D) + 1024));
    return n < 1 ? 1 : (n > 4 ? 4 : n);
}
// the shared-model generators ride along in shared memory when they fit next to the matrix slots WITHOUT costing a
// resident CTA (two CTAs per SM overlap one's element-wise phases with the other's products; the generators are then read
// through L1)
inline bool gemm_g_in_smem(int D, int K, int batched_model) {
    return !batched_model &&
           (size_t)gemm_ctas_per_sm(D) * (gemm_mats_bytes(D) + (size_t)(K + 1) * D * D * sizeof(cplx) + 1024) <= (size_t)220 * 1024;
}
inline size_t gemm_smem_bytes(int D, int K = -1, int batched_model = 1) {
    return gemm_mats_bytes(D) + ((K >= 0 && gemm_g_in_smem(D, K, batched_model)) ? (size_t)(K + 1) * D * D * sizeof(cplx) : 0);
}
// persistent grid of the CTA kernels
int cta_grid(int D, long long units);

// ---- launchers (one per kernel family; every one returns a C3B_* status) ----------------------------------------------
// k_small.cu: lane-group kernels, d <= 12 except 9
int launch_small(const RowsParams& rp, unsigned int* counter, cudaStream_t st);
int launch_fold_small(const ProductParams& pp, cudaStream_t st);   // -1: dimension not served
int launch_seq_small(const cplx* gates, int Gn, const int* idx, const int* lens, int S, int Lmax, int d, cplx* out, cudaStream_t st);  // -1: n/a
// k_d9.cu
int launch_d9(const RowsParams& rp, unsigned int* counter, int variant, cudaStream_t st);
bool d9_gated_supported(int variant);
// k_gemm.cu: four-product Taylor scheme on DMMA tiles (any d)
int launch_gemm(const CtaParams& cp, const cplx* TR, const double* RS, int grid, cudaStream_t st);
// k_cta.cu: literal Higham kernel, ordered products, model setup, Kronecker product
int launch_cta(const CtaParams& cp, const cplx* TR, int grid, cudaStream_t st);
int launch_product(ProductParams pp, cudaStream_t st);
int launch_setup_closed(const cplx* h0, const cplx* hks, cplx* G, int Bm, int K, int d, double dt, cudaStream_t st);
int launch_setup_lindblad(const cplx* h0, const cplx* hks, const cplx* col_ops, cplx* G, int Bm, int K, int C, int d, double dt,
                          cudaStream_t st);
int launch_trace_shift(cplx* G, cplx* TR, long long nmat, int D, cudaStream_t st);
int launch_rowsum(const cplx* G, double* RS, long long nrows, int D, cudaStream_t st);
int launch_kron(const cplx* A, const cplx* B, cplx* out, int batch, int ra, int ca, int rb, int cb, long long sa, long long sb,
                cudaStream_t st);
// k_grad.cu: adjoint sweeps and Frechet contraction
int launch_grad_suffix(int variant, const cplx* dUs, const cplx* Ubar, cplx* Psi, double* alpha, int nb, int N, int d, cudaStream_t st);
int launch_grad_prefix_frechet(const cplx* dUs, cplx* PsiM, int nb, int N, int d, cudaStream_t st);
int launch_grad_prefix_aug(const cplx* dUs, const cplx* Psi, const cplx* h0, const cplx* hks, const double* sig, cplx* Haug,
                           double dt, int nb, int K, int N, int d, cudaStream_t st);
int launch_grad_frechet(const cplx* G, const double* RS, const cplx* TR, const double* sig, const cplx* M, const double* alpha,
                        double* grad, int nb, int K, int N, int d, cudaStream_t st);
int launch_grad_contract(const cplx* Eaug, const cplx* hks, const double* alpha, double* grad, double dt, int nb, int K, int N, int d,
                         cudaStream_t st);

// k_grad9.cu: fused lockstep gradient kernel for closed d = 7..9 (grad_blk9.cuh)
bool grad9_supported(int K, int d);
int launch_hermitian_check(const cplx* h0, const cplx* hks, int K, int d, unsigned int* flag, cudaStream_t st);
int launch_grad9_prefix(const cplx* seg, cplx* F, cplx* U, int B, int Q, int d, cudaStream_t st);
int launch_grad9_ybound(const cplx* F, const cplx* U, const cplx* Ubar, cplx* Ybound, int B, int Q, int d, cudaStream_t st);
int launch_grad9(const Grad9Params& gp, unsigned int* counter, cudaStream_t st);

// k_grad_cta.cu: fused unitary-recurrence gradient on the DMMA product (closed 16 < d <= 32, grad_ucta.cuh)
bool grad_ucta_supported(int D);
int launch_grad_ucta(GradUParams gp, cudaStream_t st);

// k_grad_cta.cu: the same adjoint scheme on the CTA-cooperative DMMA product (closed d > 16, Lindblad superoperators)
bool grad_cta_uses_smem(int D);
int grad_cta_frechet_grid(int D, long long units);
size_t grad_cta_workspace_bytes(int Bc, int N, int D);
int launch_grad_cta(const cplx* G, const double* RS, const cplx* TR, const double* sig, const cplx* dUs, const cplx* Ubar, cplx* PsiM,
                    double* alpha, double* grad, int nb, int K, int N, int D, cplx* ws, cudaStream_t st);

}  // namespace c3b
