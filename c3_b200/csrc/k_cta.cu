// CTA-per-unit kernels that are not on the DMMA path: the literal Higham/TF restatement (cross-check), ordered products
// (tf_matmul_n / tf_matmul_left / evaluate_sequences), model setup (generators, trace shift, row sums), Kronecker product.
#include "c3b_host.cuh"
#include "pwc_cta.cuh"
#include "product.cuh"

namespace c3b {

namespace {

template <int CT, int TR, int TC>
int launch_cta_t(const CtaParams& cp, int grid, cudaStream_t st) {
    auto kern = pwc_cta_kernel<CT, TR, TC>;
    size_t smem = cp.use_smem ? cta_smem_bytes(cp.D) : (((size_t)cp.D * sizeof(int) + 15) & ~(size_t)15);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kCtaThreads, smem, st>>>(cp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

template <int CT, int TR, int TC>
int launch_product_t(const ProductParams& pp, int grid, cudaStream_t st) {
    auto kern = product_kernel<CT, TR, TC>;
    const size_t smem = pp.use_smem ? (size_t)2 * pp.D * pp.D * sizeof(cplx) : 0;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kCtaThreads, smem, st>>>(pp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace

int launch_cta(const CtaParams& cp_in, const cplx* TR, int grid, cudaStream_t st) {
    CtaParams cp = cp_in;
    cp.unshift = TR;
    const int D = cp.D;
    if (D <= 16) return launch_cta_t<16, 1, 1>(cp, grid, st);
    if (D <= 32) return launch_cta_t<32, 4, 1>(cp, grid, st);
    if (D <= 64) return launch_cta_t<32, 4, 2>(cp, grid, st);
    return launch_cta_t<32, 4, 3>(cp, grid, st);
}

int launch_product(ProductParams pp, cudaStream_t st) {
    const int D = pp.D;
    if (tuning().seq_variant != 0 && pp.idx == nullptr && pp.lens == nullptr && pp.S == 1 && pp.seg_len >= pp.M && pp.M >= 1 &&
        (long long)pp.B * 2 >= num_sms()) {      // lane-group fold for small d (a few rows only: the CTA kernel has less latency)
        const int rc = launch_fold_small(pp, st);
        if (rc >= 0) return rc;
    }
    pp.use_smem = D <= 64;
    const long long units = (long long)pp.B * pp.S;
    long long g = pp.use_smem ? (long long)num_sms() * 4 : cta_grid(D, pp.B);
    if (g > units) g = units;
    if (g < 1) g = 1;
    if (D <= 16) return launch_product_t<16, 1, 1>(pp, (int)g, st);
    if (D <= 32) return launch_product_t<32, 4, 1>(pp, (int)g, st);
    if (D <= 64) return launch_product_t<32, 4, 2>(pp, (int)g, st);
    return launch_product_t<32, 4, 3>(pp, (int)g, st);
}

int launch_setup_closed(const cplx* h0, const cplx* hks, cplx* G, int Bm, int K, int d, double dt, cudaStream_t st) {
    const long long total = (long long)Bm * (K + 1) * d * d;
    const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    setup_closed_kernel<<<blocks, 256, 0, st>>>(h0, hks, G, Bm, K, d, dt);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_setup_lindblad(const cplx* h0, const cplx* hks, const cplx* col_ops, cplx* G, int Bm, int K, int C, int d, double dt,
                          cudaStream_t st) {
    const long long total = (long long)Bm * (K + 1) * d * d * d * d;
    const int blocks = (int)((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256);
    setup_lindblad_kernel<<<blocks, 256, 0, st>>>(h0, hks, C > 0 ? col_ops : nullptr, G, Bm, K, C, d, dt);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_trace_shift(cplx* G, cplx* TR, long long nmat, int D, cudaStream_t st) {
    trace_shift_kernel<<<(int)((nmat * 32 + 255) / 256), 256, 0, st>>>(G, TR, nmat, D);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_rowsum(const cplx* G, double* RS, long long nrows, int D, cudaStream_t st) {
    rowsum_kernel<<<(int)((nrows * 32 + 255) / 256), 256, 0, st>>>(G, RS, nrows, D);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_kron(const cplx* A, const cplx* B, cplx* out, int batch, int ra, int ca, int rb, int cb, long long sa, long long sb,
                cudaStream_t st) {
    const long long total = (long long)batch * ra * rb * ca * cb;
    long long blocks = (total + 255) / 256;
    if (blocks > 65535) blocks = 65535;
    kron_kernel<<<(int)blocks, 256, 0, st>>>(A, B, out, batch, ra, ca, rb, cb, sa, sb);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace c3b
