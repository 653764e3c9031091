// Gradient kernels: adjoint sweeps, Frechet derivative of the Taylor scheme, contraction.
#include "c3b_host.cuh"
#include "grad.cuh"

namespace c3b {

namespace {
// sweep kernels: one warp per batch row with 3-4 matrices in shared memory; d = 32 needs 64 KB per warp
int sweep_warps(int d) {
    int wpb = (int)((size_t)192 * 1024 / ((size_t)4 * d * d * sizeof(cplx)));
    if (wpb > 4) wpb = 4;
    if (wpb < 1) wpb = 1;
    return wpb;
}
}  // namespace

int launch_grad_suffix(int variant, const cplx* dUs, const cplx* Ubar, cplx* Psi, double* alpha, int nb, int N, int d, cudaStream_t st) {
    const int wpb = sweep_warps(d);
    const size_t dd = (size_t)d * d;
    const size_t smem = (size_t)wpb * 3 * dd * sizeof(cplx);
    auto kern = variant == 1 ? grad_suffix2_kernel : grad_suffix_kernel;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)wpb * 4 * dd * sizeof(cplx))));
    kern<<<(nb + wpb - 1) / wpb, wpb * 32, smem, st>>>(dUs, Ubar, Psi, alpha, nb, N, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_grad_prefix_frechet(const cplx* dUs, cplx* PsiM, int nb, int N, int d, cudaStream_t st) {
    const int wpb = sweep_warps(d);
    const size_t smem = (size_t)wpb * 4 * d * d * sizeof(cplx);
    CUDA_TRY(cudaFuncSetAttribute(grad_prefix2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grad_prefix2_kernel<<<(nb + wpb - 1) / wpb, wpb * 32, smem, st>>>(dUs, PsiM, nb, N, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_grad_prefix_aug(const cplx* dUs, const cplx* Psi, const cplx* h0, const cplx* hks, const double* sig, cplx* Haug,
                           double dt, int nb, int K, int N, int d, cudaStream_t st) {
    const int wpb = sweep_warps(d);
    const size_t smem = (size_t)wpb * 4 * d * d * sizeof(cplx);
    CUDA_TRY(cudaFuncSetAttribute(grad_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grad_prefix_kernel<<<(nb + wpb - 1) / wpb, wpb * 32, smem, st>>>(dUs, Psi, h0, hks, sig, Haug, dt, nb, K, N, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_grad_frechet(const cplx* G, const double* RS, const cplx* TR, const double* sig, const cplx* M, const double* alpha,
                        double* grad, int nb, int K, int N, int d, cudaStream_t st) {
    const size_t per_warp = (size_t)kFrechetBufs * d * d * sizeof(cplx);
    int fw = (int)((size_t)96 * 1024 / per_warp);
    if (fw < 1) fw = 1;
    if (fw > 4) fw = 4;
    const size_t smem = fw * per_warp;
    CUDA_TRY(cudaFuncSetAttribute(grad_frechet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)num_sms() * per_sm;
    const long long needb = ((long long)nb * N + fw - 1) / fw;
    if (grid > needb) grid = needb;
    grad_frechet_kernel<<<(int)grid, fw * 32, smem, st>>>(G, RS, TR, sig, M, alpha, grad, nb, K, N, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_grad_contract(const cplx* Eaug, const cplx* hks, const double* alpha, double* grad, double dt, int nb, int K, int N, int d,
                         cudaStream_t st) {
    const long long warps = (long long)nb * N;
    grad_contract_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(Eaug, hks, alpha, grad, dt, nb, K, N, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace c3b
