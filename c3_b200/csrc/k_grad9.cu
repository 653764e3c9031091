// Fused d = 9 gradient path: chunk-boundary kernel + the lockstep Frechet kernel of grad_blk9.cuh,
// and the Hermiticity check that guards its unitarity assumption.
#include "c3b_host.cuh"
#include "grad_blk9.cuh"

namespace c3b {

namespace {
constexpr int kGrad9Warps = 8;

// flag[0] = 1 when some h[m] (m = 0 .. M-1, d x d) differs from its conjugate transpose by more than 1e-13 of its largest entry
__global__ void hermitian_check_kernel(const cplx* __restrict__ h0, const cplx* __restrict__ hks, const int K, const int d,
                                       unsigned int* __restrict__ flag) {
    const int m = blockIdx.x;
    const cplx* h = (m == 0) ? h0 : hks + (size_t)(m - 1) * d * d;
    __shared__ double s_max[32], s_dev[32];
    double mx = 0.0, dev = 0.0;
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
        const int i = e / d, j = e - i * d;
        const cplx a = h[e], b = h[j * d + i];
        mx = fmax(mx, fmax(fabs(a.x), fabs(a.y)));
        dev = fmax(dev, fmax(fabs(a.x - b.x), fabs(a.y + b.y)));
    }
    mx = warp_max(mx); dev = warp_max(dev);
    if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = mx; s_dev[threadIdx.x >> 5] = dev; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mx = fmax(mx, s_max[w]); dev = fmax(dev, s_dev[w]); }
        if (!(dev <= 1e-13 * mx)) atomicExch(flag, 1u);
    }
}
}  // namespace

bool grad9_supported(int K, int d) {
    return blk_template_dim(d) == 9 && K >= 1 && Grad9T::smem_bytes(K, kGrad9Warps) <= (size_t)227 * 1024;
}

int launch_hermitian_check(const cplx* h0, const cplx* hks, int K, int d, unsigned int* flag, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(unsigned int), st));
    hermitian_check_kernel<<<K + 1, 128, 0, st>>>(h0, hks, K, d, flag);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

namespace {
int boundary_wpb(int d) { return d <= 16 ? 4 : 1; }      // three d x d matrices per warp: <= 48 KB per block up to d = 32
}  // namespace

// seg [B,Q,d,d] chunk products of the forward launch (Q == 1: seg may be the forward result itself) -> prefix products
// F [B,Q,d,d] and U [B,d,d]
int launch_grad9_prefix(const cplx* seg, cplx* F, cplx* U, int B, int Q, int d, cudaStream_t st) {
    const int wpb = boundary_wpb(d);
    const size_t smem = (size_t)wpb * 3 * d * d * sizeof(cplx);
    CUDA_TRY(cudaFuncSetAttribute(grad9_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grad9_prefix_kernel<<<(B + wpb - 1) / wpb, wpb * 32, smem, st>>>(seg, F, U, B, Q, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

// Ybound [B,Q,d,d] = F_q (Ubar^dag U) F_q^dag
int launch_grad9_ybound(const cplx* F, const cplx* U, const cplx* Ubar, cplx* Ybound, int B, int Q, int d, cudaStream_t st) {
    const int wpb = boundary_wpb(d);
    const size_t smem = (size_t)wpb * 3 * d * d * sizeof(cplx);
    CUDA_TRY(cudaFuncSetAttribute(grad9_ybound_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long warps = (long long)B * Q;
    grad9_ybound_kernel<<<(int)((warps + wpb - 1) / wpb), wpb * 32, smem, st>>>(F, U, Ubar, Ybound, B, Q, d);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

int launch_grad9(const Grad9Params& gp, unsigned int* counter, cudaStream_t st) {
    const size_t smem = Grad9T::smem_bytes(gp.K, kGrad9Warps);
    auto kern = grad_blk9_kernel<kGrad9Warps>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const long long units = ((long long)gp.B * gp.Q + 2) / 3;                       // three chunks per warp
    long long grid = (units + kGrad9Warps - 1) / kGrad9Warps;
    if (grid > num_sms()) grid = num_sms();
    if (grid < 1) grid = 1;
    kern<<<(int)grid, kGrad9Warps * 32, smem, st>>>(gp, counter);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace c3b
