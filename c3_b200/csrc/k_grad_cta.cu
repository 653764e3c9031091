// CTA-level gradient kernels on the DMMA product (closed d > 16, Lindblad superoperators): launchers.
#include "c3b_host.cuh"
#include "grad_cta.cuh"
#include "grad_ucta.cuh"

namespace c3b {

namespace {

template <int TM, int TN>
int launch_sweeps_t(GradCtaParams gp, size_t smem, int grid, cudaStream_t st) {
    constexpr int NT = 256;
    gp.slots = kGradSweepSlots;
    auto ks = grad_suffix_cta_kernel<TM, TN, NT>;
    auto kp = grad_prefix_cta_kernel<TM, TN, NT>;
    CUDA_TRY(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ks<<<grid, NT, smem, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    kp<<<grid, NT, smem, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

template <int TM, int TN>
int launch_frechet_t(GradCtaParams gp, size_t smem, int grid, cudaStream_t st) {
    constexpr int NT = 256;
    gp.slots = kGradFrechetSlots;
    auto kf = grad_frechet_cta_kernel<TM, TN, NT>;
    CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kf<<<grid, NT, smem, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace

bool grad_cta_uses_smem(int D) { return round8(D) <= 32; }

// fused unitary-recurrence kernel (grad_ucta.cuh): closed systems whose matrices fit seven shared-memory slots twice per SM
bool grad_ucta_supported(int D) { return D > 16 && round8(D) <= 32; }

namespace {
template <int TM, int TN, int DPT, int KST>
int launch_ucta_t(const GradUParams& gp, size_t smem, cudaStream_t st) {
    constexpr int NT = 256;
    auto kern = grad_unitary_cta_kernel<TM, TN, DPT, KST, NT>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = 2LL * num_sms();
    const long long units = (long long)gp.B * gp.Q;
    if (grid > units) grid = units;
    kern<<<(int)grid, NT, smem, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}
}  // namespace

int launch_grad_ucta(GradUParams gp, cudaStream_t st) {
    gp.DP = round8(gp.D);
    if (gp.DP == 32) {
        gp.LD = 32;
        const size_t smem = (size_t)kGradUSlots * 32 * 32 * sizeof(cplx);
        return gp.D <= 28 ? launch_ucta_t<1, 2, 32, 7>(gp, smem, st) : launch_ucta_t<1, 2, 32, 8>(gp, smem, st);
    }
    gp.LD = gp.DP + 4;
    const size_t smem = (size_t)kGradUSlots * gp.DP * gp.LD * sizeof(cplx);
    return launch_ucta_t<1, 2, 0, 0>(gp, smem, st);
}

size_t grad_cta_workspace_bytes(int Bc, int N, int D) {
    if (grad_cta_uses_smem(D)) return 0;
    const size_t PP = (size_t)round8(D) * round8(D) * sizeof(cplx);
    const long long fgrid = grad_cta_frechet_grid(D, (long long)Bc * N);
    const size_t sweep = (size_t)Bc * kGradSweepSlots * PP, fre = (size_t)fgrid * kGradFrechetSlots * PP;
    return align_up(sweep > fre ? sweep : fre);
}

int grad_cta_frechet_grid(int D, long long units) {
    long long g = (long long)num_sms() * (grad_cta_uses_smem(D) ? 1 : 2);
    if (g > units) g = units;
    return (int)(g < 1 ? 1 : g);
}

// backward + forward sweeps (Psi, then M over it) and the Frechet / contraction kernel for one batch chunk
int launch_grad_cta(const cplx* G, const double* RS, const cplx* TR, const double* sig, const cplx* dUs, const cplx* Ubar, cplx* PsiM,
                    double* alpha, double* grad, int nb, int K, int N, int D, cplx* ws, cudaStream_t st) {
    GradCtaParams gp{};
    gp.G = G; gp.RS = RS; gp.TR = TR; gp.signals = sig; gp.dUs = dUs; gp.Ubar = Ubar; gp.PsiM = PsiM; gp.alpha = alpha; gp.grad = grad;
    gp.B = nb; gp.K = K; gp.N = N; gp.D = D; gp.DP = round8(D);
    const bool smem_path = grad_cta_uses_smem(D);
    gp.LD = smem_path ? gp.DP + 4 : gp.DP;
    gp.ws = smem_path ? nullptr : ws;
    const size_t PP = (size_t)gp.DP * gp.LD * sizeof(cplx);
    const size_t sm_sweep = smem_path ? kGradSweepSlots * PP : 0, sm_fre = smem_path ? kGradFrechetSlots * PP : 0;
    const int fgrid = grad_cta_frechet_grid(D, (long long)nb * N);
    int rc;
    if (gp.DP <= 48) {
        if ((rc = launch_sweeps_t<1, 2>(gp, sm_sweep, nb, st)) != 0) return rc;
        return launch_frechet_t<1, 2>(gp, sm_fre, fgrid, st);
    }
    if (gp.DP == 88) {
        if ((rc = launch_sweeps_t<3, 2>(gp, sm_sweep, nb, st)) != 0) return rc;
        return launch_frechet_t<3, 2>(gp, sm_fre, fgrid, st);
    }
    if ((rc = launch_sweeps_t<2, 2>(gp, sm_sweep, nb, st)) != 0) return rc;
    return launch_frechet_t<2, 2>(gp, sm_fre, fgrid, st);
}

}  // namespace c3b
