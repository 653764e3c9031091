// Taylor (degree 15+, four products) PWC propagators on fp64 tensor-core (DMMA) tiles, one CTA per (batch row, segment): d > 12, Lindblad
// superoperators (D = d^2), per-sample models.
#include "c3b_host.cuh"
#include "pwc_gemm.cuh"

namespace c3b {

namespace {

template <int TM, int TN, int DPT = 0, int KST = 0, int NT = kCtaThreads>
int launch_gemm_t(const GemmParams& gp, int grid, cudaStream_t st) {
    auto kern = pwc_taylor_cta_kernel<TM, TN, DPT, KST, NT>;
    size_t smem = gp.c.use_smem ? gemm_smem_bytes(gp.c.D, gp.g_in_smem ? gp.c.K : -1, gp.g_in_smem ? 0 : 1) : 0;
    if (TM == 0) smem = (size_t)3 * (12 * gp.DP + 8 * (gp.DP + 2)) * sizeof(cplx);      // panel ring of cta_zgemm_rows
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace

int launch_gemm(const CtaParams& cp, const cplx* TR, const double* RS, int grid, cudaStream_t st) {
    const Tuning& tn = tuning();
    GemmParams gp{};
    gp.c = cp; gp.TR = TR; gp.DP = round8(cp.D);
    gp.RS = (tn.norm_bound && TR != nullptr) ? RS : nullptr;   // RS are the row sums of the SHIFTED generators
    gp.LD = cp.use_smem ? gemm_ld_smem(cp.D) : gp.DP;
    gp.g_in_smem = (cp.use_smem && cp.hlist == nullptr && cp.G != nullptr && gemm_g_in_smem(cp.D, cp.K, cp.model_stride != 0)) ? 1 : 0;
    if (gp.DP <= 16) return launch_gemm_t<1, 1>(gp, grid, st);
    if (gp.DP == 32 && cp.use_smem) {
        if (tn.cta_threads == 512) return cp.D <= 28 ? launch_gemm_t<1, 1, 32, 7, 512>(gp, grid, st) : launch_gemm_t<1, 1, 32, 8, 512>(gp, grid, st);
        return cp.D <= 28 ? launch_gemm_t<1, 2, 32, 7>(gp, grid, st) : launch_gemm_t<1, 2, 32, 8>(gp, grid, st);
    }
    if (gp.DP <= 48) return launch_gemm_t<1, 2>(gp, grid, st);
    // 11 x 11 blocks at D = 81: 3 x 2 macro tiles give 24 tiles = 3 full rounds of the 8 warps (144 block slots for 121
    // blocks) where 2 x 2 gives 36 tiles = 5 rounds (160 slots): 20.5 -> 19.9 ms on the 296 x 40 probe (2 x 3: 20.5, 3 x 3: 23.9)
    if (gp.DP == 88 && tn.gemm_big == 0) return launch_gemm_t<3, 2>(gp, grid, st);
    // one 384-thread CTA per SM: 24 macro tiles = 2 full rounds of 12 warps, and 148 x 6 x 124 KB = 110 MB of workspace stays
    // inside the 126 MB L2 (two 256-thread CTAs per SM: 220 MB, L2 hit rate 52 %)
    if (gp.DP == 88 && tn.gemm_big == 2) return launch_gemm_t<3, 2, 0, 0, 384>(gp, grid, st);
    if (gp.DP == 88 && tn.gemm_big == 3) return launch_gemm_t<2, 2, 0, 0, 512>(gp, grid, st);   // 16 warps, one CTA per SM
    // Measured alternatives, not default: operand panels staged through shared memory by cp.async (cta_zgemm_rows), one CTA
    // per SM.  5: 12 warps, a warp owns a block-row (66 accumulators, no spills): 0.73 against 0.79 for the default -- the
    // fragment loads no longer wait on L2 (long_scoreboard 14.8 -> 4.2 per issue) but one CTA per SM exposes the per-panel
    // barrier (15 % of the samples) and leaves 3 warps per scheduler.  4: 22 warps, half a block-row each: 80 registers per
    // thread spill the accumulators inside the panel loop: 0.55.  (Two 6-warp CTAs per SM with two block-rows per warp need
    // 132 accumulators: 1.8 KB of spills, not kept.)
    if (gp.DP == 88 && !cp.use_smem && tn.gemm_big == 4) return launch_gemm_t<0, 11, 0, 0, 704>(gp, grid, st);
    if (gp.DP == 88 && !cp.use_smem && tn.gemm_big == 5) return launch_gemm_t<0, 11, 0, 0, 384>(gp, grid, st);
    return launch_gemm_t<2, 2>(gp, grid, st);
}

}  // namespace c3b
