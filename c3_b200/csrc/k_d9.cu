// d = 9 (and zero-padded d = 7): the headline lane-group kernels.
#include "c3b_host.cuh"
#include "pwc_blk.cuh"
#include "pwc_blk9.cuh"
#include "pwc_shfl9.cuh"
#include "pwc_dmma9.cuh"

namespace c3b {

namespace {

template <typename Kern>
int launch_persistent(Kern kern, size_t smem, int warps, int minb, const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d lane-group kernel", rp.K, rp.d);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const long long units = (long long)rp.B * rp.S;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > minb) per_sm = minb;
    long long grid = (long long)num_sms() * per_sm;
    const long long need = (units + warps - 1) / warps;
    if (grid > need) grid = need;
    kern<<<(int)grid, warps * 32, smem, st>>>(rp, counter);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace

bool d9_gated_supported(int variant) { return variant >= 1 && variant <= 3; }

int launch_d9(const RowsParams& rp, unsigned int* counter, int variant, cudaStream_t st) {
    if (rp.gate != nullptr && !d9_gated_supported(variant))
        return fail(C3B_EUNSUPPORTED, "C3:ERROR: gated launch is not built for d9_variant %d", variant);
    if (variant == 0)
        return launch_persistent(pwc_blk_taylor_kernel<9, 3, 4, 2>, BlkLayout<9, 3>::smem_bytes(rp.K, 4), 4, 2, rp, counter, st);
    // gated launches: rows on their own 128-byte lines may be read through L1 (see load_signal)
    const bool lines = rp.gate != nullptr && ((size_t)rp.K * rp.N * sizeof(double)) % 128 == 0 &&
                       (reinterpret_cast<uintptr_t>(rp.signals) % 128) == 0;
    if (variant == 3) {
        const size_t smem = Dmma9::smem_bytes(rp.K, 8);
        if (rp.gate != nullptr)
            return lines ? launch_persistent(pwc_dmma9_kernel<8, 2, 2>, smem, 8, 2, rp, counter, st)
                         : launch_persistent(pwc_dmma9_kernel<8, 2, 1>, smem, 8, 2, rp, counter, st);
        return launch_persistent(pwc_dmma9_kernel<8, 2, 0>, smem, 8, 2, rp, counter, st);
    }
    if (variant == 2) {
        const size_t smem8 = Shfl9::smem_bytes(rp.K, 8);
        if (rp.gate != nullptr)
            return lines ? launch_persistent(pwc_shfl9_kernel<8, 1, 2>, smem8, 8, 1, rp, counter, st)
                         : launch_persistent(pwc_shfl9_kernel<8, 1, 1>, smem8, 8, 1, rp, counter, st);
        return launch_persistent(pwc_shfl9_kernel<8, 1, 0>, smem8, 8, 1, rp, counter, st);
    }
    const size_t smem = Blk9T<true>::smem_bytes(rp.K, 4);
    if (rp.gate != nullptr)
        return lines ? launch_persistent(pwc_blk9_taylor_kernel<4, 2, true, 2>, smem, 4, 2, rp, counter, st)
                     : launch_persistent(pwc_blk9_taylor_kernel<4, 2, true, 1>, smem, 4, 2, rp, counter, st);
    return launch_persistent(pwc_blk9_taylor_kernel<4, 2, true, 0>, smem, 4, 2, rp, counter, st);
}

}  // namespace c3b
