// Lane-group kernels for d <= 12 (other than 9): fused PWC propagators, segment fold, gate-sequence products.
#include "c3b_host.cuh"
#include "pwc_blk.cuh"

namespace c3b {

namespace {

template <int D, int BS, int WARPS, int MINB>
int launch_blk_t18_t(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    using L = BlkLayout<D, BS>;
    const size_t smem = L::smem_bytes(rp.K, WARPS);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d lane-group kernel", rp.K, rp.d);
    auto kern = pwc_blk_taylor_kernel<D, BS, WARPS, MINB>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const long long units = (long long)rp.B * rp.S;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > MINB) per_sm = MINB;
    long long grid = (long long)num_sms() * per_sm;
    const long long need = (units + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(int)grid, WARPS * 32, smem, st>>>(rp, counter);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

template <int D, int BS>
int launch_fold_blk_t(const ProductParams& pp, cudaStream_t st) {
    using L = BlkLayout<D, BS>;
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * L::WARP_ELEMS * sizeof(cplx);
    auto kern = fold_blk_kernel<D, BS, WARPS>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long wunits = ((long long)pp.B + L::MPW - 1) / L::MPW;
    long long grid = (wunits + WARPS - 1) / WARPS;
    const long long cap = (long long)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(int)grid, WARPS * 32, smem, st>>>(pp.mats, pp.B, pp.M, pp.D, pp.out);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

template <int D, int BS>
int launch_seq_blk_t(const cplx* gates, int Gn, const int* idx, const int* lens, int S, int Lmax, int d, cplx* out, cudaStream_t st) {
    using L = BlkLayout<D, BS>;
    constexpr int WARPS = 4;
    const size_t smem = ((size_t)Gn * L::BUF + (size_t)WARPS * L::WARP_ELEMS) * sizeof(cplx);
    auto kern = seq_product_blk_kernel<D, BS, WARPS>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long wunits = ((long long)S + L::MPW - 1) / L::MPW;
    long long grid = (wunits + WARPS - 1) / WARPS;
    const long long cap = (long long)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(int)grid, WARPS * 32, smem, st>>>(gates, Gn, idx, lens, S, Lmax, d, out);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

}  // namespace

int launch_small(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    if (rp.gate != nullptr) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gated launch is built for the d = 9 kernel only");
    switch (blk_template_dim(rp.d)) {
        case 2: return launch_blk_t18_t<2, 2, 4, 3>(rp, counter, st);
        case 3: return launch_blk_t18_t<3, 3, 4, 2>(rp, counter, st);
        case 4: return launch_blk_t18_t<4, 2, 4, 3>(rp, counter, st);
        case 6: return launch_blk_t18_t<6, 3, 4, 2>(rp, counter, st);
        case 8: return launch_blk_t18_t<8, 2, 4, 3>(rp, counter, st);
        case 12: return launch_blk_t18_t<12, 3, 4, 2>(rp, counter, st);
    }
    return fail(C3B_EUNSUPPORTED, "C3:ERROR: no lane-group kernel for d=%d", rp.d);
}

int launch_fold_small(const ProductParams& pp, cudaStream_t st) {
    switch (blk_template_dim(pp.D)) {
        case 2: return launch_fold_blk_t<2, 2>(pp, st);
        case 3: return launch_fold_blk_t<3, 3>(pp, st);
        case 4: return launch_fold_blk_t<4, 2>(pp, st);
        case 6: return launch_fold_blk_t<6, 3>(pp, st);
        case 8: return launch_fold_blk_t<8, 2>(pp, st);
        case 9: return launch_fold_blk_t<9, 3>(pp, st);
        case 12: return launch_fold_blk_t<12, 3>(pp, st);
    }
    return -1;
}

// lane-group kernel for small dimensions when the zero-padded gate table fits in shared memory; -1 = not applicable
int launch_seq_small(const cplx* gates, int Gn, const int* idx, const int* lens, int S, int Lmax, int d, cplx* out, cudaStream_t st) {
    const int TD = blk_template_dim(d);
    if (TD == 0 || Lmax <= 0) return -1;
    if ((size_t)Gn * (TD + 2) * TD * sizeof(cplx) > (size_t)96 * 1024) return -1;
    switch (TD) {
        case 2: return launch_seq_blk_t<2, 2>(gates, Gn, idx, lens, S, Lmax, d, out, st);
        case 3: return launch_seq_blk_t<3, 3>(gates, Gn, idx, lens, S, Lmax, d, out, st);
        case 4: return launch_seq_blk_t<4, 2>(gates, Gn, idx, lens, S, Lmax, d, out, st);
        case 6: return launch_seq_blk_t<6, 3>(gates, Gn, idx, lens, S, Lmax, d, out, st);
        case 8: return launch_seq_blk_t<8, 2>(gates, Gn, idx, lens, S, Lmax, d, out, st);
        case 9: return launch_seq_blk_t<9, 3>(gates, Gn, idx, lens, S, Lmax, d, out, st);
        case 12: return launch_seq_blk_t<12, 3>(gates, Gn, idx, lens, S, Lmax, d, out, st);
    }
    return -1;
}

}  // namespace c3b
