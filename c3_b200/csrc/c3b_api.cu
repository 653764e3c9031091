// C ABI of libc3b200.so -- see include/c3b200.h for the contract of every entry point.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a.
#include "../../include/c3b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>

#include "c3b_common.cuh"
#include "pwc_cta.cuh"
#include "pwc_rows.cuh"
#include "pwc_blk.cuh"
#include "pwc_blk9.cuh"
#include "product.cuh"
#include "pwc_gemm.cuh"
#include "grad.cuh"
#include "fidelity.cuh"
#include "signal_chain.cuh"
#include "dressing.cuh"
#include "peak.cuh"

using namespace c3b;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return fail(C3B_ECUDA, "C3:ERROR: %s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

std::atomic<long long> g_launches{0};  // kernels launched by this library (bench.py's gpu_launches)
long long g_profile = 0;               // when set, the main PWC kernel of each call is bracketed by events
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
bool g_ev_valid = false;

// ---- tuning (process-wide) -------------------------------------------------------------------
long long g_target_units = 32768;  // rows kernel: aim for this many warp work units per launch
long long g_gemm_big = getenv("C3B_GEMM_BIG") ? atoll(getenv("C3B_GEMM_BIG")) : 0;   // DMMA CTA kernel, DP = 88 (D = 81): macro-tile shape
long long g_seq_variant = 1;       // evaluate_sequences: 1 = lane-group kernel for small d, 0 = CTA-per-sequence product kernel
long long g_norm_bound = 1;        // DMMA CTA kernel: scaling from the row-sum bound (1) or from the exact inf-norm of every slice (0)
long long g_force_cta = 0;         // route everything to the CTA kernel (testing)
long long g_cta_variant = getenv("C3B_CTA_VARIANT") ? atoll(getenv("C3B_CTA_VARIANT")) : 1;  // 0: Pade + pivoted Gauss-Jordan, 1: Taylor-18 on DMMA tiles
long long g_cta_threads = getenv("C3B_CTA_THREADS") ? atoll(getenv("C3B_CTA_THREADS")) : 512;   // DMMA CTA kernel, DP = 32: 256 or 512 threads
long long g_grad_variant = getenv("C3B_GRAD_VARIANT") ? atoll(getenv("C3B_GRAD_VARIANT")) : 1;   // 1: Frechet of the Taylor scheme (d <= 16), 0: augmented exponential
long long g_min_chunk = getenv("C3B_MIN_CHUNK") ? atoll(getenv("C3B_MIN_CHUNK")) : 8;         // rows kernel: minimum slices per lane group
// 1: rows v2 (one row per lane, Pade) | 4-6: rows v3 (experimental) | 7-12: block layout, Pade + Gauss-Jordan (d=9)
// 13 (default): block layout, degree-18 Taylor, trace shift, all d <= 12
long long g_rows_variant = getenv("C3B_ROWS_VARIANT") ? atoll(getenv("C3B_ROWS_VARIANT")) : 16;      // 0: v1 kernel, 1: v2 (255 regs), 2: v2 V-in-smem (168 regs), 3: v2 (168 regs)

int num_sms() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
        cached = n;
    else {
        cudaGetLastError();
        return 148;  // B200; used only for sizing when no device is visible
    }
    return cached;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

constexpr int kRowsWarps = 4;
const int kRowsDims[] = {2, 3, 4, 5, 6, 8, 9, 10, 12};

int rows_template_dim(int D) {
    for (int t : kRowsDims)
        if (D <= t) return t;
    return 0;
}

// Block-kernel instantiations: padded dimension -> (D, BS).  Every d <= 12 maps to one of them.
int blk_template_dim(int d) {
    const int dims[] = {2, 3, 4, 6, 8, 9, 12};
    for (int t : dims)
        if (d <= t) return t;
    return 0;
}

int pwc_path(int K, int D, int batched_model) {
    (void)K;
    if (!g_force_cta && !batched_model && rows_template_dim(D) != 0) return 1;
    return D <= 32 ? 2 : 3;
}

size_t cta_smem_bytes(int D) { return (((size_t)D * sizeof(int) + 15) & ~(size_t)15) + (size_t)kCtaSlots * D * D * sizeof(cplx); }

int round8(int D) { return (D + 7) & ~7; }
size_t gemm_mats_bytes(int D) { return (size_t)kGemmSlots * round8(D) * (round8(D) + 4) * sizeof(cplx); }
// the shared-model generators ride along in shared memory when they fit next to the matrix slots
bool gemm_g_in_smem(int D, int K, int batched_model) {
    return !batched_model && gemm_mats_bytes(D) + (size_t)(K + 1) * D * D * sizeof(cplx) <= (size_t)220 * 1024;
}
size_t gemm_smem_bytes(int D, int K = -1, int batched_model = 1) {
    return gemm_mats_bytes(D) + ((K >= 0 && gemm_g_in_smem(D, K, batched_model)) ? (size_t)(K + 1) * D * D * sizeof(cplx) : 0);
}

int cta_grid(int D, long long units) {
    int per_sm = 1;
    if (D <= 32) {
        const size_t sm = g_cta_variant == 1 ? gemm_smem_bytes(D) : cta_smem_bytes(D);
        per_sm = (int)((size_t)220 * 1024 / (sm + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
    } else {
        per_sm = 2;
    }
    long long g = (long long)num_sms() * per_sm;
    if (g > units) g = units;
    if (g < 1) g = 1;
    return (int)g;
}

struct Plan {
    int path;      // 1 rows, 2 cta smem, 3 cta global
    int S;         // segments per batch element
    int seg_len;
    int grid;      // CTA kernel grid (persistent)
    size_t off_G, off_RS, off_TR, off_seg, off_cta, off_prod, off_counter, total;
};

Plan make_plan(int B, int K, int N, int D, int batched_model, bool hlist) {
    Plan pl{};
    pl.path = pwc_path(K, D, batched_model);
    const int Bm = batched_model ? B : 1;
    if (pl.path == 1) {
        const int TD = rows_template_dim(D);
        int G = 32 / TD;
        if (g_rows_variant >= 13) {
            switch (blk_template_dim(D)) {
                case 2: case 3: G = 32; break;
                case 4: case 6: G = 8; break;
                case 8: case 12: G = 2; break;
                default: G = 3; break;
            }
        }
        if (g_rows_variant >= 4 && g_rows_variant <= 6 && (TD == 9 || TD == 3)) G = (TD == 9) ? (g_rows_variant >= 5 ? 6 : 10) : 32;   // v3: lane groups per warp
        long long S = (g_target_units + B - 1) / B;
        long long smax = N / (g_min_chunk * G);
        if (smax < 1) smax = 1;
        if (S > smax) S = smax;
        if (S < 1) S = 1;
        pl.S = (int)S;
    } else {
        const long long want = 4LL * num_sms() * (D <= 32 ? 2 : 2);
        long long S = (want + B - 1) / B;
        long long smax = N / 8;
        if (smax < 1) smax = 1;
        if (S > smax) S = smax;
        if (S < 1) S = 1;
        pl.S = (int)S;
    }
    pl.seg_len = (N + pl.S - 1) / pl.S;
    pl.S = (N + pl.seg_len - 1) / pl.seg_len;  // drop empty trailing segments
    if (pl.S < 1) pl.S = 1;
    pl.grid = cta_grid(D, (long long)B * pl.S);
    size_t off = 0;
    pl.off_G = off;
    if (!hlist) off += align_up((size_t)Bm * (K + 1) * D * D * sizeof(cplx));
    pl.off_RS = off;
    if (!hlist) off += align_up((size_t)Bm * (K + 1) * D * sizeof(double));
    pl.off_TR = off;
    if (!hlist) off += align_up((size_t)Bm * (K + 1) * sizeof(cplx));
    pl.off_seg = off;
    if (pl.S > 1) off += align_up((size_t)B * pl.S * D * D * sizeof(cplx));
    pl.off_cta = off;
    if (pl.path == 3) off += align_up((size_t)pl.grid * kGemmSlots * round8(D) * round8(D) * sizeof(cplx));
    pl.off_prod = off;
    if (pl.S > 1 && D > 64) off += align_up((size_t)cta_grid(D, 1LL << 40) * 2 * D * D * sizeof(cplx));
    pl.off_counter = off;
    off += 256;
    pl.total = off < 256 ? 256 : off;
    return pl;
}

// ---- launches ----------------------------------------------------------------------------------

template <int D, int MINB>
int launch_rows_t(const RowsParams& rp, cudaStream_t st) {
    using L = RowsLayout<D, kRowsWarps>;
    const size_t smem = L::smem_bytes(rp.K);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d register kernel", rp.K, rp.d);
    auto kern = pwc_rows_kernel<D, kRowsWarps, MINB>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long units = (long long)rp.B * rp.S;
    const int grid = (int)((units + kRowsWarps - 1) / kRowsWarps);
    kern<<<grid, kRowsWarps * 32, smem, st>>>(rp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

template <int D, int MINB, bool VSMEM>
int launch_rows2_t(const RowsParams& rp, cudaStream_t st) {
    using L = Rows2Layout<D, kRowsWarps, VSMEM>;
    const size_t smem = L::smem_bytes(rp.K);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d register kernel", rp.K, rp.d);
    auto kern = pwc_rows2_kernel<D, kRowsWarps, MINB, VSMEM>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long units = (long long)rp.B * rp.S;
    const int grid = (int)((units + kRowsWarps - 1) / kRowsWarps);
    kern<<<grid, kRowsWarps * 32, smem, st>>>(rp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

template <int D, int R, int WARPS>
int launch_rows3_t(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    using L = Rows3Layout<D, R>;
    const size_t smem = L::smem_bytes(rp.K, WARPS);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d register kernel", rp.K, rp.d);
    auto kern = pwc_rows3_kernel<D, R, WARPS>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const long long units = (long long)rp.B * rp.S;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)num_sms() * per_sm;
    const long long need = (units + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(int)grid, WARPS * 32, smem, st>>>(rp, counter);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

template <int D, int BS, int WARPS, int MINB>
int launch_blk_t(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    using L = BlkLayout<D, BS>;
    const size_t smem = L::smem_bytes(rp.K, WARPS);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d block kernel", rp.K, rp.d);
    auto kern = pwc_blk_kernel<D, BS, WARPS, MINB>;
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

template <int D, int BS, int WARPS, int MINB>
int launch_blk_t18_t(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    using L = BlkLayout<D, BS>;
    const size_t smem = L::smem_bytes(rp.K, WARPS);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d block kernel", rp.K, rp.d);
    auto kern = pwc_blk_t18_kernel<D, BS, WARPS, MINB>;
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

template <int WARPS, int MINB, bool NOSEL, bool GATED = false>
int launch_blk9_t18_t(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    const size_t smem = Blk9T<NOSEL>::smem_bytes(rp.K, WARPS);
    if (smem > 227 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: too many control lines (K=%d) for the d=%d block kernel", rp.K, rp.d);
    auto kern = pwc_blk9_t18_kernel<WARPS, MINB, NOSEL, GATED>;
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

// does the kernel that launch_rows() would pick accept trace-shifted generators?
bool rows_kernel_takes_shift(int d) { return g_rows_variant >= 13 && blk_template_dim(d) != 0; }

int launch_rows(const RowsParams& rp, unsigned int* counter, cudaStream_t st) {
    if (g_rows_variant >= 13) {
        switch (blk_template_dim(rp.d)) {
            case 2: return launch_blk_t18_t<2, 2, 4, 3>(rp, counter, st);
            case 3: return launch_blk_t18_t<3, 3, 4, 2>(rp, counter, st);
            case 4: return launch_blk_t18_t<4, 2, 4, 3>(rp, counter, st);
            case 6: return launch_blk_t18_t<6, 3, 4, 2>(rp, counter, st);
            case 8: return launch_blk_t18_t<8, 2, 4, 3>(rp, counter, st);
            case 9: if (g_rows_variant == 15) return launch_blk9_t18_t<4, 2, false>(rp, counter, st);
                    if (g_rows_variant == 16) return rp.rows_ready ? launch_blk9_t18_t<4, 2, true, true>(rp, counter, st)
                                                                   : launch_blk9_t18_t<4, 2, true>(rp, counter, st);
                    return (g_rows_variant == 14) ? launch_blk_t18_t<9, 3, 4, 3>(rp, counter, st)
                                                  : launch_blk_t18_t<9, 3, 4, 2>(rp, counter, st);
            case 12: return launch_blk_t18_t<12, 3, 4, 2>(rp, counter, st);
        }
    }
    if (rows_template_dim(rp.d) == 9 && g_rows_variant == 7) return launch_blk_t<9, 3, 4, 3>(rp, counter, st);
    if (rows_template_dim(rp.d) == 9 && g_rows_variant == 8) return launch_blk_t<9, 3, 4, 2>(rp, counter, st);
    if (rows_template_dim(rp.d) == 9 && g_rows_variant == 9) return launch_blk_t<9, 3, 6, 2>(rp, counter, st);
    if (rows_template_dim(rp.d) == 9 && g_rows_variant == 10) return launch_blk_t<9, 3, 5, 2>(rp, counter, st);
    if (rows_template_dim(rp.d) == 9 && g_rows_variant == 11) return launch_blk_t<9, 3, 9, 1>(rp, counter, st);
    if (rows_template_dim(rp.d) == 9 && g_rows_variant == 12) return launch_blk_t<9, 3, 11, 1>(rp, counter, st);
    if (g_rows_variant >= 4 && g_rows_variant <= 6) {
        if (rows_template_dim(rp.d) == 9 && g_rows_variant == 5) return launch_rows3_t<9, 2, 7>(rp, counter, st);
        if (rows_template_dim(rp.d) == 9 && g_rows_variant == 6) return launch_rows3_t<9, 2, 6>(rp, counter, st);
        if (rows_template_dim(rp.d) == 9) return launch_rows3_t<9, 3, 4>(rp, counter, st);
        if (rows_template_dim(rp.d) == 3) return launch_rows3_t<3, 3, 4>(rp, counter, st);
    }
    if (rows_template_dim(rp.d) == 9 && g_rows_variant > 0) {
        if (g_rows_variant == 1) return launch_rows2_t<9, 2, false>(rp, st);
        if (g_rows_variant == 2) return launch_rows2_t<9, 3, true>(rp, st);
        return launch_rows2_t<9, 3, false>(rp, st);
    }
    switch (rows_template_dim(rp.d)) {
        case 2: return launch_rows_t<2, 4>(rp, st);
        case 3: return launch_rows_t<3, 4>(rp, st);
        case 4: return launch_rows_t<4, 4>(rp, st);
        case 5: return launch_rows_t<5, 4>(rp, st);
        case 6: return launch_rows_t<6, 3>(rp, st);
        case 8: return launch_rows_t<8, 3>(rp, st);
        case 9: return launch_rows_t<9, 3>(rp, st);
        case 10: return launch_rows_t<10, 2>(rp, st);
        case 12: return launch_rows_t<12, 2>(rp, st);
    }
    return fail(C3B_EUNSUPPORTED, "C3:ERROR: no register kernel for d=%d", rp.d);
}

template <int CT, int TR, int TC>
int launch_cta_t(const CtaParams& cp, int grid, cudaStream_t st) {
    auto kern = pwc_cta_kernel<CT, TR, TC>;
    size_t smem = cp.use_smem ? cta_smem_bytes(cp.D) : (((size_t)cp.D * sizeof(int) + 15) & ~(size_t)15);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kCtaThreads, smem, st>>>(cp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

template <int TM, int TN, int DPT = 0, int KST = 0, int NT = kCtaThreads>
int launch_gemm_t(const GemmParams& gp, int grid, cudaStream_t st) {
    auto kern = pwc_t18_cta_kernel<TM, TN, DPT, KST, NT>;
    const size_t smem = gp.c.use_smem ? gemm_smem_bytes(gp.c.D, gp.g_in_smem ? gp.c.K : -1, gp.g_in_smem ? 0 : 1) : 0;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

int launch_gemm(const CtaParams& cp, const cplx* TR, const double* RS, int grid, cudaStream_t st) {
    GemmParams gp{};
    gp.c = cp; gp.TR = TR; gp.DP = round8(cp.D);
    gp.RS = (g_norm_bound && TR != nullptr) ? RS : nullptr;   // RS are the row sums of the SHIFTED generators
    gp.LD = cp.use_smem ? gp.DP + 4 : gp.DP;
    gp.g_in_smem = (cp.use_smem && cp.hlist == nullptr && cp.G != nullptr && gemm_g_in_smem(cp.D, cp.K, cp.model_stride != 0)) ? 1 : 0;
    if (gp.DP <= 16) return launch_gemm_t<1, 1>(gp, grid, st);
    if (gp.DP == 32 && cp.use_smem) {
        if (g_cta_threads == 512) return cp.D <= 28 ? launch_gemm_t<1, 1, 32, 7, 512>(gp, grid, st) : launch_gemm_t<1, 1, 32, 8, 512>(gp, grid, st);
        return cp.D <= 28 ? launch_gemm_t<1, 2, 32, 7>(gp, grid, st) : launch_gemm_t<1, 2, 32, 8>(gp, grid, st);
    }
    if (gp.DP <= 48) return launch_gemm_t<1, 2>(gp, grid, st);
    // 11 x 11 blocks at D = 81: 3 x 2 macro tiles give 24 tiles = 3 full rounds of the 8 warps (144 block slots for 121
    // blocks) where 2 x 2 gives 36 tiles = 5 rounds (160 slots): 20.5 -> 19.9 ms on the 296 x 40 probe (2 x 3: 20.5, 3 x 3: 23.9)
    if (gp.DP == 88 && g_gemm_big == 0) return launch_gemm_t<3, 2>(gp, grid, st);
    return launch_gemm_t<2, 2>(gp, grid, st);
}

int launch_cta(const CtaParams& cp, int grid, cudaStream_t st) {
    const int D = cp.D;
    if (D <= 16) return launch_cta_t<16, 1, 1>(cp, grid, st);
    if (D <= 32) return launch_cta_t<32, 4, 1>(cp, grid, st);
    if (D <= 64) return launch_cta_t<32, 4, 2>(cp, grid, st);
    return launch_cta_t<32, 4, 3>(cp, grid, st);
}

template <int CT, int TR, int TC>
int launch_product_t(const ProductParams& pp, int grid, cudaStream_t st) {
    auto kern = product_kernel<CT, TR, TC>;
    const size_t smem = pp.use_smem ? (size_t)2 * pp.D * pp.D * sizeof(cplx) : 0;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kCtaThreads, smem, st>>>(pp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

int launch_product(ProductParams pp, cudaStream_t st) {
    const int D = pp.D;
    if (g_seq_variant != 0 && pp.idx == nullptr && pp.lens == nullptr && pp.S == 1 && pp.seg_len >= pp.M && pp.M >= 1 &&
        (long long)pp.B * 2 >= num_sms()) {      // lane-group fold for small d (a few rows only: the CTA kernel has less latency)
        switch (blk_template_dim(D)) {
            case 2: return launch_fold_blk_t<2, 2>(pp, st);
            case 3: return launch_fold_blk_t<3, 3>(pp, st);
            case 4: return launch_fold_blk_t<4, 2>(pp, st);
            case 6: return launch_fold_blk_t<6, 3>(pp, st);
            case 8: return launch_fold_blk_t<8, 2>(pp, st);
            case 9: return launch_fold_blk_t<9, 3>(pp, st);
            case 12: return launch_fold_blk_t<12, 3>(pp, st);
        }
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

thread_local const unsigned int* t_rows_ready = nullptr;   // set by c3b_pwc_closed_gated around its call of c3b_pwc_closed

// common tail of the three pwc entry points once G (or the H list) is in place
int run_pwc(const Plan& pl, const cplx* G, const double* RS, const cplx* TR, const double* signals, const cplx* hlist, double dt,
            int B, int K, int N, int D, int batched_model, cplx* U_out, cplx* dUs_out, char* ws, cudaStream_t st) {
    cplx* seg = pl.S > 1 ? reinterpret_cast<cplx*>(ws + pl.off_seg) : nullptr;
    if (g_profile) {
        if (g_ev0 == nullptr) { CUDA_TRY(cudaEventCreate(&g_ev0)); CUDA_TRY(cudaEventCreate(&g_ev1)); }
        CUDA_TRY(cudaEventRecord(g_ev0, st));
    }
    if (pl.path == 1) {
        RowsParams rp{};
        rp.G = G; rp.RS = RS; rp.TR = TR; rp.signals = signals; rp.hlist = hlist;
        rp.hscale_re = 0.0; rp.hscale_im = -dt;
        rp.model_stride = 0;
        rp.B = B; rp.K = K; rp.N = N; rp.d = D; rp.S = pl.S; rp.seg_len = pl.seg_len;
        rp.U_out = U_out; rp.seg_out = seg; rp.dUs_out = dUs_out;
        rp.rows_ready = t_rows_ready;
        if (t_rows_ready != nullptr && seg != nullptr)     // a row that never arrives must surface as NaN after the fold, too
            CUDA_TRY(cudaMemsetAsync(seg, 0xff, (size_t)B * pl.S * D * D * sizeof(cplx), st));
        if (t_rows_ready != nullptr && !(g_rows_variant == 16 && D == 9))
            return fail(C3B_EUNSUPPORTED, "C3:ERROR: gated launch is built for the d = 9 kernel only");
        int rc = launch_rows(rp, reinterpret_cast<unsigned int*>(ws + pl.off_counter), st);
        if (rc) return rc;
    } else {
        if (t_rows_ready != nullptr) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gated launch is built for the d = 9 kernel only");
        CtaParams cp{};
        cp.G = G; cp.signals = signals; cp.hlist = hlist;
        cp.hscale_re = 0.0; cp.hscale_im = -dt;
        cp.model_stride = batched_model ? (long long)(K + 1) * D * D : 0;
        cp.B = B; cp.K = K; cp.N = N; cp.D = D; cp.S = pl.S; cp.seg_len = pl.seg_len;
        cp.U_out = U_out; cp.seg_out = seg; cp.dUs_out = dUs_out;
        cp.ws = reinterpret_cast<cplx*>(ws + pl.off_cta);
        cp.use_smem = pl.path == 2;
        int rc = (g_cta_variant == 1) ? launch_gemm(cp, TR, RS, pl.grid, st) : launch_cta(cp, pl.grid, st);
        if (rc) return rc;
    }
    if (g_profile) { CUDA_TRY(cudaEventRecord(g_ev1, st)); g_ev_valid = true; }
    if (pl.S > 1) {
        ProductParams pp{};
        pp.mats = seg; pp.idx = nullptr; pp.lens = nullptr;
        pp.B = B; pp.M = pl.S; pp.D = D; pp.S = 1; pp.seg_len = pl.S;
        pp.out = U_out; pp.ws = reinterpret_cast<cplx*>(ws + pl.off_prod);
        int rc = launch_product(pp, st);
        if (rc) return rc;
    }
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
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

// lane-group kernel for small dimensions when the zero-padded gate table fits in shared memory; -1 = not applicable
int launch_seq_blk(const cplx* gates, int Gn, const int* idx, const int* lens, int S, int Lmax, int d, cplx* out, cudaStream_t st) {
    const int TD = blk_template_dim(d);
    if (TD == 0 || g_seq_variant == 0 || Lmax <= 0) return -1;
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

int check_common(int B, int K, int N, int d, const void* U_out, const void* ws) {
    if (B <= 0 || N <= 0 || d <= 0 || K < 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d N=%d d=%d)", B, K, N, d);
    if (U_out == nullptr) return fail(C3B_EINVAL, "C3:ERROR: U_out is NULL");
    if (ws == nullptr) return fail(C3B_EINVAL, "C3:ERROR: workspace is NULL");
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return fail(C3B_EINVAL, "C3:ERROR: workspace must be 16-byte aligned");
    return C3B_OK;
}

}  // namespace

extern "C" {

int c3b_version(void) { return 100; }

const char* c3b_last_error(void) { return g_err; }

int c3b_set_tuning(const char* key, long long value) {
    if (key == nullptr) return fail(C3B_EINVAL, "C3:ERROR: null tuning key");
    if (!strcmp(key, "target_units")) { g_target_units = value < 1 ? 1 : value; return C3B_OK; }
    if (!strcmp(key, "force_cta")) { g_force_cta = value; return C3B_OK; }
    if (!strcmp(key, "norm_bound")) { g_norm_bound = value; return C3B_OK; }
    if (!strcmp(key, "seq_variant")) { g_seq_variant = value; return C3B_OK; }
    if (!strcmp(key, "gemm_big")) { g_gemm_big = value; return C3B_OK; }
    if (!strcmp(key, "cta_variant")) { g_cta_variant = value; return C3B_OK; }
    if (!strcmp(key, "cta_threads")) { g_cta_threads = value; return C3B_OK; }
    if (!strcmp(key, "grad_variant")) { g_grad_variant = value; return C3B_OK; }
    if (!strcmp(key, "profile")) { g_profile = value; return C3B_OK; }
    if (!strcmp(key, "rows_variant")) { g_rows_variant = value; return C3B_OK; }
    if (!strcmp(key, "min_chunk")) { g_min_chunk = value < 1 ? 1 : value; return C3B_OK; }
    return fail(C3B_EINVAL, "C3:ERROR: unknown tuning key '%s'", key);
}

int c3b_pwc_path(int K, int D, int batched_model) { return pwc_path(K, D, batched_model); }

size_t c3b_pwc_workspace_bytes(int B, int K, int N, int d, int lindblad, int batched_model) {
    if (B <= 0 || N <= 0 || d <= 0 || K < 0) return 0;
    const int D = lindblad ? d * d : d;
    return make_plan(B, K, N, D, batched_model, false).total;
}

int c3b_pwc_closed(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                   int batched_model, void* U_out, void* dUs_out, void* workspace, size_t workspace_bytes,
                   void* stream) {
    int rc = check_common(B, K, N, d, U_out, workspace);
    if (rc) return rc;
    if (h0 == nullptr) return fail(C3B_EINVAL, "C3:ERROR: h0 is NULL");
    if (K > 0 && (hks == nullptr || signals == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: K=%d but hks/signals is NULL", K);
    const Plan pl = make_plan(B, K, N, d, batched_model, false);
    if (workspace_bytes < pl.total) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    cplx* G = reinterpret_cast<cplx*>(ws + pl.off_G);
    double* RS = reinterpret_cast<double*>(ws + pl.off_RS);
    cplx* TR = ((pl.path == 1 && rows_kernel_takes_shift(d)) || (pl.path != 1 && g_cta_variant == 1)) ? reinterpret_cast<cplx*>(ws + pl.off_TR) : nullptr;
    const int Bm = batched_model ? B : 1;
    {
        const long long total = (long long)Bm * (K + 1) * d * d;
        const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
        setup_closed_kernel<<<blocks, 256, 0, st>>>(static_cast<const cplx*>(h0), static_cast<const cplx*>(hks), G, Bm, K, d, dt);
        CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
        if (TR != nullptr) {
            const long long nmat = (long long)Bm * (K + 1);
            trace_shift_kernel<<<(int)((nmat * 32 + 255) / 256), 256, 0, st>>>(G, TR, nmat, d);
            CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        const long long nrows = (long long)Bm * (K + 1) * d;
        rowsum_kernel<<<(int)((nrows * 32 + 255) / 256), 256, 0, st>>>(G, RS, nrows, d);
        CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return run_pwc(pl, G, RS, TR, signals, nullptr, dt, B, K, N, d, batched_model, static_cast<cplx*>(U_out),
                   static_cast<cplx*>(dUs_out), ws, st);
}

int c3b_pwc_closed_gated(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                         void* U_out, const uint32_t* rows_ready, void* workspace, size_t workspace_bytes, void* stream) {
    if (rows_ready == nullptr) return fail(C3B_EINVAL, "C3:ERROR: rows_ready is NULL");
    if (K <= 0) return fail(C3B_EINVAL, "C3:ERROR: gated launch needs control fields (K > 0)");
    t_rows_ready = rows_ready;
    const int rc = c3b_pwc_closed(h0, hks, signals, dt, B, K, N, d, 0, U_out, nullptr, workspace, workspace_bytes, stream);
    t_rows_ready = nullptr;
    return rc;
}

int c3b_pwc_gated_supported(int d) { return (g_rows_variant == 16 && d == 9 && g_force_cta == 0) ? 1 : 0; }

int c3b_pwc_closed_hlist(const void* Hs, double dt, int B, int N, int d, void* U_out, void* dUs_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(B, 0, N, d, U_out, workspace);
    if (rc) return rc;
    if (Hs == nullptr) return fail(C3B_EINVAL, "C3:ERROR: Hs is NULL");
    const Plan pl = make_plan(B, 0, N, d, 0, true);
    if (workspace_bytes < pl.total) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, pl.total);
    return run_pwc(pl, nullptr, nullptr, nullptr, nullptr, static_cast<const cplx*>(Hs), dt, B, 0, N, d, 0,
                   static_cast<cplx*>(U_out), static_cast<cplx*>(dUs_out), static_cast<char*>(workspace),
                   static_cast<cudaStream_t>(stream));
}

int c3b_pwc_lindblad(const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                     int B, int K, int N, int d, int batched_model, void* U_out, void* dUs_out, void* workspace,
                     size_t workspace_bytes, void* stream) {
    int rc = check_common(B, K, N, d, U_out, workspace);
    if (rc) return rc;
    if (h0 == nullptr) return fail(C3B_EINVAL, "C3:ERROR: h0 is NULL");
    if (K > 0 && (hks == nullptr || signals == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: K=%d but hks/signals is NULL", K);
    if (C < 0 || (C > 0 && col_ops == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: C=%d but col_ops is NULL", C);
    if (d > 181) return fail(C3B_EUNSUPPORTED, "C3:ERROR: Lindblad d=%d too large", d);
    const int D = d * d;
    const Plan pl = make_plan(B, K, N, D, batched_model, false);
    if (workspace_bytes < pl.total) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    cplx* G = reinterpret_cast<cplx*>(ws + pl.off_G);
    double* RS = reinterpret_cast<double*>(ws + pl.off_RS);
    cplx* TR = ((pl.path == 1 && rows_kernel_takes_shift(D)) || (pl.path != 1 && g_cta_variant == 1)) ? reinterpret_cast<cplx*>(ws + pl.off_TR) : nullptr;
    const int Bm = batched_model ? B : 1;
    {
        const long long total = (long long)Bm * (K + 1) * D * D;
        const int blocks = (int)((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256);
        setup_lindblad_kernel<<<blocks, 256, 0, st>>>(static_cast<const cplx*>(h0), static_cast<const cplx*>(hks),
                                                      C > 0 ? static_cast<const cplx*>(col_ops) : nullptr, G, Bm, K, C, d, dt);
        CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
        if (TR != nullptr) {
            const long long nmat = (long long)Bm * (K + 1);
            trace_shift_kernel<<<(int)((nmat * 32 + 255) / 256), 256, 0, st>>>(G, TR, nmat, D);
            CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        const long long nrows = (long long)Bm * (K + 1) * D;
        rowsum_kernel<<<(int)((nrows * 32 + 255) / 256), 256, 0, st>>>(G, RS, nrows, D);
        CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return run_pwc(pl, G, RS, TR, signals, nullptr, dt, B, K, N, D, batched_model, static_cast<cplx*>(U_out),
                   static_cast<cplx*>(dUs_out), ws, st);
}

static size_t product_scratch_bytes(int D) {
    return D > 64 ? align_up((size_t)cta_grid(D, 1LL << 40) * 2 * D * D * sizeof(cplx)) : 0;
}

size_t c3b_product_workspace_bytes(int B, int M, int D) {
    if (B <= 0 || M < 0 || D <= 0) return 0;
    // [scratch for D > 64][level-1 segment results]
    size_t off = product_scratch_bytes(D) + align_up((size_t)B * ((M + 7) / 8 + 1) * D * D * sizeof(cplx));
    return off < 256 ? 256 : off;
}

int c3b_ordered_product(const void* mats, int B, int M, int D, void* out, void* workspace, size_t workspace_bytes,
                        void* stream) {
    if (B <= 0 || M <= 0 || D <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d M=%d D=%d)", B, M, D);
    if (mats == nullptr || out == nullptr || workspace == nullptr) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (workspace_bytes < c3b_product_workspace_bytes(B, M, D)) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    // choose segments so that the grid is filled when B is small
    long long S = (4LL * num_sms() + B - 1) / B;
    long long smax = (M + 7) / 8;
    if (S > smax) S = smax;
    if (S < 1) S = 1;
    int seg_len = (int)((M + S - 1) / S);
    S = (M + seg_len - 1) / seg_len;
    cplx* scratch = reinterpret_cast<cplx*>(ws);
    cplx* seg = reinterpret_cast<cplx*>(ws + product_scratch_bytes(D));
    ProductParams pp{};
    pp.mats = static_cast<const cplx*>(mats); pp.B = B; pp.M = M; pp.D = D;
    pp.S = (int)S; pp.seg_len = seg_len; pp.out = (S > 1) ? seg : static_cast<cplx*>(out); pp.ws = scratch;
    int rc = launch_product(pp, st);
    if (rc) return rc;
    if (S > 1) {
        ProductParams p2{};
        p2.mats = seg; p2.B = B; p2.M = (int)S; p2.D = D; p2.S = 1; p2.seg_len = (int)S;
        p2.out = static_cast<cplx*>(out); p2.ws = scratch;
        rc = launch_product(p2, st);
    }
    return rc;
}

int c3b_seq_product(const void* gates, int Gn, const int32_t* seq_idx, const int32_t* seq_len, int S, int Lmax,
                    int D, void* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (S <= 0 || D <= 0 || Gn <= 0 || Lmax < 0) return fail(C3B_EINVAL, "C3:ERROR: bad size (S=%d D=%d Gn=%d Lmax=%d)", S, D, Gn, Lmax);
    if (gates == nullptr || out == nullptr || seq_len == nullptr || (Lmax > 0 && seq_idx == nullptr))
        return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (D > 64 && (workspace == nullptr || workspace_bytes < product_scratch_bytes(D)))
        return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small");
    {
        const int rc = launch_seq_blk(static_cast<const cplx*>(gates), Gn, seq_idx, seq_len, S, Lmax, D, static_cast<cplx*>(out),
                                      static_cast<cudaStream_t>(stream));
        if (rc >= 0) return rc;
    }
    ProductParams pp{};
    pp.mats = static_cast<const cplx*>(gates); pp.idx = seq_idx; pp.lens = seq_len;
    pp.B = S; pp.M = Lmax; pp.D = D; pp.S = 1; pp.seg_len = Lmax > 0 ? Lmax : 1;
    pp.out = static_cast<cplx*>(out);
    pp.ws = static_cast<cplx*>(workspace);
    return launch_product(pp, static_cast<cudaStream_t>(stream));
}

int c3b_kron(const void* A, const void* Bm, void* out, int batch, int ra, int ca, int rb, int cb, int a_batched,
             int b_batched, void* stream) {
    if (batch <= 0 || ra <= 0 || ca <= 0 || rb <= 0 || cb <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size");
    if (A == nullptr || Bm == nullptr || out == nullptr) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    const long long total = (long long)batch * ra * rb * ca * cb;
    long long blocks = (total + 255) / 256;
    if (blocks > 65535) blocks = 65535;
    kron_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const cplx*>(A), static_cast<const cplx*>(Bm), static_cast<cplx*>(out), batch, ra, ca, rb, cb,
        a_batched ? (long long)ra * ca : 0, b_batched ? (long long)rb * cb : 0);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

double c3b_measure_fp64_peak(int kind, int device, double seconds) {
    double tf = 0.0;
    int rc = measure_fp64_peak(kind, device, seconds, &tf);
    if (rc != 0) return (double)fail(C3B_ECUDA, "C3:ERROR: fp64 peak measurement failed (cuda error %d)", rc);
    return tf;
}

// ---- gradient -------------------------------------------------------------
// variant 1 (default, d <= 16): Frechet derivative of the Taylor scheme on (X, dX) pairs, fused contraction
// variant 0: augmented 2d x 2d exponential through the forward kernels (any d <= 32)
static int grad_variant_for(int d) { return (g_grad_variant == 1 && d <= 16) ? 1 : 0; }

// D: matrix dimension of the propagators (d, or d^2 for Lindblad); dh: Hilbert dimension passed to the workspace query
static size_t grad_chunk_bytes(int Bc, int K, int N, int dh, int lindblad, size_t* off /*[9]*/) {
    const int D = lindblad ? dh * dh : dh;
    const size_t dd = (size_t)D * D * sizeof(cplx), dd2 = 4 * dd;
    const bool aug = grad_variant_for(D) == 0;
    size_t o = 0;
    off[0] = o; o += align_up(c3b_pwc_workspace_bytes(Bc, K, N, dh, lindblad, 0));       // forward workspace
    off[1] = o; if (aug) o += align_up(c3b_pwc_workspace_bytes(Bc, 0, N, 2 * D, 0, 0));   // augmented H-list workspace
    off[2] = o; o += align_up((size_t)Bc * dd);                                    // U (forward)
    off[3] = o; o += align_up((size_t)Bc * N * dd);                                // dUs
    off[4] = o; o += align_up((size_t)Bc * N * dd);                                // Psi, then M (in place)
    off[5] = o; o += align_up((size_t)Bc * sizeof(double));                        // alpha
    off[6] = o; if (aug) o += align_up((size_t)Bc * N * dd2);                      // Haug
    off[7] = o; if (aug) o += align_up((size_t)Bc * N * dd2);                      // exp(hscale Haug)
    off[8] = o; if (aug) o += align_up((size_t)Bc * dd2);                          // product of the augmented slices (unused)
    return o;
}

size_t c3b_pwc_grad_workspace_bytes(int B, int K, int N, int d, int chunk) {
    if (B <= 0 || N <= 0 || d <= 0 || K <= 0) return 0;
    const int Bc = (chunk > 0 && chunk < B) ? chunk : B;
    size_t off[9];
    return grad_chunk_bytes(Bc, K, N, d, 0, off);
}

size_t c3b_pwc_lindblad_grad_workspace_bytes(int B, int K, int N, int d, int chunk) {
    if (B <= 0 || N <= 0 || d <= 0 || K <= 0) return 0;
    const int Bc = (chunk > 0 && chunk < B) ? chunk : B;
    size_t off[9];
    return grad_chunk_bytes(Bc, K, N, d, 1, off);
}

static int pwc_grad_impl(int lindblad, const void* h0, const void* hks, const void* col_ops, int C, const double* signals,
                         double dt, int B, int K, int N, int dh, const void* Ubar, double* grad_out, void* U_out,
                         int chunk, void* workspace, size_t workspace_bytes, void* stream) {
    if (B <= 0 || N <= 0 || dh <= 0 || K <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d N=%d d=%d)", B, K, N, dh);
    if (!h0 || !hks || !signals || !Ubar || !grad_out || !workspace) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    const int d = lindblad ? dh * dh : dh;                      // dimension of the propagators
    if (d > 32) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gradient path supports matrix dimension <= 32 (got %d)", d);
    const int variant = grad_variant_for(d);
    if (lindblad && variant != 1)
        return fail(C3B_EUNSUPPORTED, "C3:ERROR: the Lindblad gradient needs d^2 <= 16 and grad_variant 1 (got d=%d)", dh);
    const int Bc = (chunk > 0 && chunk < B) ? chunk : B;
    size_t off[9];
    const size_t need = grad_chunk_bytes(Bc, K, N, dh, lindblad, off);
    if (workspace_bytes < need) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    const size_t dd = (size_t)d * d;
    cplx* Utmp = reinterpret_cast<cplx*>(ws + off[2]);
    cplx* dUs = reinterpret_cast<cplx*>(ws + off[3]);
    cplx* Psi = reinterpret_cast<cplx*>(ws + off[4]);
    double* alpha = reinterpret_cast<double*>(ws + off[5]);
    cplx* Haug = reinterpret_cast<cplx*>(ws + off[6]);
    cplx* Eaug = reinterpret_cast<cplx*>(ws + off[7]);
    cplx* Uaug = reinterpret_cast<cplx*>(ws + off[8]);
    // sweep kernels: one warp per batch row with 3-4 matrices in shared memory; d = 32 needs 64 KB per warp
    int wpb = (int)((size_t)192 * 1024 / ((size_t)4 * dd * sizeof(cplx)));
    if (wpb > 4) wpb = 4;
    if (wpb < 1) wpb = 1;
    const int sweep_smem = (int)((size_t)wpb * 4 * dd * sizeof(cplx));
    CUDA_TRY(cudaFuncSetAttribute(grad_suffix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem));
    CUDA_TRY(cudaFuncSetAttribute(grad_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem));
    CUDA_TRY(cudaFuncSetAttribute(grad_suffix2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem));
    CUDA_TRY(cudaFuncSetAttribute(grad_prefix2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem));
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int nb = (B - b0 < Bc) ? (B - b0) : Bc;
        const double* sig = signals + (size_t)b0 * K * N;
        cplx* Udst = U_out ? static_cast<cplx*>(U_out) + (size_t)b0 * dd : Utmp;
        int rc = lindblad
            ? c3b_pwc_lindblad(h0, hks, col_ops, C, sig, dt, nb, K, N, dh, 0, Udst, dUs, ws + off[0], off[1] - off[0], stream)
            : c3b_pwc_closed(h0, hks, sig, dt, nb, K, N, dh, 0, Udst, dUs, ws + off[0], off[1] - off[0], stream);
        if (rc) return rc;
        const int blocks = (nb + wpb - 1) / wpb;
        const cplx* ub = static_cast<const cplx*>(Ubar) + (size_t)b0 * dd;
        if (variant == 1) {
            grad_suffix2_kernel<<<blocks, wpb * 32, (size_t)wpb * 3 * dd * sizeof(cplx), st>>>(dUs, ub, Psi, alpha, nb, N, d);
            CUDA_TRY(cudaGetLastError());
            g_launches.fetch_add(1, std::memory_order_relaxed);
            grad_prefix2_kernel<<<blocks, wpb * 32, (size_t)wpb * 4 * dd * sizeof(cplx), st>>>(dUs, Psi, nb, N, d);
            CUDA_TRY(cudaGetLastError());
            g_launches.fetch_add(1, std::memory_order_relaxed);
            // generators, row sums and trace shifts as the forward call left them in its workspace
            const Plan pl = make_plan(nb, K, N, d, 0, false);
            char* fws = ws + off[0];
            const cplx* G = reinterpret_cast<const cplx*>(fws + pl.off_G);
            const double* RS = reinterpret_cast<const double*>(fws + pl.off_RS);
            const bool shifted = (pl.path == 1 && rows_kernel_takes_shift(d)) || (pl.path != 1 && g_cta_variant == 1);
            const cplx* TR = shifted ? reinterpret_cast<const cplx*>(fws + pl.off_TR) : nullptr;
            const size_t per_warp = (size_t)kFrechetBufs * dd * sizeof(cplx);
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
            grad_frechet_kernel<<<(int)grid, fw * 32, smem, st>>>(G, RS, TR, sig, Psi, alpha, grad_out + (size_t)b0 * K * N,
                                                                nb, K, N, d);
            CUDA_TRY(cudaGetLastError());
            g_launches.fetch_add(1, std::memory_order_relaxed);
            continue;
        }
        grad_suffix_kernel<<<blocks, wpb * 32, (size_t)wpb * 3 * dd * sizeof(cplx), st>>>(dUs, ub, Psi, alpha, nb, N, d);
        CUDA_TRY(cudaGetLastError());
        g_launches.fetch_add(1, std::memory_order_relaxed);
        grad_prefix_kernel<<<blocks, wpb * 32, (size_t)wpb * 4 * dd * sizeof(cplx), st>>>(
            dUs, Psi, static_cast<const cplx*>(h0), static_cast<const cplx*>(hks), sig, Haug, dt, nb, K, N, d);
        CUDA_TRY(cudaGetLastError());
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = c3b_pwc_closed_hlist(Haug, dt, nb, N, 2 * d, Uaug, Eaug, ws + off[1], off[2] - off[1], stream);
        if (rc) return rc;
        const long long warps = (long long)nb * N;
        grad_contract_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(
            Eaug, static_cast<const cplx*>(hks), alpha, grad_out + (size_t)b0 * K * N, dt, nb, K, N, d);
        CUDA_TRY(cudaGetLastError());
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return C3B_OK;
}

int c3b_pwc_closed_grad(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                        const void* Ubar, double* grad_out, void* U_out, int chunk, void* workspace,
                        size_t workspace_bytes, void* stream) {
    return pwc_grad_impl(0, h0, hks, nullptr, 0, signals, dt, B, K, N, d, Ubar, grad_out, U_out, chunk, workspace,
                         workspace_bytes, stream);
}

int c3b_pwc_lindblad_grad(const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                          int B, int K, int N, int d, const void* Ubar, double* grad_out, void* U_out, int chunk,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (C < 0 || (C > 0 && col_ops == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: C=%d but col_ops is NULL", C);
    return pwc_grad_impl(1, h0, hks, col_ops, C, signals, dt, B, K, N, d, Ubar, grad_out, U_out, chunk, workspace,
                         workspace_bytes, stream);
}

// ---- goal functions on the propagators ------------------------------------
int c3b_gate_infid(const void* U, int B, int D, const void* ideal, const int32_t* sel, int C, int mode,
                   double* infid_out, void* overlap_out, void* stream) {
    if (B <= 0 || D <= 0 || C <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d D=%d C=%d)", B, D, C);
    if (C > D) return fail(C3B_EINVAL, "C3:ERROR: computational subspace (%d) larger than the matrix (%d)", C, D);
    if (mode < 0 || mode > 3) return fail(C3B_EINVAL, "C3:ERROR: unknown fidelity mode %d", mode);
    if (!U || !ideal || !sel || (!infid_out && !overlap_out)) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    const int wpb = 4;
    gate_overlap_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const cplx*>(U), static_cast<const cplx*>(ideal), sel, B, D, C, mode, infid_out,
        static_cast<cplx*>(overlap_out));
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

int c3b_gate_infid_grad(const void* overlap, const void* ideal, const int32_t* sel, const double* gbar, int B, int D,
                        int C, int mode, void* Ubar_out, void* stream) {
    if (B <= 0 || D <= 0 || C <= 0 || C > D) return fail(C3B_EINVAL, "C3:ERROR: bad size (B=%d D=%d C=%d)", B, D, C);
    if (mode != 0 && mode != 1) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gradient only for unitary_infid / average_infid (mode 0/1), got %d", mode);
    if (!overlap || !ideal || !sel || !Ubar_out) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(Ubar_out, 0, (size_t)B * D * D * sizeof(cplx), st));
    const long long total = (long long)B * C * C;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    gate_overlap_grad_kernel<<<(int)blocks, 256, 0, st>>>(static_cast<const cplx*>(overlap), static_cast<const cplx*>(ideal),
                                                         sel, gbar, B, D, C, mode, static_cast<cplx*>(Ubar_out));
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

int c3b_seq_populations(const void* gates, int Gn, const int32_t* seq_idx, const int32_t* seq_len, int S, int Lmax,
                        int D, const void* psi0, int lindblad_d, double* pops_out, void* psi_out, void* stream) {
    if (S <= 0 || D <= 0 || Gn <= 0 || Lmax < 0) return fail(C3B_EINVAL, "C3:ERROR: bad size (S=%d D=%d Gn=%d Lmax=%d)", S, D, Gn, Lmax);
    if (!gates || !seq_len || (Lmax > 0 && !seq_idx) || (!pops_out && !psi_out)) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (lindblad_d < 0 || (lindblad_d > 0 && lindblad_d * lindblad_d != D))
        return fail(C3B_EINVAL, "C3:ERROR: Lindblad populations need D = d^2 (D=%d, d=%d)", D, lindblad_d);
    constexpr int W = 4;
    const size_t smem = (size_t)W * 2 * D * sizeof(cplx);
    if (smem > 200 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: state dimension %d too large", D);
    auto kern = seq_state_kernel<W>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(S + W - 1) / W, W * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const cplx*>(gates), seq_idx, seq_len, static_cast<const cplx*>(psi0), S, Lmax > 0 ? Lmax : 1, D,
        lindblad_d, pops_out, static_cast<cplx*>(psi_out));
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

// ---- signal generation chain ----------------------------------------------------
int c3b_signal_slice_num(double t_start, double t_end, double resolution) {
    const double span = t_start - t_end;
    return (int)((span < 0 ? -span : span) * resolution);   // Device.calc_slice_num, c3/generator/devices.py:73-85
}

int c3b_generate_signals(const double* env_params, const int32_t* env_shape, const int32_t* env_flags,
                         const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                         int B, int K, int E, int N, double* signals_out, void* stream) {
    if (B <= 0 || K <= 0 || E <= 0 || N <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d E=%d N=%d)", B, K, E, N);
    if (!env_params || !env_shape || !env_flags || !lo_freq || !chain || !signals_out) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    SignalParams sp{};
    sp.env = env_params; sp.shape = env_shape; sp.flags = env_flags; sp.lo_freq = lo_freq; sp.chain = chain;
    sp.chain_batched = chain_batched; sp.t_start = t_start; sp.t_end = t_end;
    sp.B = B; sp.K = K; sp.E = E; sp.N = N; sp.out = signals_out;
    // the AWG grid and the response taps live in shared memory: bounded by the simulation grid / 4096 taps
    sp.max_awg = N + 1;
    sp.max_taps = 4096;
    const size_t smem = ((size_t)2 * sp.max_awg + sp.max_taps) * sizeof(double);
    if (smem > 200 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gate too long for the on-chip signal chain (N=%d)", N);
    CUDA_TRY(cudaFuncSetAttribute(signal_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    signal_chain_kernel<<<B * K, 128, smem, static_cast<cudaStream_t>(stream)>>>(sp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

int c3b_generate_signals_grad(const double* env_params, const int32_t* env_shape, const int32_t* env_flags,
                              const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                              int B, int K, int E, int N, int n_awg_max, const double* gsignals, double* grad_env,
                              double* grad_lo, double* grad_v2hz, void* stream) {
    if (B <= 0 || K <= 0 || E <= 0 || N <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d E=%d N=%d)", B, K, E, N);
    if (!env_params || !env_shape || !env_flags || !lo_freq || !chain || !gsignals || !grad_env || !grad_lo)
        return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    SignalGradParams gp{};
    SignalParams& sp = gp.f;
    sp.env = env_params; sp.shape = env_shape; sp.flags = env_flags; sp.lo_freq = lo_freq; sp.chain = chain;
    sp.chain_batched = chain_batched; sp.t_start = t_start; sp.t_end = t_end;
    sp.B = B; sp.K = K; sp.E = E; sp.N = N; sp.out = nullptr;
    sp.max_awg = (n_awg_max > 0 && n_awg_max <= N) ? n_awg_max : N + 1;
    sp.max_taps = 1024;
    gp.gsig = gsignals; gp.genv = grad_env; gp.glo = grad_lo; gp.gv2hz = grad_v2hz;
    const size_t smem = ((size_t)4 * sp.max_awg + sp.max_taps + (size_t)2 * N) * sizeof(double);
    if (smem > 200 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gate too long for the on-chip signal-chain gradient (N=%d)", N);
    CUDA_TRY(cudaFuncSetAttribute(signal_chain_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    signal_chain_grad_kernel<<<B * K, 128, smem, static_cast<cudaStream_t>(stream)>>>(gp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

// ---- batched dressing of model samples --------------------------------------------------
int c3b_dress_models(const void* drift, const void* ops, int ops_batched, int B, int M, int d, int ordered,
                     double* eigenframe, void* transform, void* dressed_drift, void* dressed_ops, int32_t* info,
                     void* stream) {
    if (B <= 0 || d <= 0 || M < 0) return fail(C3B_EINVAL, "C3:ERROR: bad size (B=%d M=%d d=%d)", B, M, d);
    if (d > 32) return fail(C3B_EUNSUPPORTED, "C3:ERROR: on-device dressing supports d <= 32 (got %d)", d);
    if (!drift || !eigenframe || !transform) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (M > 0 && dressed_ops && !ops) return fail(C3B_EINVAL, "C3:ERROR: dressed_ops requested but ops is NULL");
    DressParams dp{};
    dp.drift = static_cast<const cplx*>(drift); dp.ops = static_cast<const cplx*>(ops); dp.ops_batched = ops_batched;
    dp.B = B; dp.M = M; dp.d = d; dp.ordered = ordered;
    dp.eigenframe = eigenframe; dp.transform = static_cast<cplx*>(transform);
    dp.dressed_drift = static_cast<cplx*>(dressed_drift); dp.dressed_ops = static_cast<cplx*>(dressed_ops); dp.info = info;
    const size_t per_warp = (size_t)4 * d * d * sizeof(cplx) + (size_t)4 * d * sizeof(double);
    int wpb = (int)((size_t)96 * 1024 / per_warp);
    if (wpb > 4) wpb = 4;
    if (wpb < 1) wpb = 1;
    const size_t smem = wpb * per_warp;
    CUDA_TRY(cudaFuncSetAttribute(dress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dress_kernel<<<(B + wpb - 1) / wpb, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(dp);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return C3B_OK;
}

long long c3b_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

double c3b_last_kernel_ms(void) {
    if (!g_ev_valid) return (double)fail(C3B_EINVAL, "C3:ERROR: no profiled launch (set tuning 'profile' to 1 first)");
    if (cudaEventSynchronize(g_ev1) != cudaSuccess) return (double)fail(C3B_ECUDA, "C3:ERROR: event synchronize failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return (double)fail(C3B_ECUDA, "C3:ERROR: event elapsed failed");
    return (double)ms;
}

double c3b_microbench(int kind, int a, int b) {
    double tf = 0.0;
    int rc = -1;
    if (kind == 0) rc = run_mmrow_bench<9>(a, b, &tf);
    else if (kind == 1) rc = run_mmrow_bench<4>(a, b, &tf);
    else if (kind == 2) rc = run_mmrow_bench<3>(a, b, &tf);
    else return (double)fail(C3B_EINVAL, "C3:ERROR: unknown microbench kind %d", kind);
    if (rc != 0) return (double)fail(C3B_ECUDA, "C3:ERROR: microbench failed (cuda error %d)", rc);
    return tf;
}

}  // extern "C"
