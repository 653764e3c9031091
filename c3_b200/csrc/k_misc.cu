// Entry points that sit beside the propagator path: goal functions (f-3), signal chain (f-2), dressing (f-4) and the fp64
// peak measurement the roofline is quoted against.  Contracts in include/c3b200.h.
#include "c3b_host.cuh"
#include "fidelity.cuh"
#include "signal_chain.cuh"
#include "dressing.cuh"
#include "frame.cuh"
#include "peak.cuh"

using namespace c3b;

extern "C" {

double c3b_measure_fp64_peak(int kind, int device, double seconds) {
    double tf = 0.0;
    int rc = measure_fp64_peak(kind, device, seconds, &tf);
    if (rc != 0) return (double)fail(C3B_ECUDA, "C3:ERROR: fp64 peak measurement failed (cuda error %d)", rc);
    return tf;
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
    count_launch();
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
    count_launch();
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
    count_launch();
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
    return c3b_generate_signals_noisy(env_params, env_shape, env_flags, lo_freq, chain, chain_batched, t_start, t_end, B, K, E, N,
                                      nullptr, 0, 0ull, signals_out, nullptr, stream);
}

int c3b_generate_signals_noisy(const double* env_params, const int32_t* env_shape, const int32_t* env_flags,
                               const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                               int B, int K, int E, int N, const double* noise, int noise_batched, unsigned long long seed,
                               double* signals_out, double* noise_out, void* stream) {
    return c3b_generate_signals_table(env_params, env_shape, env_flags, nullptr, 0, lo_freq, chain, chain_batched, t_start, t_end, B, K, E, N,
                                      noise, noise_batched, seed, signals_out, noise_out, stream);
}

int c3b_generate_signals_table(const double* env_params, const int32_t* env_shape, const int32_t* env_flags, const double* env_table,
                               int T, const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                               int B, int K, int E, int N, const double* noise, int noise_batched, unsigned long long seed,
                               double* signals_out, double* noise_out, void* stream) {
    if (T < 0 || (T > 0 && env_table == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: T=%d but env_table is NULL", T);
    if (B <= 0 || K <= 0 || E <= 0 || N <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d E=%d N=%d)", B, K, E, N);
    if (!env_params || !env_shape || !env_flags || !lo_freq || !chain || !signals_out) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (noise_out != nullptr && noise == nullptr) return fail(C3B_EINVAL, "C3:ERROR: noise traces requested without noise parameters");
    SignalParams sp{};
    sp.env = env_params; sp.shape = env_shape; sp.flags = env_flags; sp.lo_freq = lo_freq; sp.chain = chain;
    sp.chain_batched = chain_batched; sp.t_start = t_start; sp.t_end = t_end;
    sp.B = B; sp.K = K; sp.E = E; sp.N = N; sp.out = signals_out;
    sp.noise = noise; sp.noise_batched = noise_batched; sp.seed = seed; sp.noise_out = noise_out;
    sp.table = T > 0 ? env_table : nullptr; sp.T = T;
    // the AWG grid and the response taps live in shared memory: bounded by the simulation grid / 4096 taps
    sp.max_awg = N + 1;
    sp.max_taps = 4096;
    const size_t smem = ((size_t)2 * sp.max_awg + sp.max_taps) * sizeof(double) + (noise ? (size_t)N * sizeof(int) : 0);   // + Pink_Noise sums
    if (smem > 200 * 1024) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gate too long for the on-chip signal chain (N=%d)", N);
    CUDA_TRY(cudaFuncSetAttribute(signal_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    signal_chain_kernel<<<B * K, 128, smem, static_cast<cudaStream_t>(stream)>>>(sp);
    CUDA_TRY(cudaGetLastError());
    count_launch();
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
    count_launch();
    return C3B_OK;
}

// ---- Crosstalk device (c3/generator/devices.py:225-293): the drive lines listed in chan mixed by a C x C matrix, in place:
//      out[b, chan[i], n] = sum_j M[i, j] in[b, chan[j], n]   (Generator.generate_signals applies it to the finished lines,
//      c3/generator/generator.py:229-234).  One thread per (batch row, sample); C <= 16 values in registers.
namespace c3b {
namespace {
constexpr int kMaxCrossed = 16;
__global__ void crosstalk_kernel(double* __restrict__ sig, const int* __restrict__ chan, const double* __restrict__ M, const int B,
                                 const int K, const int N, const int C) {
    const long long total = (long long)B * N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / N), n = (int)(t - (long long)b * N);
        double* row = sig + (size_t)b * K * N + n;
        double in[kMaxCrossed];
#pragma unroll
        for (int j = 0; j < kMaxCrossed; ++j) in[j] = j < C ? row[(size_t)chan[j] * N] : 0.0;
#pragma unroll
        for (int i = 0; i < kMaxCrossed; ++i) {
            if (i < C) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < kMaxCrossed; ++j)
                    if (j < C) acc += M[i * C + j] * in[j];          // same summation order as the reference's matmul row
                row[(size_t)chan[i] * N] = acc;
            }
        }
    }
}
}  // namespace
}  // namespace c3b

extern "C" int c3b_crosstalk(double* signals, int B, int K, int N, const int32_t* chan, int C, const double* matrix, void* stream) {
    using namespace c3b;
    if (B <= 0 || K <= 0 || N <= 0 || C <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d N=%d C=%d)", B, K, N, C);
    if (!signals || !chan || !matrix) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (C > kMaxCrossed || C > K) return fail(C3B_EUNSUPPORTED, "C3:ERROR: crosstalk between %d lines (at most %d, and at most K = %d)", C, kMaxCrossed, K);
    const long long total = (long long)B * N;
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    crosstalk_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(signals, chan, matrix, B, K, N, C);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return C3B_OK;
}

// ---- frame rotation / dephasing channel on the device ----------------------------------------
int c3b_frame_dephase(void* U, int B, int D, int d, const int32_t* occ, int L, const double* phases, const double* probs,
                      int lindblad, void* stream) {
    if (B <= 0 || D <= 0 || d <= 0 || L < 0) return fail(C3B_EINVAL, "C3:ERROR: bad size (B=%d D=%d d=%d L=%d)", B, D, d, L);
    if (!U || (L > 0 && !occ)) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (lindblad ? (D != d * d) : (D != d)) return fail(C3B_EINVAL, "C3:ERROR: D=%d does not match d=%d (lindblad=%d)", D, d, lindblad);
    if (probs != nullptr && !lindblad) return fail(C3B_EINVAL, "C3:ERROR: Dephasing can only be added when lindblad is on.");
    if (L == 0 || (phases == nullptr && probs == nullptr)) return C3B_OK;
    const long long warps = (long long)B * D;
    long long blocks = (warps * 32 + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    frame_dephase_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<cplx*>(U), occ, phases, probs, B, D, d, L,
                                                                                  lindblad);
    CUDA_TRY(cudaGetLastError());
    count_launch();
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
    count_launch();
    return C3B_OK;
}

}  // extern "C"
