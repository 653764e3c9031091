// C ABI of libc3b200.so -- see include/c3b200.h for the contract of every entry point.
// This file: error state, per-thread tuning, launch planning, prepared models and the propagator / product / gradient
// entry points.  The kernels live in k_*.cu (one translation unit per family, c3b_host.cuh lists their launchers).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "c3b_host.cuh"

using namespace c3b;

namespace c3b {

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};   // kernels launched by this library (bench.py's gpu_launches)

long long env_ll(const char* name, long long dflt) {
    const char* v = getenv(name);
    return v ? atoll(v) : dflt;
}

Tuning initial_tuning() {
    Tuning t;
    t.d9_variant = env_ll("C3B_D9_VARIANT", t.d9_variant);
    t.cta_variant = env_ll("C3B_CTA_VARIANT", t.cta_variant);
    t.cta_threads = env_ll("C3B_CTA_THREADS", t.cta_threads);
    t.gemm_big = env_ll("C3B_GEMM_BIG", t.gemm_big);
    t.grad_variant = env_ll("C3B_GRAD_VARIANT", t.grad_variant);
    t.min_chunk = env_ll("C3B_MIN_CHUNK", t.min_chunk);
    return t;
}

// profile events of the calling thread, created on the device of the stream they are recorded on
struct ProfileEvents {
    int dev = -1;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool valid = false;
};
thread_local ProfileEvents g_prof;

int profile_events_for_current_device() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (g_prof.dev != dev) {
        if (g_prof.e0) { cudaEventDestroy(g_prof.e0); cudaEventDestroy(g_prof.e1); g_prof.e0 = g_prof.e1 = nullptr; }
        CUDA_TRY(cudaEventCreate(&g_prof.e0));
        CUDA_TRY(cudaEventCreate(&g_prof.e1));
        g_prof.dev = dev;
        g_prof.valid = false;
    }
    return C3B_OK;
}
}  // namespace

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

Tuning& tuning() {
    thread_local Tuning t = initial_tuning();
    return t;
}

int num_sms() {
    thread_local int cached_dev = -1, cached = 0;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }   // B200; sizing only, no device visible
    if (dev == cached_dev) return cached;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); return 148; }
    cached_dev = dev;
    cached = n;
    return n;
}

int cta_grid(int D, long long units) {
    int per_sm = (tuning().gemm_big >= 2) ? 1 : 2;              // global-workspace kernels: 2 x 256 threads, or 1 x 384 / 512 / 704
    if (D <= 32) {
        if (tuning().cta_variant == 1) {
            per_sm = gemm_ctas_per_sm(D);
            if (round8(D) == 32 && tuning().cta_threads == 512) per_sm = 1;      // 512-thread CTAs: one per SM (registers)
        } else {
            per_sm = (int)((size_t)220 * 1024 / (cta_smem_bytes(D) + 1024));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 4) per_sm = 4;
        }
    }
    long long g = (long long)num_sms() * per_sm;
    if (g > units) g = units;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace c3b

namespace {

// 1: lane-group kernels (d <= 12, shared model), 2: CTA kernel with matrices in shared memory, 3: CTA kernel with a global workspace
int pwc_path(int D, int batched_model) {
    if (!tuning().force_cta && !batched_model && blk_template_dim(D) != 0) return 1;
    return D <= 32 ? 2 : 3;
}

// ---- prepared model: everything of a launch that depends on the model only ----------------------------------------------
struct ModelLayout {
    size_t off_G, off_RS, off_TR, total;
};

ModelLayout model_layout(int K, int D, int Bm) {
    ModelLayout ml{};
    size_t off = 0;
    ml.off_G = off;  off += align_up((size_t)Bm * (K + 1) * D * D * sizeof(cplx));
    ml.off_RS = off; off += align_up((size_t)Bm * (K + 1) * D * sizeof(double));
    ml.off_TR = off; off += align_up((size_t)Bm * (K + 1) * sizeof(cplx));
    ml.total = off;
    return ml;
}

// ---- launch plan ------------------------------------------------------------------------------------------------------
struct Plan {
    int path;
    int S;         // segments per batch element
    int seg_len;
    int grid;      // CTA kernel grid (persistent)
    size_t off_seg, off_cta, off_prod, off_counter, total;
};

// seg_len_override > 0: segments of exactly that many slices (the fused gradient path wants the chunk products)
Plan make_plan(int B, int N, int D, int batched_model, int seg_len_override = 0) {
    const Tuning& tn = tuning();
    Plan pl{};
    pl.path = pwc_path(D, batched_model);
    long long S, smax;
    if (seg_len_override > 0) {
        S = (N + seg_len_override - 1) / seg_len_override;
        smax = S;
    } else if (pl.path == 1) {
        const long long target = tn.target_units > 0 ? tn.target_units : (B <= 768 ? 4096 : 12288);
        S = (target + B - 1) / B;
        smax = N / (tn.min_chunk * blk_groups_per_warp(D));
    } else {
        S = (8LL * num_sms() + B - 1) / B;
        smax = N / 8;
    }
    if (smax < 1) smax = 1;
    if (S > smax) S = smax;
    if (S < 1) S = 1;
    pl.S = (int)S;
    pl.seg_len = (N + pl.S - 1) / pl.S;
    if (pl.path == 1 && seg_len_override <= 0 && pl.S > 1) {       // whole iterations: every lane group of a warp gets the same count
        const int G = blk_groups_per_warp(D);
        pl.seg_len = (pl.seg_len + G - 1) / G * G;
    }
    pl.S = (N + pl.seg_len - 1) / pl.seg_len;  // drop empty trailing segments
    if (pl.S < 1) pl.S = 1;
    pl.grid = cta_grid(D, (long long)B * pl.S);
    size_t off = 0;
    pl.off_seg = off;
    if (pl.S > 1) off += align_up((size_t)B * pl.S * D * D * sizeof(cplx));
    pl.off_cta = off;
    if (pl.path == 3) off += align_up((size_t)pl.grid * kGemmSlots * round8(D) * round8(D) * sizeof(cplx));
    pl.off_prod = off;
    if (pl.S > 1 && D > 64) off += align_up((size_t)cta_grid(D, 1LL << 40) * 2 * D * D * sizeof(cplx));
    pl.off_counter = off;
    off += 256;
    pl.total = off;
    return pl;
}

int check_common(int B, int K, int N, int d, const void* U_out, const void* ws) {
    if (B <= 0 || N <= 0 || d <= 0 || K < 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d N=%d d=%d)", B, K, N, d);
    if (U_out == nullptr) return fail(C3B_EINVAL, "C3:ERROR: U_out is NULL");
    if (ws == nullptr) return fail(C3B_EINVAL, "C3:ERROR: workspace is NULL");
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return fail(C3B_EINVAL, "C3:ERROR: workspace must be 16-byte aligned");
    return C3B_OK;
}

// generators + trace shift + row sums into a model blob (3 small kernels; once per model, not per call, for prepared models)
int build_model(const void* h0, const void* hks, const void* col_ops, int C, double dt, int K, int d, int lindblad, int Bm,
                char* blob, cudaStream_t st) {
    const int D = lindblad ? d * d : d;
    const ModelLayout ml = model_layout(K, D, Bm);
    cplx* G = reinterpret_cast<cplx*>(blob + ml.off_G);
    double* RS = reinterpret_cast<double*>(blob + ml.off_RS);
    cplx* TR = reinterpret_cast<cplx*>(blob + ml.off_TR);
    int rc = lindblad ? launch_setup_lindblad(static_cast<const cplx*>(h0), static_cast<const cplx*>(hks),
                                              static_cast<const cplx*>(col_ops), G, Bm, K, C, d, dt, st)
                      : launch_setup_closed(static_cast<const cplx*>(h0), static_cast<const cplx*>(hks), G, Bm, K, d, dt, st);
    if (rc) return rc;
    if ((rc = launch_trace_shift(G, TR, (long long)Bm * (K + 1), D, st)) != 0) return rc;
    return launch_rowsum(G, RS, (long long)Bm * (K + 1) * D, D, st);
}

// common tail of the pwc entry points once the model blob (or the H list) is in place
int run_pwc(const Plan& pl, const char* blob, const double* signals, const cplx* hlist, double dt, int B, int K, int N, int D,
            int batched_model, cplx* U_out, cplx* dUs_out, unsigned int* gate, char* ws, cudaStream_t st, bool fold = true) {
    const Tuning& tn = tuning();
    const cplx* G = nullptr;
    const double* RS = nullptr;
    const cplx* TR = nullptr;
    if (blob != nullptr) {
        const ModelLayout ml = model_layout(K, D, batched_model ? B : 1);
        G = reinterpret_cast<const cplx*>(blob + ml.off_G);
        RS = reinterpret_cast<const double*>(blob + ml.off_RS);
        TR = reinterpret_cast<const cplx*>(blob + ml.off_TR);
    }
    const bool gated_ok = pl.path == 1 && D == 9 && d9_gated_supported((int)tn.d9_variant);
    if (gate != nullptr && !gated_ok)
        return fail(C3B_EUNSUPPORTED, "C3:ERROR: gated launch is built for the d = 9 kernel only");
    cplx* seg = pl.S > 1 ? reinterpret_cast<cplx*>(ws + pl.off_seg) : nullptr;
    if (gate != nullptr && seg != nullptr)     // a row that never arrives must surface as NaN after the fold, too
        CUDA_TRY(cudaMemsetAsync(seg, 0xff, (size_t)B * pl.S * D * D * sizeof(cplx), st));
    if (tn.profile) {
        int rc = profile_events_for_current_device();
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(g_prof.e0, st));
    }
    if (pl.path == 1) {
        RowsParams rp{};
        rp.G = G; rp.RS = RS; rp.TR = TR; rp.signals = signals; rp.hlist = hlist;
        rp.hscale_re = 0.0; rp.hscale_im = -dt;
        rp.model_stride = 0;
        rp.B = B; rp.K = K; rp.N = N; rp.d = D; rp.S = pl.S; rp.seg_len = pl.seg_len;
        rp.U_out = U_out; rp.seg_out = seg; rp.dUs_out = dUs_out;
        rp.gate = gate;
        rp.skew = (int)tn.d9_skew;
        unsigned int* counter = reinterpret_cast<unsigned int*>(ws + pl.off_counter);
        int rc = (blk_template_dim(D) == 9) ? launch_d9(rp, counter, (int)tn.d9_variant, st) : launch_small(rp, counter, st);
        if (rc) return rc;
    } else {
        CtaParams cp{};
        cp.G = G; cp.signals = signals; cp.hlist = hlist;
        cp.hscale_re = 0.0; cp.hscale_im = -dt;
        cp.model_stride = batched_model ? (long long)(K + 1) * D * D : 0;
        cp.B = B; cp.K = K; cp.N = N; cp.D = D; cp.S = pl.S; cp.seg_len = pl.seg_len;
        cp.U_out = U_out; cp.seg_out = seg; cp.dUs_out = dUs_out;
        cp.ws = reinterpret_cast<cplx*>(ws + pl.off_cta);
        cp.use_smem = pl.path == 2;
        int rc = (tn.cta_variant == 1) ? launch_gemm(cp, TR, RS, pl.grid, st) : launch_cta(cp, TR, pl.grid, st);
        if (rc) return rc;
    }
    if (tn.profile) { CUDA_TRY(cudaEventRecord(g_prof.e1, st)); g_prof.valid = true; }
    if (pl.S > 1 && fold) {              // (fold == false: the caller folds the segment products itself)
        ProductParams pp{};
        pp.mats = seg; pp.idx = nullptr; pp.lens = nullptr;
        pp.B = B; pp.M = pl.S; pp.D = D; pp.S = 1; pp.seg_len = pl.S;
        pp.out = U_out; pp.ws = reinterpret_cast<cplx*>(ws + pl.off_prod);
        int rc = launch_product(pp, st);
        if (rc) return rc;
    }
    return C3B_OK;
}

int pwc_unprepared(int lindblad, const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                   int B, int K, int N, int d, int batched_model, void* U_out, void* dUs_out, unsigned int* gate,
                   void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(B, K, N, d, U_out, workspace);
    if (rc) return rc;
    if (h0 == nullptr) return fail(C3B_EINVAL, "C3:ERROR: h0 is NULL");
    if (K > 0 && (hks == nullptr || signals == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: K=%d but hks/signals is NULL", K);
    if (lindblad && (C < 0 || (C > 0 && col_ops == nullptr))) return fail(C3B_EINVAL, "C3:ERROR: C=%d but col_ops is NULL", C);
    if (lindblad && d > 181) return fail(C3B_EUNSUPPORTED, "C3:ERROR: Lindblad d=%d too large", d);
    const int D = lindblad ? d * d : d;
    const int Bm = batched_model ? B : 1;
    const ModelLayout ml = model_layout(K, D, Bm);
    const Plan pl = make_plan(B, N, D, batched_model);
    if (workspace_bytes < ml.total + pl.total)
        return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, ml.total + pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    if ((rc = build_model(h0, hks, col_ops, C, dt, K, d, lindblad, Bm, ws, st)) != 0) return rc;
    return run_pwc(pl, ws, signals, nullptr, dt, B, K, N, D, batched_model, static_cast<cplx*>(U_out), static_cast<cplx*>(dUs_out),
                   gate, ws + ml.total, st);
}

}  // namespace

extern "C" {

int c3b_version(void) { return 200; }

const char* c3b_last_error(void) { return g_err; }

int c3b_set_tuning(const char* key, long long value) {
    if (key == nullptr) return fail(C3B_EINVAL, "C3:ERROR: null tuning key");
    Tuning& t = tuning();
    if (!strcmp(key, "target_units")) { t.target_units = value < 0 ? 0 : value; return C3B_OK; }
    if (!strcmp(key, "min_chunk")) { t.min_chunk = value < 1 ? 1 : value; return C3B_OK; }
    if (!strcmp(key, "d9_variant")) { t.d9_variant = value; return C3B_OK; }
    if (!strcmp(key, "d9_skew")) { t.d9_skew = value < 0 ? 0 : value; return C3B_OK; }
    if (!strcmp(key, "force_cta")) { t.force_cta = value; return C3B_OK; }
    if (!strcmp(key, "cta_variant")) { t.cta_variant = value; return C3B_OK; }
    if (!strcmp(key, "cta_threads")) { t.cta_threads = value; return C3B_OK; }
    if (!strcmp(key, "gemm_big")) { t.gemm_big = value; return C3B_OK; }
    if (!strcmp(key, "norm_bound")) { t.norm_bound = value; return C3B_OK; }
    if (!strcmp(key, "seq_variant")) { t.seq_variant = value; return C3B_OK; }
    if (!strcmp(key, "grad_variant")) { t.grad_variant = value; return C3B_OK; }
    if (!strcmp(key, "grad_unitary")) { t.grad_unitary = value; return C3B_OK; }
    if (!strcmp(key, "grad_chunk")) { t.grad_chunk = value < 0 ? 0 : value; return C3B_OK; }
    if (!strcmp(key, "profile")) { t.profile = value; return C3B_OK; }
    return fail(C3B_EINVAL, "C3:ERROR: unknown tuning key '%s'", key);
}

int c3b_pwc_path(int K, int D, int batched_model) { (void)K; return pwc_path(D, batched_model); }

size_t c3b_pwc_workspace_bytes(int B, int K, int N, int d, int lindblad, int batched_model) {
    if (B <= 0 || N <= 0 || d <= 0 || K < 0) return 0;
    const int D = lindblad ? d * d : d;
    return model_layout(K, D, batched_model ? B : 1).total + make_plan(B, N, D, batched_model).total;
}

int c3b_pwc_closed(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                   int batched_model, void* U_out, void* dUs_out, void* workspace, size_t workspace_bytes,
                   void* stream) {
    return pwc_unprepared(0, h0, hks, nullptr, 0, signals, dt, B, K, N, d, batched_model, U_out, dUs_out, nullptr, workspace,
                          workspace_bytes, stream);
}

int c3b_pwc_lindblad(const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                     int B, int K, int N, int d, int batched_model, void* U_out, void* dUs_out, void* workspace,
                     size_t workspace_bytes, void* stream) {
    return pwc_unprepared(1, h0, hks, col_ops, C, signals, dt, B, K, N, d, batched_model, U_out, dUs_out, nullptr, workspace,
                          workspace_bytes, stream);
}

int c3b_pwc_closed_gated(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                         void* U_out, uint32_t* gate, void* workspace, size_t workspace_bytes, void* stream) {
    if (gate == nullptr) return fail(C3B_EINVAL, "C3:ERROR: gate is NULL");
    if (K <= 0) return fail(C3B_EINVAL, "C3:ERROR: gated launch needs control fields (K > 0)");
    return pwc_unprepared(0, h0, hks, nullptr, 0, signals, dt, B, K, N, d, 0, U_out, nullptr, gate, workspace, workspace_bytes,
                          stream);
}

int c3b_pwc_gated_supported(int d) {
    return (d == 9 && tuning().force_cta == 0 && d9_gated_supported((int)tuning().d9_variant)) ? 1 : 0;
}

int c3b_pwc_closed_hlist(const void* Hs, double dt, int B, int N, int d, void* U_out, void* dUs_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(B, 0, N, d, U_out, workspace);
    if (rc) return rc;
    if (Hs == nullptr) return fail(C3B_EINVAL, "C3:ERROR: Hs is NULL");
    const Plan pl = make_plan(B, N, d, 0);
    if (workspace_bytes < pl.total) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, pl.total);
    return run_pwc(pl, nullptr, nullptr, static_cast<const cplx*>(Hs), dt, B, 0, N, d, 0, static_cast<cplx*>(U_out),
                   static_cast<cplx*>(dUs_out), nullptr, static_cast<char*>(workspace), static_cast<cudaStream_t>(stream));
}

// ---- prepared models ------------------------------------------------------------------------------------------------
size_t c3b_model_bytes(int K, int d, int lindblad, int n_models) {
    if (K < 0 || d <= 0 || n_models <= 0) return 0;
    return model_layout(K, lindblad ? d * d : d, n_models).total;
}

int c3b_model_prepare(const void* h0, const void* hks, const void* col_ops, int C, double dt, int K, int d, int lindblad,
                      int n_models, void* model_out, size_t model_bytes, void* stream) {
    if (K < 0 || d <= 0 || n_models <= 0) return fail(C3B_EINVAL, "C3:ERROR: bad size (K=%d d=%d n_models=%d)", K, d, n_models);
    if (h0 == nullptr || model_out == nullptr || (K > 0 && hks == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (lindblad && (C < 0 || (C > 0 && col_ops == nullptr))) return fail(C3B_EINVAL, "C3:ERROR: C=%d but col_ops is NULL", C);
    if (lindblad && d > 181) return fail(C3B_EUNSUPPORTED, "C3:ERROR: Lindblad d=%d too large", d);
    if ((reinterpret_cast<uintptr_t>(model_out) & 15) != 0) return fail(C3B_EINVAL, "C3:ERROR: model buffer must be 16-byte aligned");
    const size_t need = c3b_model_bytes(K, d, lindblad, n_models);
    if (model_bytes < need) return fail(C3B_EWORKSPACE, "C3:ERROR: model buffer too small: %zu < %zu bytes", model_bytes, need);
    return build_model(h0, hks, col_ops, C, dt, K, d, lindblad, n_models, static_cast<char*>(model_out), static_cast<cudaStream_t>(stream));
}

size_t c3b_pwc_prepared_workspace_bytes(int B, int N, int d, int lindblad, int n_models) {
    if (B <= 0 || N <= 0 || d <= 0) return 0;
    return make_plan(B, N, lindblad ? d * d : d, n_models > 1 ? 1 : 0).total;
}

int c3b_pwc_prepared(const void* model, const double* signals, int B, int K, int N, int d, int lindblad, int n_models,
                     void* U_out, void* dUs_out, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(B, K, N, d, U_out, workspace);
    if (rc) return rc;
    if (model == nullptr) return fail(C3B_EINVAL, "C3:ERROR: model is NULL");
    if (K > 0 && signals == nullptr) return fail(C3B_EINVAL, "C3:ERROR: K=%d but signals is NULL", K);
    if (n_models != 1 && n_models != B) return fail(C3B_EINVAL, "C3:ERROR: n_models must be 1 or B (got %d, B=%d)", n_models, B);
    const int D = lindblad ? d * d : d;
    const int batched_model = n_models > 1 ? 1 : 0;
    const Plan pl = make_plan(B, N, D, batched_model);
    if (workspace_bytes < pl.total) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, pl.total);
    return run_pwc(pl, static_cast<const char*>(model), signals, nullptr, 0.0, B, K, N, D, batched_model, static_cast<cplx*>(U_out),
                   static_cast<cplx*>(dUs_out), nullptr, static_cast<char*>(workspace), static_cast<cudaStream_t>(stream));
}

// ---- ordered products -------------------------------------------------------------------------------------------------
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
        const int rc = tuning().seq_variant == 0 ? -1
            : launch_seq_small(static_cast<const cplx*>(gates), Gn, seq_idx, seq_len, S, Lmax, D, static_cast<cplx*>(out),
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
    return launch_kron(static_cast<const cplx*>(A), static_cast<const cplx*>(Bm), static_cast<cplx*>(out), batch, ra, ca, rb, cb,
                       a_batched ? (long long)ra * ca : 0, b_batched ? (long long)rb * cb : 0, static_cast<cudaStream_t>(stream));
}

// ---- gradient -------------------------------------------------------------
// variant 1 (default, d <= 16): Frechet derivative of the Taylor scheme on (X, dX) pairs, fused contraction
// variant 0: augmented 2d x 2d exponential through the forward kernels (any d <= 32)
// variant 1 (default): Frechet derivative of the Taylor scheme on (X, dX) pairs, fused contraction -- warp-per-slice kernels for
//   matrix dimension <= 16, CTA kernels on the DMMA product above (variant 2 here)
// variant 0: augmented 2d x 2d exponential through the forward kernels (closed systems, d <= 32; cross-check)
// variant 3 (default for closed d = 7..9 with Hermitian Hamiltonians): the fused lockstep kernel of grad_blk9.cuh -- no stored
//   propagators at all.  "grad_variant" 2 forces the stored-propagator kernels there (cross-check); "grad_unitary" says whether
//   the Hamiltonians are Hermitian (1), are not (0), or must be checked on the device (-1: one 4-byte read-back per call).
static int grad_variant_for(int D) {
    if (tuning().grad_variant == 0 && D <= 32) return 0;
    return D <= 16 ? 1 : 2;
}
constexpr int kGradMaxDim = 128;

// which fused (unitary-recurrence) kernel serves a closed-system call: 0 none, 1 the d = 9 lane-group kernel (grad_blk9.cuh),
// 2 the CTA kernel on the DMMA product for 16 < d <= 32 (grad_ucta.cuh)
static int grad9_wanted(int lindblad, int K, int d) {
    if (lindblad || tuning().grad_variant != 1 || tuning().grad_unitary == 0) return 0;
    if (tuning().force_cta == 0 && grad9_supported(K, d)) return 1;
    if (grad_ucta_supported(d) && pwc_path(d, 0) == 2 && tuning().cta_variant == 1) return 2;
    return 0;
}
// chunk length of the fused gradient kernels: enough chunks to fill the machine (lane groups / CTAs x 4 waves), bounded
static int grad9_chunk_len(int B, int N, int kind) {
    if (tuning().grad_chunk > 0) return (int)(tuning().grad_chunk < N ? tuning().grad_chunk : N);
    // Measured (B200): d = 9 lane-group kernel -- 24 slices per chunk when that fills the machine (12 .. 63 is flat within 2 % at
    // B = 1024), 16 for small batches (B = 1, N = 1000: 0.74 ms; the stored-propagator kernels take 4.3 ms there); CTA kernel
    // (d = 27) -- 4 waves of CTAs, at least 25 slices (B = 16, N = 400: 2.5 ms against 5.4 ms at 6 slices: the boundary kernel
    // is one warp per row and sequential in the chunks), at most 64
    long long cl;
    if (kind == 1) {
        cl = ((long long)B * N / 24 >= 3LL * 8 * num_sms() / 2) ? 24 : 16;
    } else {
        const long long want = 2LL * num_sms() * 4;
        cl = ((long long)B * N + want - 1) / want;
        if (cl < 25) cl = 25;
        if (cl > 64) cl = 64;
    }
    if (cl > N) cl = N;
    return (int)cl;
}
struct Grad9Layout { size_t off_model, off_plan, off_U, off_Y, off_F, off_counter, total; int CL, Q; Plan pl; };
static Grad9Layout grad9_layout(int Bc, int K, int N, int d) {
    Grad9Layout g{};
    g.CL = grad9_chunk_len(Bc, N, grad9_wanted(0, K, d));
    g.Q = (N + g.CL - 1) / g.CL;
    g.pl = make_plan(Bc, N, d, 0, g.CL);
    const ModelLayout ml = model_layout(K, d, 1);
    size_t o = 0;
    g.off_model = o; o += ml.total;
    g.off_plan = o; o += align_up(g.pl.total);
    g.off_U = o; o += align_up((size_t)Bc * d * d * sizeof(cplx));
    g.off_Y = o; o += align_up((size_t)Bc * g.Q * d * d * sizeof(cplx));
    g.off_F = o; o += align_up((size_t)Bc * g.Q * d * d * sizeof(cplx));          // prefix products of the chunk products
    g.off_counter = o; o += 256;
    g.total = o;
    return g;
}

// D: matrix dimension of the propagators (d, or d^2 for Lindblad); dh: Hilbert dimension passed to the workspace query
static size_t grad_chunk_bytes(int Bc, int K, int N, int dh, int lindblad, size_t* off /*[10]*/) {
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
    off[9] = o; if (grad_variant_for(D) == 2) o += grad_cta_workspace_bytes(Bc, N, D);   // per-CTA matrix slots (D > 32)
    return o;
}

size_t c3b_pwc_grad_workspace_bytes(int B, int K, int N, int d, int chunk) {
    if (B <= 0 || N <= 0 || d <= 0 || K <= 0) return 0;
    const int Bc = (chunk > 0 && chunk < B) ? chunk : B;
    size_t off[10];
    if (grad9_wanted(0, K, d)) {
        const size_t fused = grad9_layout(Bc, K, N, d).total;
        if (tuning().grad_unitary == 1) return fused;
        const size_t stored = grad_chunk_bytes(Bc, K, N, d, 0, off);     // the device check may send the call there
        return fused > stored ? fused : stored;
    }
    return grad_chunk_bytes(Bc, K, N, d, 0, off);
}

size_t c3b_pwc_lindblad_grad_workspace_bytes(int B, int K, int N, int d, int chunk) {
    if (B <= 0 || N <= 0 || d <= 0 || K <= 0) return 0;
    const int Bc = (chunk > 0 && chunk < B) ? chunk : B;
    size_t off[10];
    return grad_chunk_bytes(Bc, K, N, d, 1, off);
}

// forward pass of the fused gradient path for one batch chunk: chunk products (segments of CL slices, not folded) and their
// prefix products, which also yield U
static int fused_forward(const Grad9Layout& gl, int nb, const double* sig, double dt, int K, int N, int dh, cplx* Udst, char* w9,
                         cudaStream_t st, int* Q_out, int* CL_out) {
    const Plan pl = make_plan(nb, N, dh, 0, gl.CL);
    int rc = run_pwc(pl, w9 + gl.off_model, sig, nullptr, dt, nb, K, N, dh, 0, Udst, nullptr, nullptr, w9 + gl.off_plan, st, /*fold=*/false);
    if (rc) return rc;
    const cplx* seg = pl.S > 1 ? reinterpret_cast<const cplx*>(w9 + gl.off_plan + pl.off_seg) : Udst;
    *Q_out = pl.S;
    *CL_out = pl.seg_len;
    return launch_grad9_prefix(seg, reinterpret_cast<cplx*>(w9 + gl.off_F), Udst, nb, pl.S, dh, st);
}

// backward pass from what fused_forward left in the workspace: Y at the chunk heads, then the lockstep gradient kernel
static int fused_backward(const Grad9Layout& gl, int nb, const double* sig, int K, int N, int dh, int Q, int CL, const cplx* U,
                          const cplx* Ubar, double* grad, char* w9, cudaStream_t st) {
    const ModelLayout ml = model_layout(K, dh, 1);
    cplx* Yb = reinterpret_cast<cplx*>(w9 + gl.off_Y);
    int rc = launch_grad9_ybound(reinterpret_cast<const cplx*>(w9 + gl.off_F), U, Ubar, Yb, nb, Q, dh, st);
    if (rc) return rc;
    if (grad9_wanted(0, K, dh) == 2) {
        GradUParams gu{};
        gu.G = reinterpret_cast<const cplx*>(w9 + gl.off_model + ml.off_G);
        gu.RS = reinterpret_cast<const double*>(w9 + gl.off_model + ml.off_RS);
        gu.TR = reinterpret_cast<const cplx*>(w9 + gl.off_model + ml.off_TR);
        gu.signals = sig; gu.Ybound = Yb; gu.grad = grad;
        gu.B = nb; gu.K = K; gu.N = N; gu.D = dh; gu.Q = Q; gu.CL = CL;
        return launch_grad_ucta(gu, st);
    }
    Grad9Params gp{};
    gp.G = reinterpret_cast<const cplx*>(w9 + gl.off_model + ml.off_G);
    gp.RS = reinterpret_cast<const double*>(w9 + gl.off_model + ml.off_RS);
    gp.TR = reinterpret_cast<const cplx*>(w9 + gl.off_model + ml.off_TR);
    gp.signals = sig; gp.Ybound = Yb; gp.grad = grad;
    gp.B = nb; gp.K = K; gp.N = N; gp.d = dh; gp.Q = Q; gp.CL = CL;
    return launch_grad9(gp, reinterpret_cast<unsigned int*>(w9 + gl.off_counter), st);
}

static int pwc_grad_impl(int lindblad, const void* h0, const void* hks, const void* col_ops, int C, const double* signals,
                         double dt, int B, int K, int N, int dh, const void* Ubar, double* grad_out, void* U_out,
                         int chunk, void* workspace, size_t workspace_bytes, void* stream) {
    if (B <= 0 || N <= 0 || dh <= 0 || K <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d N=%d d=%d)", B, K, N, dh);
    if (!h0 || !hks || !signals || !Ubar || !grad_out || !workspace) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    const int d = lindblad ? dh * dh : dh;                      // dimension of the propagators
    if (d > kGradMaxDim) return fail(C3B_EUNSUPPORTED, "C3:ERROR: gradient path supports matrix dimension <= %d (got %d)", kGradMaxDim, d);
    const int variant = grad_variant_for(d);
    if (lindblad && variant == 0)
        return fail(C3B_EUNSUPPORTED, "C3:ERROR: the augmented-exponential gradient (grad_variant 0) is built for closed systems only");
    const int Bc = (chunk > 0 && chunk < B) ? chunk : B;
    if (grad9_wanted(lindblad, K, dh)) {
        cudaStream_t st9 = static_cast<cudaStream_t>(stream);
        const Grad9Layout gl = grad9_layout(Bc, K, N, dh);
        if (workspace_bytes < gl.total) return fail(C3B_EWORKSPACE, "C3:ERROR: workspace too small: %zu < %zu bytes", workspace_bytes, gl.total);
        char* w9 = static_cast<char*>(workspace);
        unsigned int* flag = reinterpret_cast<unsigned int*>(w9 + gl.off_counter) + 16;
        bool unitary = tuning().grad_unitary == 1;
        if (!unitary) {                                // -1: look at the Hamiltonians (one small kernel + a 4-byte read-back)
            int rc = launch_hermitian_check(static_cast<const cplx*>(h0), static_cast<const cplx*>(hks), K, dh, flag, st9);
            if (rc) return rc;
            unsigned int host_flag = 1;
            CUDA_TRY(cudaMemcpyAsync(&host_flag, flag, sizeof(host_flag), cudaMemcpyDeviceToHost, st9));
            CUDA_TRY(cudaStreamSynchronize(st9));
            unitary = host_flag == 0;
        }
        if (unitary) {
            int rc = build_model(h0, hks, nullptr, 0, dt, K, dh, 0, 1, w9 + gl.off_model, st9);
            if (rc) return rc;
            for (int b0 = 0; b0 < B; b0 += Bc) {
                const int nb = (B - b0 < Bc) ? (B - b0) : Bc;
                const double* sig = signals + (size_t)b0 * K * N;
                cplx* Udst = U_out ? static_cast<cplx*>(U_out) + (size_t)b0 * dh * dh : reinterpret_cast<cplx*>(w9 + gl.off_U);
                int Q = 0, CLrun = 0;
                if ((rc = fused_forward(gl, nb, sig, dt, K, N, dh, Udst, w9, st9, &Q, &CLrun)) != 0) return rc;
                rc = fused_backward(gl, nb, sig, K, N, dh, Q, CLrun, Udst, static_cast<const cplx*>(Ubar) + (size_t)b0 * dh * dh,
                                    grad_out + (size_t)b0 * K * N, w9, st9);
                if (rc) return rc;
            }
            return C3B_OK;
        }
    }
    size_t off[10];
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
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int nb = (B - b0 < Bc) ? (B - b0) : Bc;
        const double* sig = signals + (size_t)b0 * K * N;
        cplx* Udst = U_out ? static_cast<cplx*>(U_out) + (size_t)b0 * dd : Utmp;
        int rc = lindblad
            ? c3b_pwc_lindblad(h0, hks, col_ops, C, sig, dt, nb, K, N, dh, 0, Udst, dUs, ws + off[0], off[1] - off[0], stream)
            : c3b_pwc_closed(h0, hks, sig, dt, nb, K, N, dh, 0, Udst, dUs, ws + off[0], off[1] - off[0], stream);
        if (rc) return rc;
        const cplx* ub = static_cast<const cplx*>(Ubar) + (size_t)b0 * dd;
        if (variant == 2) {
            const ModelLayout ml = model_layout(K, d, 1);
            const char* fws = ws + off[0];
            rc = launch_grad_cta(reinterpret_cast<const cplx*>(fws + ml.off_G), reinterpret_cast<const double*>(fws + ml.off_RS),
                                 reinterpret_cast<const cplx*>(fws + ml.off_TR), sig, dUs, ub, Psi, alpha, grad_out + (size_t)b0 * K * N,
                                 nb, K, N, d, reinterpret_cast<cplx*>(ws + off[9]), st);
            if (rc) return rc;
            continue;
        }
        if ((rc = launch_grad_suffix(variant, dUs, ub, Psi, alpha, nb, N, d, st)) != 0) return rc;
        if (variant == 1) {
            if ((rc = launch_grad_prefix_frechet(dUs, Psi, nb, N, d, st)) != 0) return rc;
            // generators, row sums and trace shifts as the forward call left them at the head of its workspace
            const ModelLayout ml = model_layout(K, d, 1);
            const char* fws = ws + off[0];
            rc = launch_grad_frechet(reinterpret_cast<const cplx*>(fws + ml.off_G), reinterpret_cast<const double*>(fws + ml.off_RS),
                                     reinterpret_cast<const cplx*>(fws + ml.off_TR), sig, Psi, alpha,
                                     grad_out + (size_t)b0 * K * N, nb, K, N, d, st);
            if (rc) return rc;
            continue;
        }
        rc = launch_grad_prefix_aug(dUs, Psi, static_cast<const cplx*>(h0), static_cast<const cplx*>(hks), sig, Haug, dt, nb, K, N, d, st);
        if (rc) return rc;
        rc = c3b_pwc_closed_hlist(Haug, dt, nb, N, 2 * d, Uaug, Eaug, ws + off[1], off[2] - off[1], stream);
        if (rc) return rc;
        rc = launch_grad_contract(Eaug, static_cast<const cplx*>(hks), alpha, grad_out + (size_t)b0 * K * N, dt, nb, K, N, d, st);
        if (rc) return rc;
    }
    return C3B_OK;
}

int c3b_pwc_closed_grad(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                        const void* Ubar, double* grad_out, void* U_out, int chunk, void* workspace,
                        size_t workspace_bytes, void* stream) {
    return pwc_grad_impl(0, h0, hks, nullptr, 0, signals, dt, B, K, N, d, Ubar, grad_out, U_out, chunk, workspace,
                         workspace_bytes, stream);
}

// ---- forward / backward as two calls (what an autograd node needs): the forward call leaves the model blob, the chunk
//      products and their prefix products in a caller-owned state buffer, the backward call turns a cotangent into the gradient
//      without recomputing the forward pass
constexpr size_t kSavedHeaderBytes = 0;

size_t c3b_pwc_closed_saved_bytes(int B, int K, int N, int d) {
    if (B <= 0 || K <= 0 || N <= 0 || d <= 0 || !grad9_wanted(0, K, d)) return 0;
    return kSavedHeaderBytes + grad9_layout(B, K, N, d).total;
}

int c3b_pwc_closed_fwd_saved(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d, void* U_out,
                             void* state, size_t state_bytes, void* stream) {
    if (B <= 0 || K <= 0 || N <= 0 || d <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size (B=%d K=%d N=%d d=%d)", B, K, N, d);
    if (!h0 || !hks || !signals || !U_out || !state) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (!grad9_wanted(0, K, d))
        return fail(C3B_EUNSUPPORTED, "C3:ERROR: no fused gradient kernel for this shape (d=%d, K=%d): use c3b_pwc_closed_grad", d, K);
    const Grad9Layout gl = grad9_layout(B, K, N, d);
    if (state_bytes < kSavedHeaderBytes + gl.total)
        return fail(C3B_EWORKSPACE, "C3:ERROR: state buffer too small: %zu < %zu bytes", state_bytes, kSavedHeaderBytes + gl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* w9 = static_cast<char*>(state) + kSavedHeaderBytes;
    if (tuning().grad_unitary != 1) {
        unsigned int* flag = reinterpret_cast<unsigned int*>(w9 + gl.off_counter) + 16;
        int rc = launch_hermitian_check(static_cast<const cplx*>(h0), static_cast<const cplx*>(hks), K, d, flag, st);
        if (rc) return rc;
        unsigned int host_flag = 1;
        CUDA_TRY(cudaMemcpyAsync(&host_flag, flag, sizeof(host_flag), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (host_flag != 0) return fail(C3B_EUNSUPPORTED, "C3:ERROR: the Hamiltonians are not Hermitian: use c3b_pwc_closed_grad");
    }
    int rc = build_model(h0, hks, nullptr, 0, dt, K, d, 0, 1, w9 + gl.off_model, st);
    if (rc) return rc;
    cplx* Ust = reinterpret_cast<cplx*>(w9 + gl.off_U);
    int Q = 0, CL = 0;
    if ((rc = fused_forward(gl, B, signals, dt, K, N, d, Ust, w9, st, &Q, &CL)) != 0) return rc;
    CUDA_TRY(cudaMemcpyAsync(U_out, Ust, (size_t)B * d * d * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    return C3B_OK;
}

int c3b_pwc_closed_bwd_saved(const double* signals, int B, int K, int N, int d, int Q, int CL, const void* Ubar, double* grad_out,
                             void* state, size_t state_bytes, void* stream) {
    if (B <= 0 || K <= 0 || N <= 0 || d <= 0 || Q <= 0 || CL <= 0) return fail(C3B_EINVAL, "C3:ERROR: non-positive size");
    if (!signals || !Ubar || !grad_out || !state) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (!grad9_wanted(0, K, d)) return fail(C3B_EUNSUPPORTED, "C3:ERROR: no fused gradient kernel for this shape (d=%d, K=%d)", d, K);
    const Grad9Layout gl = grad9_layout(B, K, N, d);
    if (state_bytes < kSavedHeaderBytes + gl.total)
        return fail(C3B_EWORKSPACE, "C3:ERROR: state buffer too small: %zu < %zu bytes", state_bytes, kSavedHeaderBytes + gl.total);
    if (Q > gl.Q) return fail(C3B_EINVAL, "C3:ERROR: Q=%d does not belong to this state (at most %d chunks)", Q, gl.Q);
    char* w9 = static_cast<char*>(state) + kSavedHeaderBytes;
    return fused_backward(gl, B, signals, K, N, d, Q, CL, reinterpret_cast<const cplx*>(w9 + gl.off_U), static_cast<const cplx*>(Ubar), grad_out, w9,
                          static_cast<cudaStream_t>(stream));
}

int c3b_pwc_closed_saved_chunks(int B, int K, int N, int d, int* Q_out, int* CL_out) {
    if (!Q_out || !CL_out) return fail(C3B_EINVAL, "C3:ERROR: NULL pointer");
    if (B <= 0 || K <= 0 || N <= 0 || d <= 0 || !grad9_wanted(0, K, d)) return fail(C3B_EUNSUPPORTED, "C3:ERROR: no fused gradient kernel for this shape");
    const Grad9Layout gl = grad9_layout(B, K, N, d);
    const Plan pl = make_plan(B, N, d, 0, gl.CL);
    *Q_out = pl.S;
    *CL_out = pl.seg_len;
    return C3B_OK;
}

int c3b_pwc_lindblad_grad(const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                          int B, int K, int N, int d, const void* Ubar, double* grad_out, void* U_out, int chunk,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (C < 0 || (C > 0 && col_ops == nullptr)) return fail(C3B_EINVAL, "C3:ERROR: C=%d but col_ops is NULL", C);
    return pwc_grad_impl(1, h0, hks, col_ops, C, signals, dt, B, K, N, d, Ubar, grad_out, U_out, chunk, workspace,
                         workspace_bytes, stream);
}

long long c3b_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

double c3b_last_kernel_ms(void) {
    if (!g_prof.valid) return (double)fail(C3B_EINVAL, "C3:ERROR: no profiled launch on this thread (set tuning 'profile' to 1 first)");
    if (cudaEventSynchronize(g_prof.e1) != cudaSuccess) return (double)fail(C3B_ECUDA, "C3:ERROR: event synchronize failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.e0, g_prof.e1) != cudaSuccess) return (double)fail(C3B_ECUDA, "C3:ERROR: event elapsed failed");
    return (double)ms;
}

}  // extern "C"
