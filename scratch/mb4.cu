// Micro-benchmark 4: DFMA issue rate vs operand pattern (register-file bandwidth / bank limits).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
typedef double2 cplx;
__device__ __forceinline__ void cfma(cplx& c, const cplx a, const cplx b) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
// MODE 0: x[i] = fma(x[i], b, a)            (1 distinct source + 2 loop-invariant)
// MODE 1: c[i][j] = fma(a[i], b[j], c[i][j])  real outer product, NA x NB accumulators
// MODE 2: complex outer product block (cfma), NA x NB complex accumulators, operands rotate each step
template <int MODE, int NA, int NB>
__global__ void __launch_bounds__(128) k(double* sink, int iters, double s0) {
    double acc[NA][NB][2];
    double a[NA][2], b[NB][2];
#pragma unroll
    for (int i = 0; i < NA; ++i) { a[i][0] = s0 + i * 1e-3 + threadIdx.x * 1e-7; a[i][1] = s0 - i * 1e-3; }
#pragma unroll
    for (int j = 0; j < NB; ++j) { b[j][0] = 1.0 - j * 1e-3; b[j][1] = 1e-3 * j; }
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) { acc[i][j][0] = i; acc[i][j][1] = j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int j = 0; j < NB; ++j) { acc[i][j][0] = fma(acc[i][j][0], b[0][0], a[0][0]); acc[i][j][1] = fma(acc[i][j][1], b[0][0], a[0][0]); }
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int j = 0; j < NB; ++j) { acc[i][j][0] = fma(a[i][0], b[j][0], acc[i][j][0]); acc[i][j][1] = fma(a[i][1], b[j][1], acc[i][j][1]); }
            } else {
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        cplx c = make_double2(acc[i][j][0], acc[i][j][1]);
                        cfma(c, make_double2(a[i][0], a[i][1]), make_double2(b[j][0], b[j][1]));
                        acc[i][j][0] = c.x; acc[i][j][1] = c.y;
                    }
            }
            // rotate operands so that nothing is loop invariant (mimics freshly loaded operands)
            if (MODE != 0) {
                double t0 = a[0][0], t1 = a[0][1];
#pragma unroll
                for (int i = 0; i + 1 < NA; ++i) { a[i][0] = a[i + 1][0]; a[i][1] = a[i + 1][1]; }
                a[NA - 1][0] = b[0][0]; a[NA - 1][1] = b[0][1];
#pragma unroll
                for (int j = 0; j + 1 < NB; ++j) { b[j][0] = b[j + 1][0]; b[j][1] = b[j + 1][1]; }
                b[NB - 1][0] = t0; b[NB - 1][1] = t1;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 1.2345) sink[0] = s;
}
static int g_sms; static double* g_sink;
template <int MODE, int NA, int NB>
static void run(const char* name, int warps_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    const int ctas = warps_per_sm / 4;
    auto kk = k<MODE, NA, NB>;
    kk<<<g_sms * ctas, 128>>>(g_sink, iters, 0.999); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); kk<<<g_sms * ctas, 128>>>(g_sink, iters, 0.999); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double dfma_per_thread = (double)iters * 8 * NA * NB * (MODE == 2 ? 4 : 2);
    double tf = 2.0 * dfma_per_thread * 128 * g_sms * ctas / (best * 1e-3) / 1e12;
    printf("%-34s NAxNB=%dx%d warps/SM=%2d : %.2f TFLOP/s  (%.2f clk per warp-DFMA per SMSP)\n", name, NA, NB, warps_per_sm, tf,
           best * 1e-3 * 1.965e9 / (dfma_per_thread * (warps_per_sm / 4.0)));
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); g_sms = prop.multiProcessorCount;
    CK(cudaMalloc(&g_sink, 1024));
    for (int w : {4, 8, 12, 16, 32}) {
        run<0, 3, 3>("x=fma(x,b,a) invariant b,a", w);
        run<1, 3, 3>("real outer product", w);
        run<2, 3, 3>("complex outer product", w);
        run<2, 1, 9>("complex row (1x9)", w);
        run<2, 2, 9>("complex 2 rows (2x9)", w);
        run<2, 3, 9>("complex 3 rows (3x9)", w);
    }
    return 0;
}
