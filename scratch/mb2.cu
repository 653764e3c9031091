// Micro-benchmark 2: is SHFL a separate resource from the shared-memory wavefront pipe?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// NSH shuffles (32-bit) and NLD LDS.128 (3-groups-of-9 broadcast pattern) and NF DFMA per inner step
template <int NSH, int NLD, int NF>
__global__ void __launch_bounds__(256) mixk(double* sink, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) ((double*)sm)[i] = i * 1e-3;
    __syncthreads();
    const int g = lane < 27 ? lane / 9 : 2;
    const double2* base = reinterpret_cast<const double2*>(sm) + g * 85 + warp * 3;
    unsigned v[8];
    double x[8], acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = lane * 7 + i; x[i] = 1.0 + i * 1e-3; }
    int src = (lane * 5 + 3) & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NSH; ++i) v[i & 7] = __shfl_sync(0xffffffffu, v[i & 7], src) + 1;
#pragma unroll
            for (int i = 0; i < NLD; ++i) {
                const unsigned a = (unsigned)__cvta_generic_to_shared(base + (u * NLD + i) * 1);
                double vx, vy;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(a));
                acc[i & 3] += vx + vy;
            }
#pragma unroll
            for (int i = 0; i < NF; ++i) x[i & 7] = fma(x[i & 7], 0.9999999, 1e-9);
        }
        src = (src + 1) & 31;
    }
    double s = acc[0] + acc[1] + acc[2] + acc[3];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] + x[i];
    if (s == 1.2345) sink[0] = s;
}

static int g_sms;
static double* g_sink;
template <int NSH, int NLD, int NF>
static void run(int warps, int ctas) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    auto k = mixk<NSH, NLD, NF>;
    k<<<g_sms * ctas, warps * 32, 32768>>>(g_sink, iters); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<<<g_sms * ctas, warps * 32, 32768>>>(g_sink, iters); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // SM cycles per inner step per warp, assuming 1.965 GHz
    double cyc = best * 1e-3 * 1.965e9 / ((double)iters * 4 * warps * ctas);
    printf("SHFL %2d  LDS.128 %2d  DFMA %2d | warps/SM %2d : %.2f SM-clk per warp-step  (shfl %.2f, lds %.2f, dfma %.2f clk each if alone)\n",
           NSH, NLD, NF, warps * ctas, cyc, NSH ? cyc / NSH : 0.0, NLD ? cyc / NLD : 0.0, NF ? cyc / NF : 0.0);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); g_sms = prop.multiProcessorCount;
    CK(cudaMalloc(&g_sink, 1024));
    for (int w : {8, 16, 32}) {
        run<16, 0, 0>(8, w / 8);
        run<0, 16, 0>(8, w / 8);
        run<0, 0, 32>(8, w / 8);
        run<16, 16, 0>(8, w / 8);
        run<16, 4, 0>(8, w / 8);
        run<8, 8, 32>(8, w / 8);
        run<0, 8, 32>(8, w / 8);
        run<16, 0, 32>(8, w / 8);
        run<16, 4, 32>(8, w / 8);
    }
    return 0;
}
