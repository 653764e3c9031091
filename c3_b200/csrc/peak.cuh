// fp64 peak micro-benchmarks: the roofline denominators for the propagator kernels.
// MEASURED_PEAKS.json only carries HBM and bf16; SURVEY.md section 7 step 1 asks for the fp64
// DFMA and DMMA peaks to be measured on the box.
#pragma once
#include <cuda_runtime.h>
#include "pwc_rows.cuh"

namespace c3b {

// 16 independent DFMA chains per thread, fully unrolled
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i * 1e-3 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma(x[i], b, a);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// 8 independent m8n8k4 accumulator tiles per warp
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = a * i; c[i][1] = b * i; }
    const double av = a + threadIdx.x * 1e-9, bv = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma_m8n8k4(c[i][0], c[i][1], av, bv);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The core primitive of the register-resident kernel in isolation: C(row) = X(row) * Y with Y
// streamed from shared memory by broadcast LDS.128 (floor(32/D) lane groups per warp).
// Reports fp64-PIPE throughput (all 32 lanes counted), so the ratio to the DFMA peak says how
// much the shared-memory operand traffic costs.
template <int D>
__global__ void mmrow_bench_kernel(double* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int G = 32 / D;
    const int g_raw = lane / D;
    const int g = g_raw < G ? g_raw : 0;
    const int r = g_raw < G ? lane - g_raw * D : 0;
    cplx* Y = sm + ((size_t)warp * G + g) * (D * D + 4);
    for (int e = lane; e < G * (D * D + 4); e += 32) sm[(size_t)warp * G * (D * D + 4) + e] = cmake(1e-3 * (e % 7), 1e-3 * (e % 5));
    __syncwarp();
    cplx X[D];
#pragma unroll
    for (int j = 0; j < D; ++j) X[j] = cmake(0.1 * (j + r), 0.01 * j);
    for (int it = 0; it < iters; ++it) {
        cplx C[D];
        mm_row<D>(X, Y, C);
#pragma unroll
        for (int j = 0; j < D; ++j) X[j] = C[j];
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) s += X[j].x + X[j].y;
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int D>
inline int run_mmrow_bench(int warps_per_cta, int ctas_per_sm, double* tflops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr int G = 32 / D;
    const size_t smem = (size_t)warps_per_cta * G * (D * D + 4) * sizeof(cplx);
    auto kern = mmrow_bench_kernel<D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    double* out = nullptr;
    const int grid = sms * ctas_per_sm, block = warps_per_cta * 32;
    if ((e = cudaMalloc(&out, (size_t)grid * block * sizeof(double))) != cudaSuccess) return (int)e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 4000;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0);
        kern<<<grid, block, smem>>>(out, iters);
        cudaEventRecord(t1);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 8.0 * D * D * 32.0 * (double)iters * grid * warps_per_cta;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(out);
    if (e != cudaSuccess) return (int)e;
    *tflops = best;
    return 0;
}

// returns 0 on success; *tflops = achieved TFLOP/s (FMA = 2 flops)
inline int measure_fp64_peak(int kind, int device, double seconds, double* tflops) {
    cudaError_t e;
    int prev = 0;
    if ((e = cudaGetDevice(&prev)) != cudaSuccess) return (int)e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return (int)e;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double* out = nullptr;
    const int grid = sms * 8, block = 256;
    if ((e = cudaMalloc(&out, (size_t)grid * block * sizeof(double))) != cudaSuccess) return (int)e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    int iters = 2000;
    double best = 0.0, elapsed_total = 0.0;
    for (int rep = 0; rep < 64 && elapsed_total < seconds; ++rep) {
        cudaEventRecord(t0);
        if (kind == 0) peak_dfma_kernel<<<grid, block>>>(out, iters, 1.0000001, 0.9999999);
        else peak_dmma_kernel<<<grid, block>>>(out, iters, 1.0000001, 0.9999999);
        cudaEventRecord(t1);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        double flops;
        if (kind == 0) flops = 2.0 * 16 * 8 * (double)iters * grid * block;
        else flops = 2.0 * (8.0 * 8 * 4) * 8 * 4 * (double)iters * grid * (block / 32);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        elapsed_total += ms * 1e-3;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(out);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return (int)e;
    *tflops = best;
    return 0;
}

}  // namespace c3b
