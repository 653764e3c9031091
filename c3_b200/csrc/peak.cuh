// fp64 peak micro-benchmarks: the roofline denominators for the propagator kernels.
// MEASURED_PEAKS.json only carries HBM and bf16; SURVEY.md section 7 step 1 asks for the fp64
// DFMA and DMMA peaks to be measured on the box.
#pragma once
#include <cuda_runtime.h>
#include "c3b_common.cuh"

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
