// Micro-benchmarks that decide the lane/block layout of the small-d propagator kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb mb.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef double2 cplx;
__device__ __forceinline__ cplx cmake(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ void cfma(cplx& c, const cplx a, const cplx b) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
__device__ __forceinline__ cplx lds_c(const cplx* p, int mode) {
    if (mode == 0) return *p;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    cplx v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v.x) : "r"(a));
    asm("ld.shared.f64 %0, [%1+8];" : "=d"(v.y) : "r"(a));
    return v;
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// ------------------------------------------------------------------------------------------
// 1. LDS throughput by address pattern.  pattern: 0 all-same, 1 three groups of 9 (+5 alias),
//    2 all distinct consecutive, 3 eight groups of 4, 4 five groups of 6 (+2), 5: 16 groups of 2
// ------------------------------------------------------------------------------------------
template <int VEC>  // 16 or 8 bytes per lane
__global__ void lds_bench(long long* cyc, double* sink, int iters, int pattern) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8192 / 8; i += blockDim.x) ((double*)sm)[i] = i;
    __syncthreads();
    int idx;
    switch (pattern) {
        case 0: idx = 0; break;
        case 1: idx = (lane < 27 ? lane / 9 : 0) * 85; break;
        case 2: idx = lane; break;
        case 3: idx = (lane / 4) * 21; break;
        case 4: idx = (lane < 30 ? lane / 6 : 0) * 53; break;
        default: idx = (lane / 2) * 11; break;
    }
    const unsigned char* base = sm + (size_t)idx * VEC + warp * 16;
    double acc0 = 0, acc1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const unsigned addr = (unsigned)__cvta_generic_to_shared(base + u * 256);
            if (VEC == 16) {
                double vx, vy;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
                acc0 += vx; acc1 += vy;
            } else {
                double v;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
                acc0 += v;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc0 + acc1 == 1.2345) sink[0] = acc0;
}

// ------------------------------------------------------------------------------------------
// 2. complex D x D matmul primitives, chained X <- X * Y with a shared-memory round trip and a
//    __syncwarp per product (as in the real Pade power chain).
// ------------------------------------------------------------------------------------------
// (a) row layout: lane = one row, X row in registers, Y broadcast from smem.  ROWS rows per lane.
template <int D, int ROWS, int MODE, int ALIAS>
__global__ void mm_rows_bench(double* sink, int iters) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int LPM = (D + ROWS - 1) / ROWS;  // lanes per matrix
    constexpr int MPW = 32 / LPM;
    constexpr int MS = D * D + 4;                // matrix stride (cplx) incl. pad
    cplx* sm = reinterpret_cast<cplx*>(smraw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_raw = lane / LPM;
    const bool on = m_raw < MPW;
    const int m = on ? m_raw : (ALIAS ? MPW - 1 : 0);
    const int li = on ? lane - m_raw * LPM : (ALIAS ? LPM - 1 : 0);
    cplx* Y = sm + ((size_t)warp * MPW + m) * MS;
    for (int e = lane; e < MPW * MS; e += 32) sm[(size_t)warp * MPW * MS + e] = cmake(0.05 * ((e % 7) - 3), 0.03 * ((e % 5) - 2));
    __syncwarp();
    cplx X[ROWS][D];
#pragma unroll
    for (int a = 0; a < ROWS; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) X[a][j] = cmake(0.1 * (j == li), 0.01 * j);
    for (int it = 0; it < iters; ++it) {
        cplx C[ROWS][D];
#pragma unroll
        for (int a = 0; a < ROWS; ++a)
#pragma unroll
            for (int j = 0; j < D; ++j) C[a][j] = cmake(0, 0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const cplx y = lds_c(&Y[k * D + j], MODE);
#pragma unroll
                for (int a = 0; a < ROWS; ++a) cfma(C[a][j], X[a][k], y);
            }
        }
        __syncwarp();
#pragma unroll
        for (int a = 0; a < ROWS; ++a) {
            const int row = li * ROWS + a;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                X[a][j] = C[a][j];
                if (on && row < D) Y[row * D + j] = cmake(C[a][j].x * 0.5, C[a][j].y * 0.5);
            }
        }
        __syncwarp();
    }
    double s = 0;
#pragma unroll
    for (int a = 0; a < ROWS; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) s += X[a][j].x + X[a][j].y;
    if (s == 1.2345) sink[0] = s;
}

// (b) block layout: GR x GC lanes per matrix, lane owns an R x CC block of C; both operands
//     streamed from shared memory; result stored back as the next X operand.
template <int D, int GR, int GC, int PAD, int MODE>
__global__ void mm_block_bench(double* sink, int iters) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int R = (D + GR - 1) / GR, CC = (D + GC - 1) / GC;
    constexpr int LPM = GR * GC, MPW = 32 / LPM;
    constexpr int ROWS = GR * R, LD = GC * CC;   // padded extents
    constexpr int MS = ROWS * LD + PAD;          // per-matrix buffer (cplx)
    cplx* sm = reinterpret_cast<cplx*>(smraw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_raw = lane / LPM;
    const bool on = m_raw < MPW;
    const int m = on ? m_raw : 0;
    const int li = on ? lane - m_raw * LPM : 0;
    const int gi = li / GC, gj = li - gi * GC;
    cplx* base = sm + ((size_t)warp * MPW + m) * 3 * MS;  // X0, X1 (ping-pong), Y
    for (int e = lane; e < MPW * 3 * MS; e += 32) sm[(size_t)warp * MPW * 3 * MS + e] = cmake(0.05 * ((e % 7) - 3), 0.03 * ((e % 5) - 2));
    __syncwarp();
    const cplx* Y = base + 2 * MS + gj * CC;
    cplx acc_keep = cmake(0, 0);
    for (int it = 0; it < iters; ++it) {
        const cplx* X = base + (it & 1) * MS + gi * R * LD;
        cplx* Xn = base + ((it & 1) ^ 1) * MS + gi * R * LD + gj * CC;
        cplx C[R][CC];
#pragma unroll
        for (int a = 0; a < R; ++a)
#pragma unroll
            for (int c = 0; c < CC; ++c) C[a][c] = cmake(0, 0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            cplx x[R], y[CC];
#pragma unroll
            for (int a = 0; a < R; ++a) x[a] = lds_c(&X[a * LD + k], MODE);
#pragma unroll
            for (int c = 0; c < CC; ++c) y[c] = lds_c(&Y[k * LD + c], MODE);
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int c = 0; c < CC; ++c) cfma(C[a][c], x[a], y[c]);
        }
        if (on) {
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int c = 0; c < CC; ++c) Xn[a * LD + c] = C[a][c];
        }
        acc_keep.x += C[0][0].x;
        __syncwarp();
    }
    if (acc_keep.x == 1.2345) sink[0] = acc_keep.x;
}

// ------------------------------------------------------------------------------------------
// 3. DFMA and DMMA issued together: are they one pipe or two?
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NDMMA, int NDFMA>
__global__ void __launch_bounds__(256) mix_bench(double* sink, int iters, double a, double b) {
    double c[8][2], x[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = a * i; c[i][1] = b * i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NDMMA; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < NDFMA; ++i) x[i] = fma(x[i], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 1.2345) sink[0] = s;
}

// ------------------------------------------------------------------------------------------
static int g_sms = 148;
static double* g_sink;

template <typename F>
static float time_ms(F launch, int reps = 3) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int D, int ROWS, int MODE, int ALIAS>
static void run_rows(const char* name, int warps, int ctas) {
    constexpr int LPM = (D + ROWS - 1) / ROWS, MPW = 32 / LPM, MS = D * D + 4;
    size_t smem = (size_t)warps * MPW * MS * sizeof(cplx);
    auto k = mm_rows_bench<D, ROWS, MODE, ALIAS>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, warps * 32, smem);
    const int iters = 3000;
    float ms = time_ms([&] { k<<<g_sms * ctas, warps * 32, smem>>>(g_sink, iters); });
    double useful = 8.0 * D * D * D * (double)iters * MPW * warps * ctas * g_sms / (ms * 1e-3) / 1e12;
    printf("%-22s mode=%d alias=%d warps=%2d ctas=%d occ=%d smem=%6zu  useful %.2f TF  (lanes %d/32)\n", name, MODE, ALIAS, warps, ctas, occ, smem, useful, MPW * LPM);
}

template <int D, int GR, int GC, int PAD, int MODE>
static void run_block(const char* name, int warps, int ctas) {
    constexpr int R = (D + GR - 1) / GR, CC = (D + GC - 1) / GC, LPM = GR * GC, MPW = 32 / LPM;
    constexpr int MS = GR * R * GC * CC + PAD;
    size_t smem = (size_t)warps * MPW * 3 * MS * sizeof(cplx);
    auto k = mm_block_bench<D, GR, GC, PAD, MODE>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, warps * 32, smem);
    const int iters = 3000;
    float ms = time_ms([&] { k<<<g_sms * ctas, warps * 32, smem>>>(g_sink, iters); });
    double useful = 8.0 * D * D * D * (double)iters * MPW * warps * ctas * g_sms / (ms * 1e-3) / 1e12;
    printf("%-22s warps=%2d ctas=%d occ=%d smem=%6zu  useful %.2f TF  (lanes %d/32, blk %dx%d pad %d mode %d)\n", name, warps, ctas, occ, smem, useful, MPW * LPM, R, CC, PAD, MODE);
}

template <int VEC>
static void run_lds(int pattern, int warps) {
    long long* cyc; CK(cudaMalloc(&cyc, g_sms * sizeof(long long)));
    const int iters = 2000;
    lds_bench<VEC><<<g_sms, warps * 32, 16384>>>(cyc, g_sink, iters, pattern);
    CK(cudaDeviceSynchronize());
    lds_bench<VEC><<<g_sms, warps * 32, 16384>>>(cyc, g_sink, iters, pattern);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(g_sms);
    cudaMemcpy(h.data(), cyc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= g_sms;
    printf("LDS.%d pattern %d warps %2d: %.2f SM-cycles per warp-instruction\n", VEC * 8, pattern, warps, avg / ((double)iters * 16 * warps));
    cudaFree(cyc);
}

template <int NDMMA, int NDFMA>
static void run_mix() {
    const int iters = 4000, grid = g_sms * 8;
    float ms = time_ms([&] { mix_bench<NDMMA, NDFMA><<<grid, 256>>>(g_sink, iters, 1.0000001, 0.9999999); });
    double dm = 2.0 * 256 * NDMMA * 4.0 * iters * grid * 8 / (ms * 1e-3) / 1e12;
    double df = 2.0 * 32 * NDFMA * 4.0 * iters * grid * 8 / (ms * 1e-3) / 1e12;
    printf("mix DMMA x%d + DFMA x%2d : DMMA %.2f TF + DFMA %.2f TF = %.2f TF\n", NDMMA, NDFMA, dm, df, dm + df);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, g_sms);
    CK(cudaMalloc(&g_sink, 1024));
    for (int w : {4, 8}) for (int c : {1, 2, 3}) {
        if (w * c > 16) continue;
        run_rows<9, 1, 0, 0>("rows d9 1row/lane", w, c);
        run_rows<9, 1, 0, 1>("rows d9 1row/lane", w, c);
        run_rows<9, 1, 1, 0>("rows d9 1row/lane", w, c);
        run_rows<9, 1, 1, 1>("rows d9 1row/lane", w, c);
    }
    for (int w : {4, 8}) for (int c : {1, 2}) {
        if (w * c > 8) continue;
        run_rows<9, 2, 0, 0>("rows d9 2rows/lane", w, c);
        run_rows<9, 2, 1, 0>("rows d9 2rows/lane", w, c);
        run_rows<9, 2, 1, 1>("rows d9 2rows/lane", w, c);
        run_rows<9, 3, 1, 0>("rows d9 3rows/lane", w, c);
    }
    for (int w : {4, 8}) for (int c : {1, 2}) {
        run_block<9, 3, 3, 4, 0>("block d9 3x3 lanes", w, c);
        run_block<9, 3, 3, 4, 1>("block d9 3x3 lanes", w, c);
        run_block<9, 3, 3, 0, 1>("block d9 3x3 lanes", w, c);
        if (w * c <= 8) run_block<9, 2, 3, 3, 1>("block d9 2x3 lanes", w, c);
        if (w * c <= 4) run_block<9, 2, 2, 0, 1>("block d9 2x2 lanes", w, c);
    }
    run_rows<3, 1, 0, 0>("rows d3", 8, 2); run_rows<3, 1, 1, 0>("rows d3", 8, 2);
    run_rows<4, 1, 0, 0>("rows d4", 8, 2); run_rows<4, 1, 1, 0>("rows d4", 8, 2);
    run_rows<8, 1, 0, 0>("rows d8", 8, 2); run_rows<8, 1, 1, 0>("rows d8", 8, 2);
    run_rows<12, 1, 0, 0>("rows d12", 8, 1); run_rows<12, 1, 1, 0>("rows d12", 8, 1);
    run_rows<16, 1, 1, 0>("rows d16", 8, 1);
    return 0;
}
