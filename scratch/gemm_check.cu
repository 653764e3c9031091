// compile-only harness: register/spill report of the DMMA CTA kernel variants (nvcc -cubin -Xptxas -v)
#include "../c3_b200/csrc/pwc_gemm.cuh"
namespace c3b {
template __global__ void pwc_t18_cta_kernel<1, 2, 32, 7>(const GemmParams);
template __global__ void pwc_t18_cta_kernel<1, 2, 32, 8>(const GemmParams);
template __global__ void pwc_t18_cta_kernel<1, 1, 32, 7, 512>(const GemmParams);
template __global__ void pwc_t18_cta_kernel<1, 2, 0>(const GemmParams);
template __global__ void pwc_t18_cta_kernel<2, 2, 0>(const GemmParams);
template __global__ void pwc_t18_cta_kernel<1, 1, 0>(const GemmParams);
}
