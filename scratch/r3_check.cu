// compile-only harness (nvcc -I c3_b200/csrc): register/spill report of the R-rows-per-lane Taylor kernel variants
#include "pwc_r3t18.cuh"
namespace c3b {
template __global__ void pwc_r3t18_kernel<9, 3, 4>(const RowsParams, unsigned int*);
template __global__ void pwc_r3t18_kernel<9, 2, 7>(const RowsParams, unsigned int*);
template __global__ void pwc_r3t18_kernel<9, 2, 6>(const RowsParams, unsigned int*);
}
