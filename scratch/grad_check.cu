// compile-only harness for the gradient kernels
#include "../c3_b200/csrc/grad.cuh"
