// compile-only harness for the signal chain kernels
#include "../c3_b200/csrc/signal_chain.cuh"
