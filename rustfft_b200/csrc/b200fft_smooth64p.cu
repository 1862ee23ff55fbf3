// libb200fft.so: f64 run-time-radix two-pass kernels with the prime butterflies
#include "rt_cuda.h"
#define B2_PART_SMOOTH64P 1
#include "impl.inl"
