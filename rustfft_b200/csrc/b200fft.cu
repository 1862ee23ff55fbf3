// libb200fft.so -- translation unit 1 of 5: the C ABI (include/b200fft.h), plan object, host-slice pipeline.
// The kernel instantiations live in b200fft_f32.cu / b200fft_f64.cu so the three compile in parallel.
// Build: rustfft_b200/csrc/Makefile   (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 ...)
#include "rt_cuda.h"
#define B2_PART_CABI 1
#include "impl.inl"
