// libb200fft.so: f32 run-time-radix kernels without the prime butterflies (3 CTAs per SM)
#include "rt_cuda.h"
#define B2_PART_SMOOTH32S 1
#include "impl.inl"
