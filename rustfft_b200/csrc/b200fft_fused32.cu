// libb200fft.so, translation unit 6 of 6: the fused single-launch four-step kernels (f32), see fused.h
#include "rt_cuda.h"
#define B2_PART_FUSED32 1
#include "impl.inl"
