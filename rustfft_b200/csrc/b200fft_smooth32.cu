// libb200fft.so -- translation unit 4 of 5: the run-time-radix Complex<f32> kernels (Smooth, SmoothFourStep) + their planner.
#include "rt_cuda.h"
#define B2_PART_SMOOTH32 1
#include "impl.inl"
