// libb200fft.so -- translation unit 5 of 5: the run-time-radix Complex<f64> kernels (Smooth, SmoothFourStep) + their planner.
#include "rt_cuda.h"
#define B2_PART_SMOOTH64 1
#include "impl.inl"
