// libb200fft.so -- translation unit 2 of 5: every Complex<f32> kernel instantiation + its planner.
#include "rt_cuda.h"
#define B2_PART_F32 1
#include "impl.inl"
