// libb200fft.so -- translation unit 3 of 5: every Complex<f64> kernel instantiation + its planner.
#include "rt_cuda.h"
#define B2_PART_F64 1
#include "impl.inl"
