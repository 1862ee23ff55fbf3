// libb200fft.so, translation unit 7 of 7: compiled composite tiles of the f32 two-pass plans (SmoothTileGeo, impl.inl)
#include "rt_cuda.h"
#define B2_PART_CTILE32 1
#include "impl.inl"
