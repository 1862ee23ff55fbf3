// b200fft -- common definitions shared by the sm_100a kernels and the host planner.
//
// Every kernel body in this library is written as a sequence of `phase<P>()` functions that are
// separated by a CTA-wide barrier.  On the GPU `run_kernel<K>` calls them back to back with
// __syncthreads() in between; the test-only CPU harness (tests/emu) replays the same phases
// thread by thread so that index maths, twiddle tables and planning can be checked without a GPU.
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__
#else
#define B2_HD inline
#define B2_D inline
#endif

#if defined(__CUDA_ARCH__)
#define B2_UNROLL _Pragma("unroll")
#else
#define B2_UNROLL
#endif

namespace b2 {

// Complex<T> of the reference is repr(C) {re, im} (CHANGELOG.md:139) == float2 / double2.
template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};

template <typename T> B2_HD cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename T> B2_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> B2_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
// (a.x + i a.y)(w.x + i w.y); compiles to 2 mul + 2 fma
template <typename T> B2_HD cx<T> cmul(cx<T> a, cx<T> w) {
    return mk<T>(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
// a * conj(w)
template <typename T> B2_HD cx<T> cmulc(cx<T> a, cx<T> w) {
    return mk<T>(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}
template <typename T> B2_HD cx<T> conj(cx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by -i  (forward quarter turn, twiddle(1,4))
template <typename T> B2_HD cx<T> mul_mi(cx<T> a) { return mk<T>(a.y, -a.x); }
// swap re <-> im.  ifft(x) = swap(fft(swap(x))): the whole inverse direction is a register
// renaming at the outermost load and store of a plan, every table stays "forward".
template <typename T> B2_HD cx<T> swap_ri(cx<T> a) { return mk<T>(a.y, a.x); }

// read-only global load through the non-coherent path on device (tables shared by all CTAs)
template <typename T> B2_HD cx<T> ldg(const cx<T>* p) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        float2 v = __ldg(reinterpret_cast<const float2*>(p));
        return mk<T>(v.x, v.y);
    } else {
        double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return mk<T>(v.x, v.y);
    }
#else
    return *p;
#endif
}
B2_HD uint32_t ldg_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// streaming global load / store of signal data: read once, written once -> keep it out of L1
template <typename T> B2_HD cx<T> ld_stream(const cx<T>* p) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        float2 v = __ldcs(reinterpret_cast<const float2*>(p));
        return mk<T>(v.x, v.y);
    } else {
        double2 v = __ldcs(reinterpret_cast<const double2*>(p));
        return mk<T>(v.x, v.y);
    }
#else
    return *p;
#endif
}
template <typename T> B2_HD void st_stream(cx<T>* p, cx<T> v) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
    } else {
        __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
    }
#else
    *p = v;
#endif
}

// compile-time list of stage radices
template <int... Rs>
struct Radices {
    static constexpr int N = sizeof...(Rs);
    static constexpr int get(int i) {
        constexpr int a[] = {Rs...};
        return a[i];
    }
    static constexpr int product(int upto = N) {  // product of the first `upto` radices
        constexpr int a[] = {Rs...};
        int p = 1;
        for (int i = 0; i < upto; ++i) p *= a[i];
        return p;
    }
    // offset (in elements) of stage s inside the packed stage-twiddle table; stage 0 has none
    static constexpr int tw_offset(int s) {
        constexpr int a[] = {Rs...};
        int off = 0, p = 1;
        for (int i = 0; i < s; ++i) {
            if (i >= 1) off += (a[i] - 1) * p;
            p *= a[i];
        }
        return off;
    }
    static constexpr int tw_total() { return tw_offset(N); }
};

template <int A, int B> struct StaticMax { static constexpr int v = A > B ? A : B; };

}  // namespace b2
