// Memory-system ceilings of one B200, measured with plain coalesced 16-byte accesses (no FFT): what the SM <-> L2 fabric, the L2
// slices and HBM sustain for the traffic mixes the FFT plans generate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o membw tools/membw.cu ; run under gpurun.  Output: one line per experiment, GB/s of *kernel-visible* bytes (loads + stores).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                              \
        }                                                                              \
    } while (0)

__device__ __forceinline__ float4 ldcs(const float4* p) { return __ldcs(p); }

// each CTA streams over chunks of CH float4 (CH * 16 bytes contiguous); U loads in flight per thread
template <int U>
__global__ void k_read(const float4* __restrict__ a, size_t n4, float* sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i + (U - 1) * blockDim.x < n4; i += stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldcs(a + i + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 1.2345f) *sink = acc;
}
template <int U>
__global__ void k_write(float4* __restrict__ a, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i + (U - 1) * blockDim.x < n4; i += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(a + i + (size_t)u * blockDim.x, v);
    }
}
// copy src (ns4 float4, wraps) -> dst (nd4 float4, wraps): total n4 float4 moved
template <int U>
__global__ void k_copy(const float4* __restrict__ src, size_t ns4, float4* __restrict__ dst, size_t nd4, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i + (U - 1) * blockDim.x < n4; i += stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = src[(i + (size_t)u * blockDim.x) % ns4];
#pragma unroll
        for (int u = 0; u < U; ++u) dst[(i + (size_t)u * blockDim.x) % nd4] = v[u];
    }
}
// the traffic of a two-pass FFT plan without the FFT: big -> ring (L2), ring -> big2; "pass A" and "pass B" chunks interleaved
// at 64 KiB granularity inside every CTA, the ring read lags the ring write by `lag` chunks
template <int U>
__global__ void k_twopass(const float4* __restrict__ big, float4* __restrict__ out, float4* __restrict__ ring, size_t ring4, size_t n4,
                          size_t lag4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i + (U - 1) * blockDim.x < n4; i += stride) {
        float4 v[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldcs(big + i + (size_t)u * blockDim.x);
        const size_t j = i >= lag4 ? i - lag4 : i;
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = __ldcg(ring + (j + (size_t)u * blockDim.x) % ring4);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcg(ring + (i + (size_t)u * blockDim.x) % ring4, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(out + i + (size_t)u * blockDim.x, w[u]);
    }
}

template <class F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    const size_t BIG = 4ull << 30, SMALL = 32ull << 20;
    float4 *a, *b, *s1, *s2;
    float* sink;
    CK(cudaMalloc(&a, BIG));
    CK(cudaMalloc(&b, BIG));
    CK(cudaMalloc(&s1, SMALL));
    CK(cudaMalloc(&s2, SMALL));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, BIG));
    CK(cudaMemset(b, 1, BIG));
    CK(cudaMemset(s1, 1, SMALL));
    CK(cudaMemset(s2, 1, SMALL));
    const size_t big4 = BIG / 16, small4 = SMALL / 16;
    const int NT = 512;
    for (int per_sm : {2, 4}) {
        const int grid = 148 * per_sm;
        float ms;
        ms = time_ms([&] { k_read<8><<<grid, NT>>>(a, big4, sink); });
        std::printf("grid %4d  HBM read-only      %8.1f GB/s\n", grid, BIG / ms / 1e6);
        ms = time_ms([&] { k_write<8><<<grid, NT>>>(a, big4); });
        std::printf("grid %4d  HBM write-only     %8.1f GB/s\n", grid, BIG / ms / 1e6);
        ms = time_ms([&] { k_copy<8><<<grid, NT>>>(a, big4, b, big4, big4); });
        std::printf("grid %4d  HBM copy (rd+wr)   %8.1f GB/s\n", grid, 2.0 * BIG / ms / 1e6);
        // L2 resident: loop the small buffer 128 times (4 GiB of traffic each way)
        ms = time_ms([&] { k_copy<8><<<grid, NT>>>(s1, small4, b, big4, 0); });  // (warm-up no-op)
        ms = time_ms([&] {
            for (int r = 0; r < 32; ++r) k_read<8><<<grid, NT>>>(s1, small4, sink);
        });
        std::printf("grid %4d  L2 read-only       %8.1f GB/s\n", grid, 32.0 * SMALL / ms / 1e6);
        ms = time_ms([&] {
            for (int r = 0; r < 32; ++r) k_write<8><<<grid, NT>>>(s1, small4);
        });
        std::printf("grid %4d  L2 write-only      %8.1f GB/s\n", grid, 32.0 * SMALL / ms / 1e6);
        ms = time_ms([&] { k_copy<8><<<grid, NT>>>(s1, small4, s2, small4, big4); });
        std::printf("grid %4d  L2 copy (rd+wr)    %8.1f GB/s\n", grid, 2.0 * BIG / ms / 1e6);
        ms = time_ms([&] { k_copy<8><<<grid, NT>>>(a, big4, s2, small4, big4); });
        std::printf("grid %4d  HBM rd -> L2 wr    %8.1f GB/s\n", grid, 2.0 * BIG / ms / 1e6);
        ms = time_ms([&] { k_copy<8><<<grid, NT>>>(s1, small4, b, big4, big4); });
        std::printf("grid %4d  L2 rd -> HBM wr    %8.1f GB/s\n", grid, 2.0 * BIG / ms / 1e6);
        ms = time_ms([&] { k_twopass<4><<<grid, NT>>>(a, b, s1, small4, big4, small4 / 2); });
        std::printf("grid %4d  two-pass mix       %8.1f GB/s kernel-visible (= %.1f GB/s algorithmic, HBM rd+wr only)\n", grid, 4.0 * BIG / ms / 1e6,
                    2.0 * BIG / ms / 1e6);
    }
    return 0;
}
