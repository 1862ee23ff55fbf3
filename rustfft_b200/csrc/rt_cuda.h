// `rt::` runtime layer on the CUDA runtime (streams, events, memory, kernel launches) -- included by every
// CUDA translation unit of libb200fft.so.  tests/emu/b200fft_emu.cpp provides the same names on the CPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <string>

#include "common.h"

namespace b2 {
namespace rt {

typedef cudaStream_t stream_t;

inline thread_local std::string g_err;  // (inline: ONE instance for all translation units of the library)
static bool check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return false;
}
static std::string last_error() { return g_err; }

// devices this library can run on: compute capability 10.x (the cubin is sm_100a only)
static int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++usable;
        else break;  // keep device indices dense
    }
    return usable;
}
static bool set_device(int d) { return check(cudaSetDevice(d), "cudaSetDevice"); }
static void* dmalloc(size_t bytes) {
    void* p = nullptr;
    if (!check(cudaMalloc(&p, bytes), "cudaMalloc")) return nullptr;
    return p;
}
static void dfree(void* p) { cudaFree(p); }
static bool h2d_sync(void* d, const void* h, size_t n) { return check(cudaMemcpy(d, h, n, cudaMemcpyHostToDevice), "cudaMemcpy H2D"); }
static bool h2d_async(void* d, const void* h, size_t n, stream_t s) {
    return check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
}
static bool d2h_async(void* h, const void* d, size_t n, stream_t s) {
    return check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
}
static bool d2d_async(void* dst, const void* src, size_t n, stream_t s) {
    return check(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync D2D");
}
static bool memset_async(void* d, int v, size_t n, stream_t s) { return check(cudaMemsetAsync(d, v, n, s), "cudaMemsetAsync"); }
// The default pool returns freed memory to the OS at the next synchronisation point, so a caller that synchronises
// between two execs paid a fresh multi-millisecond allocation for the same workspace every time (the 0.68 <-> 3.9 ms
// bimodal timing of the 65537-point plan in round 1).  The pool keeps what it has been given: one threshold per device.
static void keep_pool_memory() {
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    const uint64_t bit = 1ull << (dev & 63);
    if (configured.load(std::memory_order_acquire) & bit) return;
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    configured.fetch_or(bit, std::memory_order_release);
}
static void* malloc_async(size_t bytes, stream_t s) {
    keep_pool_memory();
    void* p = nullptr;
    if (!check(cudaMallocAsync(&p, bytes, s), "cudaMallocAsync")) return nullptr;
    return p;
}
static void free_async(void* p, stream_t s) { cudaFreeAsync(p, s); }
static stream_t stream_create() {
    cudaStream_t s = nullptr;
    if (!check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate")) return nullptr;
    return s;
}
static void stream_destroy(stream_t s) { cudaStreamDestroy(s); }
static bool stream_sync(stream_t s) { return check(cudaStreamSynchronize(s), "cudaStreamSynchronize"); }
typedef cudaEvent_t event_t;
static event_t event_create() {
    cudaEvent_t e = nullptr;
    if (!check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    return e;
}
static void event_destroy(event_t e) { cudaEventDestroy(e); }
static bool event_record(event_t e, stream_t s) { return check(cudaEventRecord(e, s), "cudaEventRecord"); }
static bool stream_wait(stream_t s, event_t e) { return check(cudaStreamWaitEvent(s, e, 0), "cudaStreamWaitEvent"); }
static bool event_sync(event_t e) { return check(cudaEventSynchronize(e), "cudaEventSynchronize"); }
// page-locked host memory for the staging rings of the host-slice path
static void* host_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (!check(cudaHostAlloc(&p, bytes, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
    return p;
}
static void host_free_pinned(void* p) { cudaFreeHost(p); }
// true when the device can DMA straight from / to this host pointer (pinned or registered); pageable memory is staged
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
// RAII: the library never leaves the calling thread on another device than it found it on
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = check(cudaSetDevice(dev), "cudaSetDevice");
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- L2 persistence for the L2-resident intermediate of two-pass plans ---------------------------------------
// B200FFT_L2_PERSIST_MB = N (N > 0): reserve N MiB of L2 for persisting lines (cudaLimitPersistingL2CacheSize, a
// DEVICE-WIDE setting, hence opt-in) and launch the passes with an access-policy window over their workspace, so the
// intermediate is protected from eviction by the streamed input/output.  Measured without it (profiles/r1x): 42-92 %
// of the intermediate is written back to HBM although the second pass still finds it in L2.
struct L2Window {
    void* base;
    size_t bytes;
};
static thread_local L2Window g_l2win{nullptr, 0};
static size_t l2_persist_bytes() {
    static size_t v = [] {
        const char* e = std::getenv("B200FFT_L2_PERSIST_MB");
        const long mb = e ? std::atol(e) : 0;
        return mb > 0 ? (size_t)mb << 20 : (size_t)0;
    }();
    return v;
}
// called by a two-pass exec around its launches; returns false (and changes nothing) when persistence is off
static bool set_l2_window(void* base, size_t bytes) {
    if (!base || !bytes) {
        g_l2win = L2Window{nullptr, 0};
        return false;
    }
    if (!l2_persist_bytes()) return false;
    static std::atomic<uint64_t> configured{0};
    static std::atomic<size_t> max_window{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return false;
        const size_t want = std::min<size_t>(l2_persist_bytes(), (size_t)prop.persistingL2CacheMaxSize);
        if (want == 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        max_window.store((size_t)prop.accessPolicyMaxWindowSize, std::memory_order_relaxed);
        configured.fetch_or(bit, std::memory_order_release);
    }
    g_l2win = L2Window{base, std::min(bytes, max_window.load(std::memory_order_relaxed))};
    return true;
}
template <class P>
static bool launch_ex(void (*kern)(P), unsigned grid, unsigned block, size_t smem, stream_t s, const P& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    unsigned na = 0;
    if (g_l2win.bytes) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = g_l2win.base;
        attr[0].val.accessPolicyWindow.num_bytes = g_l2win.bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return check(cudaLaunchKernelEx(&cfg, kern, p), "kernel launch");
}

}  // namespace rt
}  // namespace b2

#include "fused.h"
#include "cluster.h"

namespace b2 {
namespace rt {

// thread-block cluster launch (cluster.h): grid = clusters x C CTAs, cluster dimension C (> 8 needs the non-portable opt-in)
template <class KT>
static int cluster_max_active() {
    static std::atomic<int> cached[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int v = cached[dev & 63].load(std::memory_order_acquire);
    if (v != 0) return v;
    v = -1;
    bool ok = check(cudaFuncSetAttribute(run_cluster<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT::SMEM_BYTES),
                    "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    if (ok && KT::C > 8)
        ok = check(cudaFuncSetAttribute(run_cluster<KT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                   "cudaFuncSetAttribute(NonPortableClusterSizeAllowed)");
    if (ok) {
        cudaFuncSetAttribute(run_cluster<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaGetLastError();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(KT::C, 1, 1);
        cfg.blockDim = dim3(KT::NT, 1, 1);
        cfg.dynamicSmemBytes = KT::SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = KT::C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (check(cudaOccupancyMaxActiveClusters(&n, run_cluster<KT>, &cfg), "cudaOccupancyMaxActiveClusters") && n > 0) v = n;
        else if (n <= 0) g_err = "no cluster of this size fits on the device";
    }
    cached[dev & 63].store(v, std::memory_order_release);
    return v;
}
template <class KT>
static bool launch_cluster(const typename KT::Params& p, uint64_t clusters, stream_t s) {
    if (clusters == 0) return true;
    if (clusters * KT::C > 0x7fffffffull) {
        g_err = "grid too large";
        return false;
    }
    if (cluster_max_active<KT>() <= 0) {
        if (g_err.empty()) g_err = "thread-block clusters of this size cannot be launched on this device";
        return false;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * KT::C), 1, 1);
    cfg.blockDim = dim3(KT::NT, 1, 1);
    cfg.dynamicSmemBytes = KT::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = KT::C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return check(cudaLaunchKernelEx(&cfg, run_cluster<KT>, p), "cluster kernel launch");
}

// once per (kernel, device): opt in to > 48 KiB dynamic shared memory, and ask for a shared-memory
// carveout that fits as many CTAs as registers and threads allow (the driver's default carveout left
// the 70 KiB tile kernels at 1 CTA/SM in the first round-1 capture)
template <class KT>
static bool ensure_configured() {
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (configured.load(std::memory_order_acquire) & bit) return true;
    if (KT::SMEM_BYTES > 48 * 1024 &&
        !check(cudaFuncSetAttribute(run_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT::SMEM_BYTES),
               "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
        return false;
    if (KT::SMEM_BYTES > 0) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, run_kernel<KT>) == cudaSuccess) {
            const int regs = ((fa.numRegs + 7) / 8) * 8;
            int want = 65536 / (regs * KT::NT);
            if (want > 2048 / KT::NT) want = 2048 / KT::NT;
            if (want < 1) want = 1;
            const size_t need = (size_t)want * (KT::SMEM_BYTES + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            cudaFuncSetAttribute(run_kernel<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        cudaGetLastError();
    }
    configured.fetch_or(bit, std::memory_order_release);
    return true;
}

// CTAs of this kernel the whole device holds at once (one "wave"); the planner sizes the L2 chunks of
// multi-pass plans so that every launch is close to a whole number of waves
template <class KT>
static int resident_ctas() {
    if (!ensure_configured<KT>()) return 0;
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, run_kernel<KT>, KT::NT, KT::SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
}

template <class KT>
static bool launch(const typename KT::Params& p, uint64_t ctas, stream_t s) {
    if (ctas == 0) return true;
    if (ctas > 0x7fffffffull) {
        g_err = "grid too large";
        return false;
    }
    if (!ensure_configured<KT>()) return false;
    return launch_ex(run_kernel<KT>, (unsigned)ctas, KT::NT, KT::SMEM_BYTES, s, p);
}

// persistent variant: grid = resident CTAs of the kernel (queried once per kernel and device)
template <class KT>
static bool launch_persistent(const typename KT::Params& p, uint64_t ctas, stream_t s) {
    if (ctas == 0) return true;
    if (ctas > 0x7fffffffull) {
        g_err = "grid too large";
        return false;
    }
    static std::atomic<int> grid_for_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int grid = grid_for_dev[dev & 63].load(std::memory_order_acquire);
    if (grid == 0) {
        if (KT::SMEM_BYTES > 48 * 1024 &&
            !check(cudaFuncSetAttribute(run_kernel_persistent<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT::SMEM_BYTES),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return false;
        cudaFuncSetAttribute(run_kernel_persistent<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        int per_sm = 0, sms = 0;
        if (!check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, run_kernel_persistent<KT>, KT::NT, KT::SMEM_BYTES),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
            return false;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1 || sms < 1) {
            g_err = "persistent kernel does not fit on an SM";
            return false;
        }
        grid = per_sm * sms;
        grid_for_dev[dev & 63].store(grid, std::memory_order_release);
    }
    const unsigned g = (unsigned)std::min<uint64_t>((uint64_t)grid, ctas);
    run_kernel_persistent<KT><<<g, KT::NT, KT::SMEM_BYTES, s>>>(p, (uint32_t)ctas);
    return check(cudaGetLastError(), "kernel launch");
}

// kernels with run-time sized dynamic shared memory (<= max_smem bytes, configured once)
template <class KT>
static bool launch_dyn(const typename KT::Params& p, uint64_t ctas, size_t smem_bytes, size_t max_smem, stream_t s) {
    if (ctas == 0) return true;
    if (ctas > 0x7fffffffull) {
        g_err = "grid too large";
        return false;
    }
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        if (!check(cudaFuncSetAttribute(run_kernel_dyn<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return false;
        cudaFuncSetAttribute(run_kernel_dyn<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured.fetch_or(bit, std::memory_order_release);
    }
    run_kernel_dyn<KT><<<(unsigned)ctas, KT::NT, smem_bytes, s>>>(p);
    return check(cudaGetLastError(), "kernel launch");
}

// kernels whose step count is run-time data (run_kernel_loop), run-time sized dynamic shared memory
template <class KT>
static bool launch_loop(const typename KT::Params& p, uint64_t ctas, uint32_t n_steps, size_t smem_bytes, size_t max_smem, stream_t s) {
    if (ctas == 0) return true;
    if (ctas > 0x7fffffffull) {
        g_err = "grid too large";
        return false;
    }
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        if (!check(cudaFuncSetAttribute(run_kernel_loop<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return false;
        cudaFuncSetAttribute(run_kernel_loop<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured.fetch_or(bit, std::memory_order_release);
    }
    run_kernel_loop<KT><<<(unsigned)ctas, KT::NT, smem_bytes, s>>>(p, n_steps);
    return check(cudaGetLastError(), "kernel launch");
}

// persistent pipelined kernels: grid = SMs x resident CTAs (queried once per kernel and device)
template <class KT>
static bool launch_pipelined(const typename KT::Params& p, stream_t s) {
    if (p.n_items == 0) return true;
    static std::atomic<int> grid_for_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int grid = grid_for_dev[dev & 63].load(std::memory_order_acquire);
    if (grid == 0) {
        if (!check(cudaFuncSetAttribute(run_pipelined<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT::SMEM_BYTES),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return false;
        cudaFuncSetAttribute(run_pipelined<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        int per_sm = 0, sms = 0;
        if (!check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, run_pipelined<KT>, KT::NT, KT::SMEM_BYTES),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
            return false;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1 || sms < 1) {
            g_err = "pipelined kernel does not fit on an SM";
            return false;
        }
        grid = per_sm * sms;
        grid_for_dev[dev & 63].store(grid, std::memory_order_release);
    }
    const unsigned g = (unsigned)((uint64_t)grid < (uint64_t)p.n_items ? grid : (int)p.n_items);
    run_pipelined<KT><<<g, KT::NT, KT::SMEM_BYTES, s>>>(p);
    return check(cudaGetLastError(), "kernel launch");
}

// ---- TMA tensor maps -----------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).  A map describes
// `slabs` matrices of `rows` x `inner_cx` complex numbers (row-major, dense) and a box of box_rows x box_cx.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encode_tiled() {
    static encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (encode_tiled_fn) nullptr;
        }
        return (encode_tiled_fn)p;
    }();
    return fn;
}
static bool tma_available() { return encode_tiled() != nullptr; }
static bool make_tile_map(TMap* out, bool f64, const void* base, uint64_t inner_cx, uint64_t rows, uint64_t slabs, uint32_t box_cx,
                          uint32_t box_rows) {
    static_assert(sizeof(TMap) == sizeof(CUtensorMap), "TMap must mirror CUtensorMap");
    encode_tiled_fn fn = encode_tiled();
    if (!fn) {
        g_err = "cuTensorMapEncodeTiled is not available";
        return false;
    }
    const uint64_t esz = f64 ? 8 : 4;
    const cuuint64_t dims[3] = {2 * inner_cx, rows, slabs};
    const cuuint64_t strides[2] = {2 * inner_cx * esz, 2 * inner_cx * rows * esz};
    const cuuint32_t box[3] = {2 * box_cx, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    // L2 promotion of the boxes' sectors: none by default (a 64-byte box row must not drag in its 128-byte line's other
    // half); B200FFT_TMA_L2PROMO=1/2/3 selects 64 / 128 / 256 bytes for A/B measurements
    static const CUtensorMapL2promotion promo = [] {
        const char* e = std::getenv("B200FFT_TMA_L2PROMO");
        const int k = e ? std::atoi(e) : 0;
        return k == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : k == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
             : k == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    }();
    const CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                          const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return false;
    }
    return true;
}
// the tile-major ring of the fused plans as a 4-D tensor [slot][pass-A tile][row k1][fo columns]: box = all (<= 256) tiles x
// box_rows rows x fo columns -- the rows of one pass-B tile, gathered from every pass-A tile of the transform
static bool make_ring_map(TMap* out, const void* base, uint32_t fo, uint64_t rows, uint64_t tiles, uint64_t slots, uint32_t box_rows,
                          uint32_t box_tiles) {
    encode_tiled_fn fn = encode_tiled();
    if (!fn) {
        g_err = "cuTensorMapEncodeTiled is not available";
        return false;
    }
    const cuuint64_t dims[4] = {2ull * fo, rows, tiles, slots};
    const cuuint64_t strides[3] = {2ull * fo * 4, 2ull * fo * 4 * rows, 2ull * fo * 4 * rows * tiles};
    const cuuint32_t box[4] = {2 * fo, box_rows, box_tiles, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_err = "cuTensorMapEncodeTiled (ring) failed with CUresult " + std::to_string((int)r);
        return false;
    }
    return true;
}
template <class KT>
static bool launch_tma(const typename KT::Params& p, uint64_t ctas, stream_t s) {
    if (ctas == 0) return true;
    if (ctas > 0x7fffffffull) {
        g_err = "grid too large";
        return false;
    }
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        if (KT::SMEM_BYTES + 144 > 48 * 1024 &&
            !check(cudaFuncSetAttribute(run_kernel_tma<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KT::SMEM_BYTES + 144)),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return false;
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, run_kernel_tma<KT>) == cudaSuccess) {
            const int regs = ((fa.numRegs + 7) / 8) * 8;
            int want = 65536 / (regs * KT::NT);
            if (want > 2048 / KT::NT) want = 2048 / KT::NT;
            if (want < 1) want = 1;
            const size_t need = (size_t)want * (KT::SMEM_BYTES + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            cudaFuncSetAttribute(run_kernel_tma<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        cudaGetLastError();
        configured.fetch_or(bit, std::memory_order_release);
    }
    return launch_ex(run_kernel_tma<KT>, (unsigned)ctas, KT::NT, KT::SMEM_BYTES + 144, s, p);
}

// single-launch dataflow four-step: persistent grid = resident CTAs (queried once per kernel and device);
// the control block is zeroed on the stream right before the launch (both are graph-capturable)
template <class KA, class KB>
static int flow_grid() {
    using FK = FlowKernel<KA, KB>;
    static std::atomic<int> grid_for_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int grid = grid_for_dev[dev & 63].load(std::memory_order_acquire);
    if (grid == 0) {
        if (FK::SMEM_BYTES > 48 * 1024 &&
            !check(cudaFuncSetAttribute(run_flow<KA, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FK::SMEM_BYTES),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return 0;
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, run_flow<KA, KB>) == cudaSuccess) {
            const int regs = ((fa.numRegs + 7) / 8) * 8;
            int want = 65536 / (regs * FK::NT);
            if (want > 2048 / FK::NT) want = 2048 / FK::NT;
            if (want < 1) want = 1;
            const size_t need = (size_t)want * (FK::SMEM_BYTES + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            cudaFuncSetAttribute(run_flow<KA, KB>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        cudaGetLastError();
        int per_sm = 0, sms = 0;
        if (!check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, run_flow<KA, KB>, FK::NT, FK::SMEM_BYTES),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
            return 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1 || sms < 1) {
            g_err = "dataflow kernel does not fit on an SM";
            return 0;
        }
        grid = per_sm * sms;
        grid_for_dev[dev & 63].store(grid, std::memory_order_release);
    }
    return grid;
}
template <class KA, class KB>
static bool launch_flow(const typename FlowKernel<KA, KB>::Params& p, uint64_t ctl_bytes, stream_t s) {
    using FK = FlowKernel<KA, KB>;
    if (p.sched.total == 0) return true;
    const int grid = flow_grid<KA, KB>();
    if (grid <= 0) return false;
    if (!memset_async(p.ctl, 0, ctl_bytes, s)) return false;
    const unsigned g = (unsigned)std::min<uint64_t>((uint64_t)grid, (uint64_t)p.sched.total);
    return launch_ex(run_flow<KA, KB>, g, FK::NT, FK::SMEM_BYTES, s, p);
}

// fused single-launch four-step (fused.h): persistent grid = one CTA per resident slot (queried once per kernel and
// device); the control block is zeroed on the stream right before the launch (both are graph-capturable)
template <class KA, class KB, int NG, int NS>
static int fused_grid() {
    using FK = FusedKernel<KA, KB, NG, NS>;
    static std::atomic<int> grid_for_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int grid = grid_for_dev[dev & 63].load(std::memory_order_acquire);
    if (grid == 0) {
        if (!check(cudaFuncSetAttribute(run_fused<KA, KB, NG, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FK::SMEM_BYTES),
                   "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
            return 0;
        cudaFuncSetAttribute(run_fused<KA, KB, NG, NS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaGetLastError();
        int per_sm = 0, sms = 0;
        if (!check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, run_fused<KA, KB, NG, NS>, FK::NT, FK::SMEM_BYTES),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
            return 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1 || sms < 1) {
            g_err = "fused kernel does not fit on an SM";
            return 0;
        }
        grid = per_sm * sms;
        grid_for_dev[dev & 63].store(grid, std::memory_order_release);
    }
    return grid;
}
template <class KA, class KB, int NG, int NS>
static bool launch_fused(const typename FusedKernel<KA, KB, NG, NS>::Params& p, uint64_t ctl_bytes, stream_t s) {
    using FK = FusedKernel<KA, KB, NG, NS>;
    if (p.sched.total == 0) return true;
    const int grid = fused_grid<KA, KB, NG, NS>();
    if (grid <= 0) return false;
    if (!memset_async(p.ctl, 0, ctl_bytes, s)) return false;
    const unsigned g = (unsigned)std::min<uint64_t>((uint64_t)grid, (uint64_t)p.sched.total);
    return launch_ex(run_fused<KA, KB, NG, NS>, g, FK::NT, FK::SMEM_BYTES, s, p);
}

}  // namespace rt
}  // namespace b2

