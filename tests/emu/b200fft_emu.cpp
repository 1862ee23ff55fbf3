// TEST INFRASTRUCTURE -- NOT PRODUCT CODE, never linked into libb200fft.so.
//
// Replays the library's kernel *phase functions* (rustfft_b200/csrc/kernels.h) on the CPU, one
// CTA at a time, thread by thread, with a barrier between phases -- exactly the structure
// run_kernel<K> has on the GPU.  It lets the CPU test-suite (no GPU in the build container) check the
// planner, the twiddle / chirp / index tables and every kernel's index arithmetic through the same
// C ABI (include/b200fft.h) before GPU time is spent.  "Device" memory is host memory here.
// What it cannot see: real races, shared-memory limits, launch configuration -- those are what the
// -m gpu tests are for.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../rustfft_b200/csrc/common.h"

namespace b2 {
namespace rt {
typedef void* stream_t;
static std::string g_err;
static std::string last_error() { return g_err; }
static int device_count() { return 1; }
static bool set_device(int) { return true; }
// (256-byte aligned like cudaMalloc: the library rejects workspaces that are not 128-byte aligned)
static void* dmalloc(size_t n) { return std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); }
static void dfree(void* p) { std::free(p); }
static bool h2d_sync(void* d, const void* h, size_t n) { std::memcpy(d, h, n); return true; }
static bool h2d_async(void* d, const void* h, size_t n, stream_t) { std::memcpy(d, h, n); return true; }
static bool d2h_async(void* h, const void* d, size_t n, stream_t) { std::memcpy(h, d, n); return true; }
static bool d2d_async(void* dst, const void* src, size_t n, stream_t) { std::memmove(dst, src, n); return true; }
static bool set_l2_window(void*, size_t) { return false; }
static bool memset_async(void* d, int v, size_t n, stream_t) { std::memset(d, v, n); return true; }
static void* malloc_async(size_t n, stream_t) { return dmalloc(n); }
static void free_async(void* p, stream_t) { std::free(p); }
static stream_t stream_create() { return (stream_t)1; }
static void stream_destroy(stream_t) {}
static bool stream_sync(stream_t) { return true; }
typedef void* event_t;
static event_t event_create() { return (event_t)1; }
static void event_destroy(event_t) {}
static bool event_record(event_t, stream_t) { return true; }
static bool stream_wait(stream_t, event_t) { return true; }
static bool event_sync(event_t) { return true; }
static void* host_alloc_pinned(size_t n) { return std::malloc(n ? n : 1); }
static void host_free_pinned(void* p) { std::free(p); }
static bool host_is_pinned(const void*) { return false; }  // exercise the staged path on the CPU
struct DeviceGuard {
    bool ok = true;
    explicit DeviceGuard(int) {}
};
}  // namespace rt
}  // namespace b2

#include "../../rustfft_b200/csrc/fused.h"
#include "../../rustfft_b200/csrc/cluster.h"

namespace b2 {
namespace rt {

static uint64_t g_launches = 0;

template <class KT, int P>
struct EmuPhases {
    template <class S>
    static void run(const typename KT::Params& p, uint32_t bid, std::vector<typename KT::Regs>& regs, S* smem) {
        for (int tid = 0; tid < KT::NT; ++tid) KT::template phase<P>(p, bid, tid, regs[(size_t)tid], smem);
        if constexpr (P + 1 < KT::NPHASE) EmuPhases<KT, P + 1>::run(p, bid, regs, smem);
    }
};

template <class KT>
static bool launch(const typename KT::Params& p, uint64_t ctas, stream_t) {
    ++g_launches;
    std::vector<typename KT::Regs> regs((size_t)KT::NT);
    // poison shared memory so a read of a slot nobody wrote shows up as NaN
    std::vector<cx<typename KT::T>> smem(KT::SMEM_BYTES / sizeof(cx<typename KT::T>) + 1);
    for (uint64_t bid = 0; bid < ctas; ++bid) {
        std::memset(smem.data(), 0xff, smem.size() * sizeof(smem[0]));
        EmuPhases<KT, 0>::run(p, (uint32_t)bid, regs, smem.data());
    }
    return true;
}

template <class KT>
static bool launch_persistent(const typename KT::Params& p, uint64_t ctas, stream_t s) { return launch<KT>(p, ctas, s); }

template <class KT>
static bool launch_dyn(const typename KT::Params& p, uint64_t ctas, size_t smem_bytes, size_t, stream_t) {
    ++g_launches;
    std::vector<typename KT::Regs> regs((size_t)KT::NT);
    std::vector<cx<typename KT::T_>> smem(smem_bytes / sizeof(cx<typename KT::T_>) + 1);
    for (uint64_t bid = 0; bid < ctas; ++bid) {
        std::memset(smem.data(), 0xff, smem.size() * sizeof(smem[0]));
        EmuPhases<KT, 0>::run(p, (uint32_t)bid, regs, smem.data());
    }
    return true;
}

template <class KT>
static bool launch_loop(const typename KT::Params& p, uint64_t ctas, uint32_t n_steps, size_t smem_bytes, size_t, stream_t) {
    ++g_launches;
    std::vector<cx<typename KT::T_>> smem(smem_bytes / sizeof(cx<typename KT::T_>) + 1);
    for (uint64_t bid = 0; bid < ctas; ++bid) {
        std::memset(smem.data(), 0xff, smem.size() * sizeof(smem[0]));
        for (uint32_t st = 0; st < n_steps; ++st)
            for (int tid = 0; tid < KT::NT; ++tid) KT::step(p, (uint32_t)bid, tid, st, smem.data());
    }
    return true;
}

// thread-block cluster kernels (cluster.h): the C CTAs of a cluster advance phase by phase together (every boundary is treated
// as a cluster barrier); the DSMEM window is the array of the C shared-memory buffers
template <class KT, int P>
struct EmuClusterPhases {
    static void run(const typename KT::Params& p, uint32_t cluster, std::vector<typename KT::Regs>& regs,
                    std::vector<std::vector<cx<typename KT::T>>>& smem, cx<typename KT::T>* const* remote) {
        for (int c = 0; c < KT::C; ++c)
            for (int tid = 0; tid < KT::NT; ++tid)
                KT::template phase<P>(p, cluster * KT::C + c, tid, regs[(size_t)c * KT::NT + tid], smem[(size_t)c].data(), remote);
        if constexpr (P + 1 < KT::NPHASE) EmuClusterPhases<KT, P + 1>::run(p, cluster, regs, smem, remote);
    }
};
template <class KT>
static int cluster_max_active() { return 32; }
template <class KT>
static bool launch_cluster(const typename KT::Params& p, uint64_t clusters, stream_t) {
    ++g_launches;
    using C = cx<typename KT::T>;
    std::vector<typename KT::Regs> regs((size_t)KT::C * KT::NT);
    std::vector<std::vector<C>> smem((size_t)KT::C, std::vector<C>(KT::SMEM_BYTES / sizeof(C) + 1));
    C* remote[KT::C];
    for (int c = 0; c < KT::C; ++c) remote[c] = smem[(size_t)c].data();
    for (uint64_t cl = 0; cl < clusters; ++cl) {
        for (auto& b : smem) std::memset(b.data(), 0xff, b.size() * sizeof(C));
        EmuClusterPhases<KT, 0>::run(p, (uint32_t)cl, regs, smem, remote);
    }
    return true;
}

template <class KT>
static int resident_ctas() { return 296; }  // what a B200 reports for a 2-CTA/SM kernel; only steers chunk sizes

// persistent pipelined kernels: tiles in order; the TMA bulk copy of a tile becomes a memcpy
template <class KT>
static bool launch_pipelined(const typename KT::Params& p, stream_t) {
    ++g_launches;
    std::vector<typename KT::Regs> regs((size_t)KT::NT);
    std::vector<cx<typename KT::T>> buf(KT::BUF_ELEMS + 1);
    for (uint32_t item = 0; item < p.n_items; ++item) {
        std::memset(buf.data(), 0xff, buf.size() * sizeof(buf[0]));
        std::memcpy(buf.data(), KT::fetch_src(p, item), KT::fetch_bytes(p, item));
        EmuPhases<KT, 0>::run(p, item, regs, buf.data());
    }
    return true;
}

// TMA-tiled passes: the tensor map is never read here -- phase 0 / the last phase of TmaTileKernel copy the box
static bool tma_available() { return true; }
static bool make_tile_map(TMap* out, bool, const void*, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t) {
    std::memset(out, 0, sizeof(*out));
    return true;
}
static bool make_ring_map(TMap* out, const void*, uint32_t, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t) {
    std::memset(out, 0, sizeof(*out));
    return true;
}
template <class KT>
static bool launch_tma(const typename KT::Params& p, uint64_t ctas, stream_t s) { return launch<KT>(p, ctas, s); }

// single-launch dataflow four-step: ONE emulated CTA takes the tickets in order, so every dependency (which always
// points to a smaller ticket) must already be satisfied when its tile starts -- checked here; the counters are
// kept exactly as the device kernel keeps them
template <class KA, class KB>
static int flow_grid() { return 296; }
template <class KA, class KB>
static bool launch_flow(const typename FlowKernel<KA, KB>::Params& p, uint64_t ctl_bytes, stream_t) {
    using FK = FlowKernel<KA, KB>;
    using C = cx<typename FK::T>;
    ++g_launches;
    std::memset(p.ctl, 0, ctl_bytes);
    const FlowSched& sc = p.sched;
    uint32_t* ready = p.ctl + FLOW_CTL_HEAD;
    uint32_t* freed = ready + sc.ring_w;
    std::vector<typename KA::Regs> ra((size_t)KA::NT);
    std::vector<typename KB::Regs> rb((size_t)KB::NT);
    std::vector<C> smem(FK::SMEM_BYTES / sizeof(C) + 1);
    std::vector<uint32_t> seenA((size_t)sc.batch * sc.TA, 0), seenB((size_t)sc.batch * sc.TB, 0);
    for (uint32_t ticket = 0; ticket < sc.total; ++ticket) {
        int kind;
        uint32_t t, tile;
        bool valid;
        sc.decode(ticket, kind, t, tile, valid);
        if (!valid) continue;
        const uint32_t slot = t % sc.ring_w, gen = t / sc.ring_w;
        std::memset(smem.data(), 0xff, smem.size() * sizeof(smem[0]));
        if (kind == 0) {
            if (tile >= sc.TA || freed[slot] < gen * sc.TB) { g_err = "flow: pass-A tile scheduled before its ring slot was free"; return false; }
            ++seenA[(size_t)t * sc.TA + tile];
            EmuPhases<KA, 0>::run(p.a, t * sc.TA + tile, ra, smem.data());
            ++ready[slot];
        } else {
            if (tile >= sc.TB || ready[slot] < (gen + 1u) * sc.TA) { g_err = "flow: pass-B tile scheduled before pass A finished"; return false; }
            ++seenB[(size_t)t * sc.TB + tile];
            EmuPhases<KB, 0>::run(p.b, t * sc.TB + tile, rb, smem.data());
            ++freed[slot];
        }
    }
    for (uint32_t v : seenA) if (v != 1) { g_err = "flow: a pass-A tile did not run exactly once"; return false; }
    for (uint32_t v : seenB) if (v != 1) { g_err = "flow: a pass-B tile did not run exactly once"; return false; }
    return true;
}

// fused single-launch four-step (fused.h): same ticket order and counters as the device kernel, one emulated CTA, tiles
// in ticket order; the store phase of TmaTileKernel (its last phase) stands in for the storer thread
template <class KA, class KB, int NG, int NS>
static int fused_grid() { return 148; }
template <class KA, class KB, int NG, int NS>
static bool launch_fused(const typename FusedKernel<KA, KB, NG, NS>::Params& p, uint64_t ctl_bytes, stream_t) {
    using FK = FusedKernel<KA, KB, NG, NS>;
    using C = cx<typename FK::T>;
    ++g_launches;
    std::memset(p.ctl, 0, ctl_bytes);
    const FlowSched& sc = p.sched;
    uint32_t* ready = p.ctl + FLOW_CTL_HEAD;
    uint32_t* freed = ready + sc.ring_w;
    std::vector<typename KA::Regs> ra((size_t)KA::NT);
    std::vector<typename KB::Regs> rb((size_t)KB::NT);
    std::vector<C> smem(FK::STAGE_BYTES / sizeof(C) + 1);
    std::vector<uint32_t> seenA((size_t)sc.batch * sc.TA, 0), seenB((size_t)sc.batch * sc.TB, 0);
    for (uint32_t ticket = 0; ticket < sc.total; ++ticket) {
        int kind;
        uint32_t t, tile;
        bool valid;
        sc.decode(ticket, kind, t, tile, valid);
        if (!valid) continue;
        const uint32_t slot = t % sc.ring_w, gen = t / sc.ring_w;
        std::memset(smem.data(), 0xff, smem.size() * sizeof(smem[0]));
        if (kind == 0) {
            if (tile >= sc.TA || freed[slot] < gen * sc.TB) { g_err = "fused: pass-A tile scheduled before its ring slot was free"; return false; }
            ++seenA[(size_t)t * sc.TA + tile];
            EmuPhases<KA, 0>::run(p.a, t * sc.TA + tile, ra, smem.data());
            ++ready[slot];
        } else {
            if (tile >= sc.TB || ready[slot] < (gen + 1u) * sc.TA) { g_err = "fused: pass-B tile scheduled before pass A finished"; return false; }
            ++seenB[(size_t)t * sc.TB + tile];
            EmuPhases<KB, 0>::run(p.b, t * sc.TB + tile, rb, smem.data());
            ++freed[slot];
        }
    }
    for (uint32_t v : seenA) if (v != 1) { g_err = "fused: a pass-A tile did not run exactly once"; return false; }
    for (uint32_t v : seenB) if (v != 1) { g_err = "fused: a pass-B tile did not run exactly once"; return false; }
    return true;
}

}  // namespace rt
}  // namespace b2

#include "../../rustfft_b200/csrc/impl.inl"

extern "C" uint64_t b200fft_emu_launch_count(void) { return b2::rt::g_launches; }
