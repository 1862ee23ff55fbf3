// Host side of libb200fft: planner + C ABI, written against the small `rt::` runtime layer that the
// including translation unit provides (b200fft.cu: CUDA runtime, sm_100a kernels;
// tests/emu/b200fft_emu.cpp: a thread-by-thread CPU replay of the same kernel phases, test-only).
//
// Planner (what RustFFT's planners decide in src/plan.rs:412-665 / src/avx/avx_planner.rs:205-216,
// re-decided for a GPU):
//   len 0, 1                      -> Identity
//   2^k <= 16384 (f64: 8192)      -> Direct      one CTA pass, Stockham radix-4/8/16 in registers+smem
//   larger 2^k (<= 2^24)          -> FourStep    two passes, intermediate kept in L2 by chunking
//   prime factors <= 31, n <= 4096 -> Smooth     one CTA pass, run-time radix list (31..11/7/5/3/16/8/4/2)
//   prime n, n-1 = 2^k            -> Rader       (fused single pass when n-1 <= 256, else over FourStep)
//   anything else                 -> Bluestein   M = next_pow2(2n-1)  (fused single pass when M <= 4096)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/b200fft.h"
#include "host_math.h"
#include "fused.h"
#include "cluster.h"
#include "real.h"

namespace b2 {

// (inline: one instance shared by the translation units of the library)
inline thread_local std::string g_last_error;
inline int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// which parts this translation unit compiles (b200fft.cu / b200fft_f32.cu / b200fft_f64.cu; the CPU replay
// harness defines none and gets everything)
#if !defined(B2_PART_CABI) && !defined(B2_PART_F32) && !defined(B2_PART_F64) && !defined(B2_PART_SMOOTH32) && !defined(B2_PART_SMOOTH64) && \
    !defined(B2_PART_FUSED32) && !defined(B2_PART_CTILE32) && !defined(B2_PART_SMOOTH32S) && !defined(B2_PART_SMOOTH64S) && \
    !defined(B2_PART_SMOOTH32P) && !defined(B2_PART_SMOOTH64P)
#define B2_PART_FUSED32 1
#define B2_PART_SMOOTH32P 1
#define B2_PART_SMOOTH64P 1
#define B2_PART_CTILE32 1
#define B2_PART_SMOOTH32S 1
#define B2_PART_SMOOTH64S 1
#define B2_PART_CABI 1
#define B2_PART_F32 1
#define B2_PART_F64 1
#define B2_PART_SMOOTH32 1
#define B2_PART_SMOOTH64 1
#endif

// ---- geometry registry ---------------------------------------------------------------------
// V = tuning variant: 0 = radix <= 16 stages (16 elements per thread), 1 = radix-32 stages (32 per thread)
template <typename T, int L, int V = 0> struct DirectGeo;  // whole transform per CTA pass, threads: j fastest
template <typename T, int L, int V = 0> struct TileGeo;    // four-step tiles: F FFTs side by side

#define B2_DIRECT(T, L, E, F, ...) \
    template <> struct DirectGeo<T, L, 0> { using type = Geo<T, L, E, F, Radices<__VA_ARGS__>>; };
#define B2_TILE(T, L, E, F, ...) \
    template <> struct TileGeo<T, L, 0> { using type = Geo<T, L, E, F, Radices<__VA_ARGS__>>; };
#define B2_DIRECT_V1(T, L, E, F, ...) \
    template <> struct DirectGeo<T, L, 1> { using type = Geo<T, L, E, F, Radices<__VA_ARGS__>>; };
#define B2_TILE_V1(T, L, E, F, ...) \
    template <> struct TileGeo<T, L, 1> { using type = Geo<T, L, E, F, Radices<__VA_ARGS__>>; };
// 1-in-2^PS padding of the 1024-point radix-32 tile: tests/test_engine_model.py predicts 2-way bank conflicts for the
// default PS = 4 (ncu agrees: ~10 % of its shared-memory wavefronts are conflict replays) and none for PS = 5.
// Build-time switch until it has been timed:  make OUT=../libb200fft_ps5.so BUILD=build_ps5 EXTRA=-DB2_TILE1024_PS=5
#if !defined(B2_TILE1024_PS)
#define B2_TILE1024_PS 4
#endif

B2_DIRECT(float, 2, 2, 128, 2)
B2_DIRECT(float, 4, 4, 128, 4)
B2_DIRECT(float, 8, 8, 128, 8)
B2_DIRECT(float, 16, 4, 32, 4, 4)
B2_DIRECT(float, 32, 8, 32, 4, 8)
B2_DIRECT(float, 64, 8, 16, 8, 8)
B2_DIRECT(float, 128, 16, 16, 8, 16)
B2_DIRECT(float, 256, 16, 8, 16, 16)
B2_DIRECT(float, 512, 16, 8, 2, 16, 16)
B2_DIRECT(float, 1024, 16, 4, 4, 16, 16)
B2_DIRECT(float, 2048, 16, 2, 8, 16, 16)
B2_DIRECT(float, 4096, 16, 1, 16, 16, 16)
B2_DIRECT(float, 8192, 16, 1, 2, 16, 16, 16)
B2_DIRECT(float, 16384, 16, 1, 4, 16, 16, 16)

B2_DIRECT(double, 2, 2, 128, 2)
B2_DIRECT(double, 4, 4, 128, 4)
B2_DIRECT(double, 8, 8, 64, 8)
B2_DIRECT(double, 16, 4, 32, 4, 4)
B2_DIRECT(double, 32, 8, 32, 4, 8)
B2_DIRECT(double, 64, 8, 16, 8, 8)
B2_DIRECT(double, 128, 8, 8, 2, 8, 8)
B2_DIRECT(double, 256, 8, 8, 4, 8, 8)
B2_DIRECT(double, 512, 8, 4, 8, 8, 8)
B2_DIRECT(double, 1024, 8, 2, 2, 8, 8, 8)
B2_DIRECT(double, 2048, 8, 1, 4, 8, 8, 8)
B2_DIRECT(double, 4096, 8, 1, 8, 8, 8, 8)
B2_DIRECT(double, 8192, 8, 1, 2, 8, 8, 8, 8)

B2_TILE(float, 64, 8, 16, 8, 8)
B2_TILE(float, 128, 16, 16, 8, 16)
B2_TILE(float, 256, 16, 16, 16, 16)
B2_TILE(float, 512, 16, 16, 2, 16, 16)
B2_TILE(float, 1024, 16, 8, 16, 16, 4)
// radix-32 variants (f32): every four-step pass becomes two stages = one shared-memory exchange
B2_TILE_V1(float, 512, 32, 16, 16, 32)
template <> struct TileGeo<float, 1024, 1> { using type = Geo<float, 1024, 32, 8, Radices<32, 32>, B2_TILE1024_PS>; };
// "narrow" variants (B200FFT_NARROW=1, experiment queued for the next GPU session): half the columns per tile, 128
// threads, so an SM holds FOUR independent 32 KiB tiles in different phases instead of two 64 KiB ones at the same
// register cost; meant to be paired with tensor-map L2 promotion (B200FFT_TMA_L2PROMO=2) so that the first tile to
// touch a 128-byte line brings in its neighbours' 32-byte runs
#define B2_TILE_V2(T, L, E, F, ...) \
    template <> struct TileGeo<T, L, 2> { using type = Geo<T, L, E, F, Radices<__VA_ARGS__>>; };
B2_TILE_V2(float, 512, 32, 8, 16, 32)
B2_TILE_V2(float, 1024, 32, 4, 32, 32)
B2_DIRECT_V1(float, 8192, 32, 1, 16, 16, 32)
B2_DIRECT_V1(float, 16384, 32, 1, 16, 32, 32)

B2_TILE(double, 64, 8, 16, 8, 8)
B2_TILE(double, 128, 8, 16, 2, 8, 8)
B2_TILE(double, 256, 8, 8, 4, 8, 8)
B2_TILE(double, 512, 8, 8, 8, 8, 8)
B2_TILE(double, 1024, 8, 4, 2, 8, 8, 8)
// 2048- and 4096-point tiles: lengths 2^21..2^24 (narrow tiles: 32- / 16-byte runs, the price of a 4096-point
// row or column having to fit one CTA); the radix-32 geometries are the f32 defaults
B2_TILE(float, 2048, 32, 4, 2, 32, 32)
B2_TILE(float, 4096, 32, 2, 4, 32, 32)
B2_TILE(double, 2048, 8, 2, 4, 8, 8, 8)
B2_TILE(double, 4096, 8, 1, 8, 8, 8, 8)

// tiles of the fused single-launch four-step (fused.h): 32 elements per thread, 256 threads per consumer group, every
// tile 8192 elements = 64 KiB whatever the pass length -- (columns | rows) per tile = 8192 / L
template <typename T, int L> struct FusedGeo;
#define B2_FUSED(L, F, ...) \
    template <> struct FusedGeo<float, L> { using type = Geo<float, L, 32, F, Radices<__VA_ARGS__>>; };
B2_FUSED(128, 64, 4, 32)
B2_FUSED(256, 32, 8, 32)
B2_FUSED(512, 16, 16, 32)
template <> struct FusedGeo<float, 1024> { using type = Geo<float, 1024, 32, 8, Radices<32, 32>, B2_TILE1024_PS>; };
B2_FUSED(2048, 4, 2, 32, 32)
B2_FUSED(4096, 2, 4, 32, 32)
static constexpr int FUSED_NG = 2, FUSED_NS = 3;  // consumer groups, shared-memory stages
// half tiles for the cluster plans (4096 points, 128 threads per CTA: four CTAs per SM, twice the cluster size)
template <int L> struct ClusterHalfGeo;
template <> struct ClusterHalfGeo<128> { using type = Geo<float, 128, 32, 32, Radices<4, 32>>; };
template <> struct ClusterHalfGeo<256> { using type = Geo<float, 256, 32, 16, Radices<8, 32>>; };
template <> struct ClusterHalfGeo<512> { using type = Geo<float, 512, 32, 8, Radices<16, 32>>; };  // (only named by the full-tile 2^17 plan's type selection)

// compiled tiles of COMPOSITE pass lengths (f32): the same CTA engine with radix-3 / 5 / 7 stages, for the two-pass plans of
// 10000 = 100 x 100, 44100 = 196 x 225, 48000 = 128 x 375, 100000 = 100 x 1000 and 10^6 = 1000 x 1000 (the run-time-radix
// SmoothPassKernel is instruction bound at ~0.12 of the roofline; every radix must divide the E elements a thread holds)
template <int L> struct SmoothTileGeo;
template <> struct SmoothTileGeo<64> { using type = Geo<float, 64, 8, 16, Radices<8, 8>>; };
template <> struct SmoothTileGeo<100> { using type = Geo<float, 100, 20, 32, Radices<4, 5, 5>>; };
template <> struct SmoothTileGeo<125> { using type = Geo<float, 125, 25, 32, Radices<5, 5, 5>>; };
template <> struct SmoothTileGeo<128> { using type = Geo<float, 128, 16, 16, Radices<8, 16>>; };
template <> struct SmoothTileGeo<196> { using type = Geo<float, 196, 28, 16, Radices<4, 7, 7>>; };
template <> struct SmoothTileGeo<200> { using type = Geo<float, 200, 40, 32, Radices<8, 5, 5>>; };
template <> struct SmoothTileGeo<225> { using type = Geo<float, 225, 15, 16, Radices<3, 3, 5, 5>>; };
template <> struct SmoothTileGeo<250> { using type = Geo<float, 250, 10, 16, Radices<2, 5, 5, 5>>; };
template <> struct SmoothTileGeo<256> { using type = Geo<float, 256, 16, 16, Radices<16, 16>>; };
template <> struct SmoothTileGeo<375> { using type = Geo<float, 375, 15, 16, Radices<3, 5, 5, 5>>; };
template <> struct SmoothTileGeo<400> { using type = Geo<float, 400, 20, 16, Radices<4, 4, 5, 5>>; };
template <> struct SmoothTileGeo<500> { using type = Geo<float, 500, 20, 16, Radices<4, 5, 5, 5>>; };
template <> struct SmoothTileGeo<512> { using type = Geo<float, 512, 16, 16, Radices<2, 16, 16>>; };
template <> struct SmoothTileGeo<625> { using type = Geo<float, 625, 25, 8, Radices<5, 5, 5, 5>>; };
template <> struct SmoothTileGeo<1000> { using type = Geo<float, 1000, 40, 8, Radices<8, 5, 5, 5>>; };
template <> struct SmoothTileGeo<1024> { using type = Geo<float, 1024, 16, 8, Radices<16, 16, 4>>; };
// pass lengths with a compiled tile; a two-pass plan exists for every product a * b of two of them (a <= b)
static constexpr uint32_t COMPILED_TILE_LENGTHS[] = {64, 100, 125, 128, 196, 200, 225, 250, 256, 375, 400, 500, 512, 625, 1000, 1024};


// largest transform one CTA keeps in shared memory: 16384 c32 (136 KiB) / 8192 c64 (136 KiB)
template <typename T> struct DirectMax { static constexpr uint32_t v = sizeof(T) == 4 ? 16384 : 8192; };
static constexpr size_t MAX_SMEM_PER_CTA = 227 * 1024;  // opt-in dynamic shared memory limit of sm_100
static constexpr uint32_t FUSED_CONV_MAX = 4096;  // largest inner FFT of the fused Bluestein / Rader kernels
static constexpr uint32_t TILE_MIN = 64, TILE_MAX = 4096;

// ---- plan object ----------------------------------------------------------------------------
struct ExecCtx {
    const void* in;
    void* out;
    void* work;
    uint64_t batch;
    rt::stream_t stream;
    int max_streams;  // 0 = plan default; the host-slice pipeline passes 1 (it is PCIe bound and already staged)
};

// resources of one host-slice pipeline (exec_host_impl); grow-only, owned by a plan
struct HostPipe {
    static constexpr int NB = 4;  // ring slots
    rt::stream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    rt::event_t ev_in[NB] = {}, ev_k[NB] = {}, ev_out[NB] = {};
    void* dbuf[NB] = {};
    uint64_t dbuf_bytes = 0;
    void* pin_in[NB] = {};   // pinned staging for pageable callers
    void* pin_out[NB] = {};
    uint64_t pin_bytes = 0;
    void* wbuf = nullptr;
    uint64_t wbytes = 0;
};
inline void host_pipe_destroy(HostPipe* hp) {
    if (!hp) return;
    for (int i = 0; i < HostPipe::NB; ++i) {
        if (hp->dbuf[i]) rt::dfree(hp->dbuf[i]);
        if (hp->pin_in[i]) rt::host_free_pinned(hp->pin_in[i]);
        if (hp->pin_out[i]) rt::host_free_pinned(hp->pin_out[i]);
        if (hp->ev_in[i]) rt::event_destroy(hp->ev_in[i]);
        if (hp->ev_k[i]) rt::event_destroy(hp->ev_k[i]);
        if (hp->ev_out[i]) rt::event_destroy(hp->ev_out[i]);
    }
    if (hp->wbuf) rt::dfree(hp->wbuf);
    if (hp->s_in) rt::stream_destroy(hp->s_in);
    if (hp->s_k) rt::stream_destroy(hp->s_k);
    if (hp->s_out) rt::stream_destroy(hp->s_out);
    delete hp;
}

}  // namespace b2

struct b200fft_plan {
    uint64_t len = 0;
    int direction = 0, precision = 0, device = 0;
    std::string desc;
    uint64_t chunk = 0;         // transforms per L2 chunk of multi-pass plans
    std::vector<void*> tables;  // device allocations owned by the plan
    // auxiliary streams for plans that run independent chunks on two streams (created lazily, reused)
    std::mutex aux_mutex;
    std::vector<b2::rt::stream_t> aux_streams;
    b2::rt::stream_t aux_get() {
        std::lock_guard<std::mutex> g(aux_mutex);
        if (!aux_streams.empty()) {
            b2::rt::stream_t s = aux_streams.back();
            aux_streams.pop_back();
            return s;
        }
        return b2::rt::stream_create();
    }
    void aux_put(b2::rt::stream_t s) {
        std::lock_guard<std::mutex> g(aux_mutex);
        aux_streams.push_back(s);
    }
    // host-slice path: pipelines (streams, events, device ring, pinned staging ring, workspace) are created on first use and kept;
    // one per concurrent caller
    std::mutex pipe_mutex;
    std::vector<b2::HostPipe*> pipes;
    std::vector<b200fft_recipe_node> recipe;  // b200fft_plan_create_from_recipe: the caller's decomposition (empty = plan here)
    std::vector<b200fft_recipe_node> chosen;  // the decomposition that was built, in the same vocabulary (b200fft_plan_recipe)
    std::function<bool(const b2::ExecCtx&)> exec;
    std::function<uint64_t(uint64_t)> work_bytes = [](uint64_t) { return (uint64_t)0; };
    std::function<uint64_t(uint64_t)> launches = [](uint64_t) { return (uint64_t)0; };
    ~b200fft_plan() {
        for (void* p : tables) b2::rt::dfree(p);
        for (auto s : aux_streams) b2::rt::stream_destroy(s);
        for (b2::HostPipe* hp : pipes) b2::host_pipe_destroy(hp);
    }
};

namespace b2 {

// records what was built as a recipe (node 0 + optional inner node): b200fft_plan_recipe() hands it back, so a plan can be
// stored as data and rebuilt with b200fft_plan_create_from_recipe()
static void set_recipe(b200fft_plan& pl, uint32_t kind, uint64_t a = 0, uint64_t b = 0, uint32_t child_kind = 0, uint64_t child_len = 0,
                       uint64_t child_a = 0, uint64_t child_b = 0) {
    pl.chosen.clear();
    pl.chosen.push_back(b200fft_recipe_node{kind, child_kind ? 1u : 0u, pl.len, a, b});
    if (child_kind) pl.chosen.push_back(b200fft_recipe_node{child_kind, 0u, child_len, child_a, child_b});
}

template <class V>
static const V* upload(b200fft_plan& pl, const std::vector<V>& host) {
    if (host.empty()) return nullptr;
    void* d = rt::dmalloc(host.size() * sizeof(V));
    if (!d) return nullptr;
    pl.tables.push_back(d);
    if (!rt::h2d_sync(d, host.data(), host.size() * sizeof(V))) return nullptr;
    return reinterpret_cast<const V*>(d);
}

// packed stage twiddles of a geometry: stage s >= 1, layout [(r-1)*p + k] = W_{pR}^{k r}
template <class G>
static std::vector<cx<typename G::T>> stage_twiddles() {
    using T = typename G::T;
    using RL = typename G::RL;
    std::vector<cx<T>> tw((size_t)RL::tw_total());
    int p = RL::get(0);
    for (int s = 1; s < RL::N; ++s) {
        const int R = RL::get(s);
        const int off = RL::tw_offset(s);
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < p; ++k) tw[(size_t)off + (size_t)(r - 1) * p + k] = hm::twiddle<T>((uint64_t)k * r, (uint64_t)p * R);
        p *= R;
    }
    return tw;
}

static bool chunk_bytes_forced() { return std::getenv("B200FFT_CHUNK_MB") != nullptr; }
static uint64_t chunk_bytes() {
    // target footprint of the L2-resident intermediate of multi-pass plans (B200 L2: ~126 MB)
    static uint64_t v = [] {
        const char* e = std::getenv("B200FFT_CHUNK_MB");
        uint64_t mb = e ? std::strtoull(e, nullptr, 10) : 64;
        if (mb < 1) mb = 1;
        return mb << 20;
    }();
    return v;
}

// B200FFT_PIPELINE=1 selects the persistent TMA-pipelined kernels (cp.async.bulk + mbarrier double
// buffering).  Measured in round 1 (profiles/r1e_*): correct, but 10-13 % slower than the plain kernels --
// the second tile buffer halves the resident CTAs of the 70 KiB tiles and the dense tile costs one more
// shared-memory pass -- so they stay opt-in until that is fixed.
static bool use_pipelined() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_PIPELINE");
        return e && std::atoi(e) == 1;
    }();
    return v;
}

// B200FFT_OVERLAP=0: all chunks of a multi-pass plan on the caller's stream.  Default: consecutive chunks
// alternate between the caller's stream and an auxiliary one (two workspaces), so the tail of one chunk's
// launches overlaps the head of the next chunk's -- chunks are independent, only pass A -> pass B within a
// chunk is ordered.
static bool use_overlap() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_OVERLAP");
        return !(e && std::atoi(e) == 0);
    }();
    return v;
}
// number of streams the chunks of a multi-pass plan are spread over: B200FFT_STREAMS (1..4) or, by default, what
// measured best on B200 for the plan kind (profiles/r2a): 4 for two-pass plans below 2^20, 3 from 2^20 on, 2 for the
// four-pass convolution plans
static int overlap_streams(int auto_default = 2) {
    static int forced = [] {
        if (!use_overlap()) return 1;
        const char* e = std::getenv("B200FFT_STREAMS");
        const int k = e ? std::atoi(e) : 0;
        return k < 1 ? 0 : (k > 4 ? 4 : k);
    }();
    return forced ? forced : auto_default;
}

// B200FFT_FLOW=1: two-pass plans run as ONE launch of the dataflow kernel (kernels.h, run_flow) instead of one launch pair
// per L2 chunk.  Opt-in: measured on B200 (profiles/r1t) it reaches 0.30-0.53 of the HBM roofline against 0.42-0.55 for
// the chunked path -- the per-tile dependency/ticket/fence work costs more than the launch ramps and tails it removes.
// B200FFT_FLOW_LOOKAHEAD = tickets pass A runs ahead of pass B (default 700), B200FFT_FLOW_W forces the ring slots.
static bool use_flow() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_FLOW");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
static uint32_t flow_ring_slots(uint32_t per_round, uint64_t bytes_per_transform) {
    static const uint32_t forced = [] {
        const char* e = std::getenv("B200FFT_FLOW_W");
        return e ? (uint32_t)std::atoi(e) : 0u;
    }();
    static const uint32_t look = [] {
        const char* e = std::getenv("B200FFT_FLOW_LOOKAHEAD");
        const int k = e ? std::atoi(e) : 700;
        return (uint32_t)(k < 1 ? 1 : k);
    }();
    if (forced >= 2) return forced;
    // pass A runs D = W/2 rounds ahead of pass B and a slot is reused W/2 rounds after pass B read it: both gaps
    // must exceed the tickets that are claimed but unfinished (~1.5 per resident CTA, ~450) or tiles stall
    uint32_t D = (look + per_round - 1) / per_round;
    if (D < 1) D = 1;
    uint32_t W = 2 * D;
    // the ring must stay L2 resident: at most 64 MiB, but never fewer than two slots
    while (W > 2 && (uint64_t)W * bytes_per_transform > (64ull << 20)) W -= 2;
    return W;
}

// Power-of-two two-pass plans (f32) run as ONE launch of the fused warp-specialised kernel (fused.h) whenever the
// buffers allow TMA; B200FFT_FUSED=0 selects the chunked launch pairs instead.  B200FFT_FUSED_LOOKAHEAD = tickets pass A
// runs ahead of pass B (default 1000 > the ~900 tickets in flight or claimed on 148 SMs), B200FFT_FUSED_W forces the ring slots,
// B200FFT_FUSED_NOCOMPUTE=1 skips the butterflies (memory-pipeline ceiling; results are garbage).
static bool use_fused() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_FUSED");
        return !(e && std::atoi(e) == 0) && rt::tma_available();
    }();
    return v;
}
// B200FFT_FUSED_HINTS = bit mask of L2 eviction hints on the kernel's TMA copies (1: input loads evict-first, 2: ring loads
// evict-first, 4: ring stores evict-last, 8: output stores evict-first); default 15 (measured +3..8 % over no hints)
static uint32_t fused_flags() {
    static uint32_t v = [] {
        const char* e = std::getenv("B200FFT_FUSED_NOCOMPUTE");
        const char* h = std::getenv("B200FFT_FUSED_HINTS");
        const char* x = std::getenv("B200FFT_FUSED_XFLAGS");  // experiment bits (fused.h), shifted above the hints
        return ((e && std::atoi(e) == 1) ? 1u : 0u) | ((h ? (uint32_t)std::atoi(h) & 15u : 15u) << 1) | ((x ? (uint32_t)std::atoi(x) : 0u) << 5);
    }();
    return v;
}
// B200FFT_FUSED_TILED=1: tile-major ring (pass A stores from its registers with coalesced 8-byte stores, pass B gathers its rows with a
// 4-D tensor copy) instead of the [k1][n2] layout of the chunked path (pass A stores its tile with a TMA tensor store, pass B reads
// contiguous rows).  Measured on B200 (profiles/): slower -- 64 KiB of register->global stores stall a consumer group for ~1.7 us, the TMA
// store does the same work in the background -- so it is opt-in.
static bool fused_tiled() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_FUSED_TILED");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
// B200FFT_FUSED_TW2=1: two-level inter-pass twiddles from 16 KiB of L1-resident tables instead of the N-entry table streamed from L2
// (measured: no gain -- the table loads are prefetched while the tile is in flight -- and one more rounding; opt-in)
static bool fused_tw2() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_FUSED_TW2");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
// B200FFT_FUSED_BDIRECT = bit mask over log2 N - 15: pass B of the fused kernel stores its results from the registers (coalesced 64-256
// byte runs) instead of through a TMA store of the stage; measured per size (profiles/), default = the sizes where it won
static bool fused_bdirect(uint32_t lgN) {
    static uint32_t v = [] {
        const char* e = std::getenv("B200FFT_FUSED_BDIRECT");
        return e ? (uint32_t)std::atoi(e) : 0u;
    }();
    return lgN >= 15 && lgN < 47 && ((v >> (lgN - 15)) & 1u);
}
static bool fused_trace() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_FUSED_TRACE");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
static uint32_t fused_ring_slots(uint32_t per_round, uint64_t bytes_per_transform) {
    static const uint32_t forced = [] {
        const char* e = std::getenv("B200FFT_FUSED_W");
        return e ? (uint32_t)std::atoi(e) : 0u;
    }();
    static const uint32_t look = [] {
        const char* e = std::getenv("B200FFT_FUSED_LOOKAHEAD");
        const int k = e ? std::atoi(e) : 1000;
        return (uint32_t)(k < 1 ? 1 : k);
    }();
    if (forced >= 2) return forced;
    // measured ring sweep (profiles/r2j_ab_fused_ring_size.txt): 2^18 is best at 32 slots (64 MiB: 0.591 against 0.485 at 16 and 0.572 at
    // 48), 2^19 at 16-20 slots (0.578 / 0.582), 2^20 at 10 slots = 80 MiB (0.561 against 0.541 at 8 and 0.530 at 12): from 8 MiB
    // transforms on, one more pair of slots than the 1000-ticket look-ahead asks for, and the residency cap that allows it
    const bool big = bytes_per_transform >= (8ull << 20);
    const uint32_t lk = big ? std::max<uint32_t>(look, 1280u) : look;
    const uint64_t cap = big ? (80ull << 20) : (64ull << 20);
    uint32_t D = (lk + per_round - 1) / per_round;
    if (D < 1) D = 1;
    uint32_t W = 2 * D;
    while (W > 2 && (uint64_t)W * bytes_per_transform > cap) W -= 2;  // the ring must stay L2 resident
    return W;
}

// The chunked two-pass plans move their tiles with TMA tensor copies (kernels.h, TmaTileKernel) whenever the caller's
// buffers are 16-byte aligned and the driver exports cuTensorMapEncodeTiled; B200FFT_TMA_TILES=0 selects the LDG/STG
// passes instead (measured on B200, profiles/r2a: TMA tiles +3..11 % at 2^17..2^20).
static bool use_tma_tiles() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_TMA_TILES");
        return !(e && std::atoi(e) == 0) && rt::tma_available();
    }();
    return v;
}

// B200FFT_DISCARD=0: keep the consumed intermediate of two-pass plans in L2 until it is evicted (and written back to HBM).
// Default: the second pass drops the lines it has read with discard.global.L2 -- they are dead scratch data.
static bool use_discard() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_DISCARD");
        return !(e && std::atoi(e) == 0);
    }();
    return v;
}

// B200FFT_RADIX32=0 disables the radix-32 geometries (A/B measurements)
static bool use_radix32() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_RADIX32");
        return !(e && std::atoi(e) == 0);
    }();
    return v;
}
// B200FFT_PREFETCH=1: pass B of a chunk issues cp.async.bulk.prefetch.L2 for the input of the same stream's next chunk
// (DRAM is ~50 % busy in these plans and a pass-A tile spends most of its life waiting for DRAM); queued for a timed A/B,
// to be tried with a smaller B200FFT_CHUNK_MB since the prefetched input shares L2 with the intermediates
static bool use_prefetch() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_PREFETCH");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
// B200FFT_PERSIST=1: Direct{16384} runs as a persistent kernel (rt::launch_persistent); queued for a timed A/B
static bool use_persistent() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_PERSIST");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
static bool use_narrow() {
    static bool v = [] {
        const char* e = std::getenv("B200FFT_NARROW");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
template <typename T, int L> struct HasV2 { static constexpr bool tile = false; };
template <> struct HasV2<float, 512> { static constexpr bool tile = true; };
template <> struct HasV2<float, 1024> { static constexpr bool tile = true; };
template <typename T, int L> struct HasV1 { static constexpr bool direct = false, tile = false; };
template <> struct HasV1<float, 512> { static constexpr bool direct = false, tile = true; };   // Direct{512,1024}:
template <> struct HasV1<float, 1024> { static constexpr bool direct = false, tile = true; };  // radix-16 measured faster
template <> struct HasV1<float, 8192> { static constexpr bool direct = true, tile = false; };
template <> struct HasV1<float, 16384> { static constexpr bool direct = true, tile = false; };

// The run-time-radix kernels (Smooth, SmoothFourStep: 14 butterfly sizes x 8 stages each) are instantiated in
// translation units of their own (b200fft_smooth32.cu / b200fft_smooth64.cu) so the library builds in parallel.
// kind 0: one-pass Smooth plan of pl.len;  kind 1: SmoothFourStep{a x b}
bool build_smooth_f32(b200fft_plan& pl, int kind, uint32_t a, uint32_t b);
bool build_smooth_f64(b200fft_plan& pl, int kind, uint32_t a, uint32_t b);
bool build_smooth_small_f32(b200fft_plan& pl, int kind, uint32_t a, uint32_t b);
bool build_smooth_small_f64(b200fft_plan& pl, int kind, uint32_t a, uint32_t b);
bool build_smooth_passes_f32(b200fft_plan& pl, int kind, uint32_t a, uint32_t b);
bool build_smooth_passes_f64(b200fft_plan& pl, int kind, uint32_t a, uint32_t b);
// single-pass cluster plans (cluster.h), compiled in the same translation unit as the fused kernels
bool build_cluster_f32(b200fft_plan& pl, uint32_t lgN, bool half);
// compiled composite tiles (b200fft_ctile32.cu)
bool build_compiled_smooth_f32(b200fft_plan& pl, uint32_t a, uint32_t b);
bool build_cluster_conv_f32(b200fft_plan& pl, uint64_t M, int mode);
// likewise the fused single-launch four-step kernels (b200fft_fused32.cu)
typedef std::function<bool(const void* in, void* out, void* work, uint64_t batch, rt::stream_t)> FusedFn;
bool build_fused_f32(b200fft_plan& pl, uint32_t L1, uint32_t L2, uint32_t lgN, const void* full_tw, FusedFn& fn, uint32_t& W);

template <typename T>
struct Builder {
    typedef cx<T> C;
    // the run-time-radix kernels exist in two instantiations -- with the prime butterflies 11..31 (128 registers, 2 CTAs per SM) and
    // without them (3 CTAs per SM) -- each in translation units of its own; which one a plan needs is decided here
    static bool smooth_kind_is_small(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) {
        std::vector<uint32_t> r;
        switch (kind) {
            case 0: return smooth_factor(pl.len, r) && max_radix(r) <= 16;
            case 2: return smooth_factor((uint64_t)b - 1, r) && max_radix(r) <= 16;
            case 3: return smooth_factor(a, r) && max_radix(r) <= 16;
            case 7: return smooth_factor(a, r) && max_radix(r) <= 16;
            default: return small_radices_only(a, b);
        }
    }
    static bool smooth_dispatch(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) {
        const bool small = smooth_kind_is_small(pl, kind, a, b);
        const bool two_pass = kind == 1 || kind == 4 || kind == 5 || kind == 6 || kind == 7;
        if constexpr (sizeof(T) == 4) {
            if (small) return build_smooth_small_f32(pl, kind, a, b);
            return two_pass ? build_smooth_passes_f32(pl, kind, a, b) : build_smooth_f32(pl, kind, a, b);
        } else {
            if (small) return build_smooth_small_f64(pl, kind, a, b);
            return two_pass ? build_smooth_passes_f64(pl, kind, a, b) : build_smooth_f64(pl, kind, a, b);
        }
    }
    // kind 0: Smooth; 1: SmoothFourStep{a x b}; 2: one-pass Rader, outer radix a, prime b; 3: one-pass Bluestein, inner M = a;
    //      4: GoodThomas{a x b}; 5: Rader over SmoothFourStep{a x b}; 6: Bluestein over SmoothFourStep{a x b}
    // (one-pass kinds 0 / 2 / 3 and two-pass kinds 1 / 4 / 5 / 6 are instantiated in different translation units: build time)
    template <int RMAX>
    static bool smooth_build_onepass(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) {
        const bool sw = pl.direction != 0;
        if (kind == 2) return sw ? make_smooth_conv_t<true, RMAX>(pl, 0, a, b) : make_smooth_conv_t<false, RMAX>(pl, 0, a, b);
        if (kind == 3) return sw ? make_smooth_conv_t<true, RMAX>(pl, 1, 1, a) : make_smooth_conv_t<false, RMAX>(pl, 1, 1, a);
        std::vector<uint32_t> radices;
        if (kind != 0 || !smooth_factor(pl.len, radices)) return false;
        return sw ? make_smooth_t<true, RMAX>(pl, radices) : make_smooth_t<false, RMAX>(pl, radices);
    }
    // stand-alone column pass of the 2-D plans: H-point FFTs down the columns of row-major [H][W] images, in place or out of place
    template <bool SW, int RMAX>
    static bool make_smooth_columns_t(b200fft_plan& pl, uint32_t H, uint32_t W) {
        using KA = SmoothPassKernel<T, SW, 1, RMAX>;
        std::vector<uint32_t> ra;
        if (H > SMOOTH_MAX || !smooth_factor(H, ra)) return false;
        typename KA::Params pa;
        if (!fill_smooth_pass<KA>(pl, pa, H, ra)) return false;
        const uint64_t N = (uint64_t)H * W;
        pa.NN = N;
        pa.other = W;
        pa.div_other = make_fastdiv(W);
        const uint32_t FA = std::max<uint32_t>(1, std::min<uint32_t>(64, SMOOTH_MAX / H));
        pa.f_per_cta = FA;
        pa.pitch = H;
        pa.div_f = make_fastdiv(FA);
        pa.smem_bytes = ra.size() > 1 ? (uint32_t)(2ull * FA * H * sizeof(C)) : 0;
        pa.swap_out = SW ? 1u : 0u;
        const size_t max_smem = 2ull * (SMOOTH_MAX + 64) * sizeof(C);
        const uint64_t seg = std::max<uint64_t>(1, ((1ull << 31) - 1) / W);  // FFT indices of a launch stay below 2^31
        pl.exec = [=](const ExecCtx& c) {
            for (uint64_t b0 = 0; b0 < c.batch; b0 += seg) {
                const uint64_t nb = std::min(seg, c.batch - b0);
                typename KA::Params q = pa;
                q.in = (const C*)c.in + b0 * N;
                q.out = (C*)c.out + b0 * N;
                q.n_fft = nb * W;
                if (!rt::launch_dyn<KA>(q, (q.n_fft + FA - 1) / FA, q.smem_bytes, max_smem, c.stream)) return false;
            }
            return true;
        };
        pl.launches = [=](uint64_t batch) { return (batch + seg - 1) / seg; };
        pl.desc = "Columns{" + std::to_string(H) + " down [" + std::to_string(H) + "x" + std::to_string(W) + "]}";
        set_recipe(pl, B200FFT_RECIPE_COLUMNS, H, W);
        return true;
    }
    template <int RMAX>
    static bool smooth_build_twopass(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) {
        const bool sw = pl.direction != 0;
        if (kind == 7) return sw ? make_smooth_columns_t<true, RMAX>(pl, a, b) : make_smooth_columns_t<false, RMAX>(pl, a, b);
        if (kind == 1 || kind == 4)
            return sw ? make_smooth_four_step_t<true, RMAX>(pl, a, b, kind == 4 ? 1 : 0) : make_smooth_four_step_t<false, RMAX>(pl, a, b, kind == 4 ? 1 : 0);
        if (kind == 5 || kind == 6)
            return sw ? make_smooth_big_conv_t<true, RMAX>(pl, a, b, kind == 5) : make_smooth_big_conv_t<false, RMAX>(pl, a, b, kind == 5);
        return false;
    }

    // ---------------- Direct ----------------
    template <int L, bool SW, int V = 0>
    static bool make_direct_t(b200fft_plan& pl) {
        using G = typename DirectGeo<T, L, V>::type;
        using KT = FftKernel<G, JF, JF, LoadRows<T, SW>, StoreRows<T, SW>>;
        const C* tw = nullptr;
        if (G::TW_ELEMS) {
            tw = upload(pl, stage_twiddles<G>());
            if (!tw) return false;
        }
        pl.exec = [tw](const ExecCtx& c) {
            if constexpr (G::NS >= 2 && 2 * G::F * G::LP * sizeof(C) + 2048 <= MAX_SMEM_PER_CTA) {
                // persistent TMA-pipelined kernel whenever the caller's buffer allows 16-byte bulk copies
                if (use_pipelined() && (reinterpret_cast<uintptr_t>(c.in) & 15u) == 0) {
                    using KP = PipeKernel<G, JF, XformSwap<T, SW>, StoreRows<T, SW>>;
                    typename KP::Params q;
                    q.in = (const C*)c.in;
                    q.xform = XformSwap<T, SW>{};
                    q.store = StoreRows<T, SW>{(C*)c.out, (uint32_t)L};
                    q.tw = tw;
                    q.n_fft = c.batch;
                    q.n_items = (uint32_t)((c.batch + G::F - 1) / G::F);
                    return rt::launch_pipelined<KP>(q, c.stream);
                }
            }
            typename KT::Params p;
            p.load = LoadRows<T, SW>{(const C*)c.in, (uint32_t)L};
            p.store = StoreRows<T, SW>{(C*)c.out, (uint32_t)L};
            p.tw = tw;
            p.n_fft = c.batch;
            if constexpr (L >= 16384) {  // one CTA per SM: the persistent form is the experiment queued for these
                if (use_persistent()) return rt::launch_persistent<KT>(p, (c.batch + G::F - 1) / G::F, c.stream);
            }
            return rt::launch<KT>(p, (c.batch + G::F - 1) / G::F, c.stream);
        };
        pl.launches = [](uint64_t) { return (uint64_t)1; };
        pl.desc = "Direct{" + std::to_string(L) + "}";
        set_recipe(pl, B200FFT_RECIPE_POW2);
        return true;
    }
    template <int L>
    static bool make_direct(b200fft_plan& pl) {
        if constexpr (HasV1<T, L>::direct) {
            if (use_radix32()) return pl.direction ? make_direct_t<L, true, 1>(pl) : make_direct_t<L, false, 1>(pl);
        }
        return pl.direction ? make_direct_t<L, true>(pl) : make_direct_t<L, false>(pl);
    }
    static bool make_direct_rt(b200fft_plan& pl, uint32_t L) {
        switch (L) {
            case 2: return make_direct<2>(pl);
            case 4: return make_direct<4>(pl);
            case 8: return make_direct<8>(pl);
            case 16: return make_direct<16>(pl);
            case 32: return make_direct<32>(pl);
            case 64: return make_direct<64>(pl);
            case 128: return make_direct<128>(pl);
            case 256: return make_direct<256>(pl);
            case 512: return make_direct<512>(pl);
            case 1024: return make_direct<1024>(pl);
            case 2048: return make_direct<2048>(pl);
            case 4096: return make_direct<4096>(pl);
            case 8192: return make_direct<8192>(pl);
            case 16384:
                if constexpr (sizeof(T) == 4) return make_direct<16384>(pl);
        }
        return false;
    }

    // ---------------- FourStep ----------------
    // N = N1*N2.  pass A: N2 strided columns, N1-point FFT each, times W_N^(n2 k1), written to the
    // workspace in the same [k1][n2] layout; pass B: N1 contiguous rows, N2-point FFT each, written
    // transposed: X[k1 + N1 k2].  Chunked over the batch so the workspace stays L2 resident.
    struct PassFns {
        std::function<bool(const C* in, C* work, uint64_t nb, rt::stream_t)> a;
        std::function<bool(const C* work, C* out, uint64_t nb, rt::stream_t)> b;
        uint64_t ctas_per_transform_a = 0, ctas_per_transform_b = 0;  // tiles per transform
        int wave_a = 0, wave_b = 0;                                    // resident CTAs of each kernel
        // TMA-tiled variants (tensor maps are built per exec call: they hold the caller's pointers)
        std::function<bool(const TMap& m_in, const TMap& m_ws, const C* in, C* work, uint32_t z_in, uint64_t nb, rt::stream_t)> a_tma;
        std::function<bool(const TMap& m_out, const C* work, C* out, uint32_t z_out, uint64_t nb, const void* pf, uint64_t pf_total, rt::stream_t)> b_tma;
        uint32_t f_a = 0, f_b = 0, box_a = 0, box_b = 0;  // tile width and box rows of each pass
    };
    // transforms per L2 chunk: inside [24, 80] MiB of workspace, as close to whole waves as possible for
    // both passes (a 1024-CTA launch on 296 resident CTAs runs 3.46 waves = 13 % idle; 1152 CTAs run 3.89)
    static uint64_t pick_chunk(uint64_t bytes_per_transform, const PassFns& f) {
        if (chunk_bytes_forced()) return std::max<uint64_t>(1, chunk_bytes() / bytes_per_transform);
        // cost per transform = (whole waves of pass A + whole waves of pass B + one wave-time of fixed
        // launch/drain overhead per launch) / nb, searched over 32..64 MiB of workspace (above ~64 MiB the
        // second pass starts missing L2: measured, profiles/r1g_*)
        const uint64_t lo = std::max<uint64_t>(1, (32ull << 20) / bytes_per_transform);
        const uint64_t hi = std::max<uint64_t>(lo, (64ull << 20) / bytes_per_transform);
        uint64_t best = hi;
        double best_cost = 1e300;
        for (uint64_t nb = hi; nb >= lo; --nb) {
            double cost = 0.0;
            const uint64_t cta[2] = {nb * f.ctas_per_transform_a, nb * f.ctas_per_transform_b};
            const int wave[2] = {f.wave_a, f.wave_b};
            for (int i = 0; i < 2; ++i) cost += (wave[i] > 0 ? std::ceil((double)cta[i] / wave[i]) : 1.0) + 1.0;
            cost /= (double)nb;
            if (cost < best_cost - 1e-12) {
                best_cost = cost;
                best = nb;
            }
            if (nb == 1) break;
        }
        return best;
    }
    template <int L1, bool SW, int V = 0>
    static bool make_pass_a(b200fft_plan& pl, uint32_t lgN, uint32_t lg2, PassFns& fns) {
        if constexpr (V == 0 && HasV2<T, L1>::tile) {
            if (use_narrow()) return make_pass_a<L1, SW, 2>(pl, lgN, lg2, fns);
        }
        if constexpr (V == 0 && HasV1<T, L1>::tile) {
            if (use_radix32()) return make_pass_a<L1, SW, 1>(pl, lgN, lg2, fns);
        }
        using G = typename TileGeo<T, L1, V>::type;
        using KT = FftKernel<G, FF, FF, LoadCols<T, SW>, StoreCols<T>>;
        const C* tw = upload(pl, stage_twiddles<G>());
        if (!tw) return false;
        fns.ctas_per_transform_a = ((1ull << lg2) + G::F - 1) / G::F;
        fns.wave_a = rt::resident_ctas<KT>();
        fns.a = [=](const C* in, C* work, uint64_t nb, rt::stream_t s) {
            typename KT::Params p;
            p.load = LoadCols<T, SW>{in, lgN, lg2};
            p.store = StoreCols<T>{work, lgN, lg2};
            p.tw = tw;
            p.n_fft = nb << lg2;
            return rt::launch<KT>(p, (p.n_fft + G::F - 1) / G::F, s);
        };
        if constexpr (G::NS >= 2) {
            using KM = TmaTileKernel<G, FF, FF, 0, SW>;
            fns.f_a = G::F;
            fns.box_a = KM::BOX_ROWS;
            fns.a_tma = [=](const TMap& m_in, const TMap& m_ws, const C* in, C* work, uint32_t z_in, uint64_t nb, rt::stream_t s) {
                typename KM::Params p;
                p.map_in = m_in;
                p.map_out = m_ws;
                p.in = in;
                p.out = work;
                p.tw = tw;
                p.full_tw = nullptr;
                p.tw_lo = p.tw_hi = nullptr;
                p.n_fft = nb << lg2;
                p.lgN = lgN;
                p.lg_other = lg2;
                p.z_in = z_in;
                p.z_out = 0;
                p.discard = 0;
                p.pf = nullptr;
                p.pf_bytes = 0;
                p.ring_w = 0;
                return rt::launch_tma<KM>(p, p.n_fft / G::F, s);
            };
        }
        return true;
    }
    template <int L2, bool SW, int V = 0>
    static bool make_pass_b(b200fft_plan& pl, uint32_t lgN, uint32_t lg1, const C* full_tw, PassFns& fns) {
        if constexpr (V == 0 && HasV2<T, L2>::tile) {
            if (use_narrow()) return make_pass_b<L2, SW, 2>(pl, lgN, lg1, full_tw, fns);
        }
        if constexpr (V == 0 && HasV1<T, L2>::tile) {
            if (use_radix32()) return make_pass_b<L2, SW, 1>(pl, lgN, lg1, full_tw, fns);
        }
        using G = typename TileGeo<T, L2, V>::type;
        using KT = FftKernel<G, JF, FF, LoadRowsTw<T>, StoreTransposed<T, SW>>;
        const C* tw = upload(pl, stage_twiddles<G>());
        if (!tw) return false;
        fns.ctas_per_transform_b = ((1ull << lg1) + G::F - 1) / G::F;
        fns.wave_b = rt::resident_ctas<KT>();
        fns.b = [=](const C* work, C* out, uint64_t nb, rt::stream_t s) {
            if (use_pipelined() && (reinterpret_cast<uintptr_t>(work) & 15u) == 0) {
                using KP = PipeKernel<G, FF, XformRowTw<T>, StoreTransposed<T, SW>>;
                typename KP::Params q;
                q.in = work;
                q.xform = XformRowTw<T>{full_tw, (uint32_t)G::L, lg1};
                q.store = StoreTransposed<T, SW>{out, lgN, lg1};
                q.tw = tw;
                q.n_fft = nb << lg1;
                q.n_items = (uint32_t)((q.n_fft + G::F - 1) / G::F);
                return rt::launch_pipelined<KP>(q, s);
            }
            typename KT::Params p;
            p.load = LoadRowsTw<T>{work, full_tw, (uint32_t)G::L, lg1, use_discard() ? 1u : 0u};
            p.store = StoreTransposed<T, SW>{out, lgN, lg1};
            p.tw = tw;
            p.n_fft = nb << lg1;
            return rt::launch<KT>(p, (p.n_fft + G::F - 1) / G::F, s);
        };
        if constexpr (G::NS >= 2) {
            using KM = TmaTileKernel<G, JF, FF, 1, SW>;
            fns.f_b = G::F;
            fns.box_b = KM::BOX_ROWS;
            fns.b_tma = [=](const TMap& m_out, const C* work, C* out, uint32_t z_out, uint64_t nb, const void* pf, uint64_t pf_total, rt::stream_t s) {
                typename KM::Params p;
                p.map_in = TMap{};
                p.map_out = m_out;
                p.in = work;
                p.out = out;
                p.tw = tw;
                p.full_tw = full_tw;
                p.tw_lo = p.tw_hi = nullptr;
                p.n_fft = nb << lg1;
                p.lgN = lgN;
                p.lg_other = lg1;
                p.z_in = 0;
                p.z_out = z_out;
                p.discard = use_discard() ? 1u : 0u;
                const uint64_t ctas = p.n_fft / G::F;
                const uint64_t share = (pf && ctas) ? std::min<uint64_t>((pf_total / ctas) & ~15ull, 1u << 20) : 0;
                p.pf = share ? pf : nullptr;
                p.pf_bytes = (uint32_t)share;
                p.ring_w = 0;
                return rt::launch_tma<KM>(p, ctas, s);
            };
        }
        return true;
    }
    template <bool SW>
    static bool make_pass_a_rt(b200fft_plan& pl, uint32_t L1, uint32_t lgN, uint32_t lg2, PassFns& f) {
        switch (L1) {
            case 64: return make_pass_a<64, SW>(pl, lgN, lg2, f);
            case 128: return make_pass_a<128, SW>(pl, lgN, lg2, f);
            case 256: return make_pass_a<256, SW>(pl, lgN, lg2, f);
            case 512: return make_pass_a<512, SW>(pl, lgN, lg2, f);
            case 1024: return make_pass_a<1024, SW>(pl, lgN, lg2, f);
            case 2048: return make_pass_a<2048, SW>(pl, lgN, lg2, f);
            case 4096: return make_pass_a<4096, SW>(pl, lgN, lg2, f);
        }
        return false;
    }
    template <bool SW>
    static bool make_pass_b_rt(b200fft_plan& pl, uint32_t L2, uint32_t lgN, uint32_t lg1, const C* tw, PassFns& f) {
        switch (L2) {
            case 64: return make_pass_b<64, SW>(pl, lgN, lg1, tw, f);
            case 128: return make_pass_b<128, SW>(pl, lgN, lg1, tw, f);
            case 256: return make_pass_b<256, SW>(pl, lgN, lg1, tw, f);
            case 512: return make_pass_b<512, SW>(pl, lgN, lg1, tw, f);
            case 1024: return make_pass_b<1024, SW>(pl, lgN, lg1, tw, f);
            case 2048: return make_pass_b<2048, SW>(pl, lgN, lg1, tw, f);
            case 4096: return make_pass_b<4096, SW>(pl, lgN, lg1, tw, f);
        }
        return false;
    }
    // inter-pass twiddles W_N^(k1*n2), laid out [k1][n2] (N entries)
    static const C* make_full_twiddles(b200fft_plan& pl, uint32_t lg1, uint32_t lg2) {
        const uint64_t N1 = 1ull << lg1, N2 = 1ull << lg2, N = N1 * N2;
        std::vector<C> t((size_t)N);
        for (uint64_t k1 = 0; k1 < N1; ++k1)
            for (uint64_t n2 = 0; n2 < N2; ++n2) t[(size_t)(k1 * N2 + n2)] = hm::twiddle<T>(k1 * n2, N);
        return upload(pl, t);
    }
    // ---- single-launch dataflow variant (kernels.h: run_flow) ----
    template <int L> struct FlowGeo { using type = typename TileGeo<T, L, HasV1<T, L>::tile ? 1 : 0>::type; };
    typedef std::function<bool(const C* in, C* out, void* work, uint64_t batch, rt::stream_t)> FlowFn;
    template <int L1, int L2, bool SW>
    static bool make_flow_t(b200fft_plan& pl, uint32_t lgN, const C* full_tw, FlowFn& fn, uint32_t& W_out) {
        using GA = typename FlowGeo<L1>::type;
        using GB = typename FlowGeo<L2>::type;
        using KA = FftKernel<GA, FF, FF, LoadCols<T, SW>, StoreColsRing<T>>;
        using KB = FftKernel<GB, JF, FF, LoadRowsTwRing<T>, StoreTransposed<T, SW>>;
        using FK = FlowKernel<KA, KB>;
        const uint32_t lg1 = hm::ilog2(L1), lg2 = hm::ilog2(L2);
        const C* twa = upload(pl, stage_twiddles<GA>());
        const C* twb = upload(pl, stage_twiddles<GB>());
        if (!twa || !twb) return false;
        if (rt::flow_grid<KA, KB>() <= 0) return false;
        const uint32_t TA = (uint32_t)L2 / GA::F, TB = (uint32_t)L1 / GB::F;
        const uint64_t N = 1ull << lgN;
        const uint32_t W = flow_ring_slots(TA + TB, N * sizeof(C));
        W_out = W;
        const uint64_t ctl_bytes = flow_ctl_bytes(W);
        fn = [=](const C* in, C* out, void* work, uint64_t batch, rt::stream_t s) {
            // 2^24 transforms per launch keeps the ticket count inside 32 bits for every geometry
            const uint64_t seg = 1ull << 24;
            for (uint64_t b0 = 0; b0 < batch; b0 += seg) {
                const uint64_t nb = std::min(seg, batch - b0);
                typename FK::Params p;
                C* ring = (C*)((char*)work + ctl_bytes);
                p.a.load = LoadCols<T, SW>{in + b0 * N, lgN, lg2};
                p.a.store = StoreColsRing<T>{ring, lgN, lg2, W};
                p.a.tw = twa;
                p.a.n_fft = nb << lg2;
                p.b.load = LoadRowsTwRing<T>{ring, full_tw, (uint32_t)L2, lg1, lgN, W, use_discard() ? 1u : 0u};
                p.b.store = StoreTransposed<T, SW>{out + b0 * N, lgN, lg1};
                p.b.tw = twb;
                p.b.n_fft = nb << lg1;
                p.ctl = (uint32_t*)work;
                p.trace = nullptr;
#if defined(B2_FLOW_TRACE)
                p.trace = (unsigned long long*)((char*)work + ctl_bytes + (uint64_t)W * N * sizeof(C));
                if (!rt::memset_async(p.trace, 0, (size_t)FLOW_TRACE_CTAS * FLOW_TRACE_WORDS * 8, s)) return false;
#endif
                if (!make_flow_sched(p.sched, nb, TA, TB, W)) {
                    rt::g_err = "dataflow schedule overflow";
                    return false;
                }
                if (!rt::launch_flow<KA, KB>(p, ctl_bytes, s)) return false;
            }
            return true;
        };
        return true;
    }
    template <int L1>
    static bool make_flow_l1(b200fft_plan& pl, uint32_t L2, uint32_t lgN, const C* tw, FlowFn& fn, uint32_t& W) {
        const bool sw = pl.direction != 0;
        if (L2 == (uint32_t)L1) return sw ? make_flow_t<L1, L1, true>(pl, lgN, tw, fn, W) : make_flow_t<L1, L1, false>(pl, lgN, tw, fn, W);
        if constexpr (2 * L1 <= (int)TILE_MAX) {
            if (L2 == 2u * L1)
                return sw ? make_flow_t<L1, 2 * L1, true>(pl, lgN, tw, fn, W) : make_flow_t<L1, 2 * L1, false>(pl, lgN, tw, fn, W);
        }
        return false;
    }
    static bool make_flow_rt(b200fft_plan& pl, uint32_t L1, uint32_t L2, uint32_t lgN, const C* tw, FlowFn& fn, uint32_t& W) {
        switch (L1) {
            case 128: return make_flow_l1<128>(pl, L2, lgN, tw, fn, W);
            case 256: return make_flow_l1<256>(pl, L2, lgN, tw, fn, W);
            case 512: return make_flow_l1<512>(pl, L2, lgN, tw, fn, W);
            case 1024: return make_flow_l1<1024>(pl, L2, lgN, tw, fn, W);
            case 2048: return make_flow_l1<2048>(pl, L2, lgN, tw, fn, W);
            case 4096: return make_flow_l1<4096>(pl, L2, lgN, tw, fn, W);
        }
        return false;
    }
    // ---- fused single-launch variant (fused.h: run_fused) ----
    template <int L1, int L2, bool SW, bool TILED = true, bool DOUT = false>
    static bool make_fused_t(b200fft_plan& pl, uint32_t lgN, const C* full_tw, FusedFn& fn, uint32_t& W_out) {
        if constexpr (sizeof(T) == 4) {
            using GA = typename FusedGeo<T, L1>::type;
            using GB = typename FusedGeo<T, L2>::type;
            if constexpr (TILED) {
                if (!fused_tiled()) return make_fused_t<L1, L2, SW, false>(pl, lgN, full_tw, fn, W_out);
            }
            if constexpr (!TILED && !DOUT) {
                if (fused_bdirect(lgN)) return make_fused_t<L1, L2, SW, false, true>(pl, lgN, full_tw, fn, W_out);
            }
            // TILED: tile-major ring -- pass A stores straight from its registers, pass B gathers its rows with one 4-D tensor copy
            using KA = TmaTileKernel<GA, FF, FF, 0, SW, TILED ? 1 : 0>;
            using KB = TmaTileKernel<GB, JF, FF, 1, SW, TILED ? GA::F : 0, DOUT>;
            using FK = FusedKernel<KA, KB, FUSED_NG, FUSED_NS>;
            static_assert(FK::SMEM_BYTES <= MAX_SMEM_PER_CTA, "fused stages must fit one SM");
            const uint32_t lg1 = hm::ilog2(L1), lg2 = hm::ilog2(L2);
            const C* twa = upload(pl, stage_twiddles<GA>());
            const C* twb = upload(pl, stage_twiddles<GB>());
            if (!twa || !twb) return false;
            // optional two-level inter-pass twiddles (B200FFT_FUSED_TW2=1)
            const C *tw_lo = nullptr, *tw_hi = nullptr;
            if (fused_tw2() && KB::PRE) {
                const uint64_t Nn = 1ull << lgN;
                std::vector<C> lo(1024), hi((size_t)(Nn >> 10));
                for (uint64_t i = 0; i < 1024; ++i) lo[(size_t)i] = hm::twiddle<T>(i, Nn);
                for (uint64_t h = 0; h < (Nn >> 10); ++h) hi[(size_t)h] = hm::twiddle<T>(h, Nn >> 10);
                tw_lo = upload(pl, lo);
                tw_hi = upload(pl, hi);
                if (!tw_lo || !tw_hi) return false;
            }
            if (rt::fused_grid<KA, KB, FUSED_NG, FUSED_NS>() <= 0) return false;
            const uint32_t TA = (uint32_t)L2 / GA::F, TB = (uint32_t)L1 / GB::F;
            const uint64_t N = 1ull << lgN;
            const uint32_t W = fused_ring_slots(TA + TB, N * sizeof(C));
            W_out = W;
            const uint64_t ctl_bytes = flow_ctl_bytes(W);
            fn = [=](const void* in_v, void* out_v, void* work, uint64_t batch, rt::stream_t s) {
                const C* in = (const C*)in_v;
                C* out = (C*)out_v;
                // segments keep the ticket count inside 32 bits and the slab index of the tensor maps inside 31
                const uint64_t seg = std::max<uint64_t>(1, std::min<uint64_t>(1ull << 24, ((1ull << 30) / (TA + TB))));
                C* ring = (C*)((char*)work + ctl_bytes);
                for (uint64_t b0 = 0; b0 < batch; b0 += seg) {
                    const uint64_t nb = std::min(seg, batch - b0);
                    typename FK::Params p;
                    std::memset(&p, 0, sizeof(p));
                    if (!rt::make_tile_map(&p.a.map_in, false, in + b0 * N, L2, L1, nb, GA::F, KA::BOX_ROWS) ||
                        !rt::make_tile_map(&p.b.map_out, false, out + b0 * N, L1, L2, nb, GB::F, KB::BOX_ROWS))
                        return false;
                    if (TILED ? !rt::make_ring_map(&p.b.map_in, ring, GA::F, L1, TA, W, GB::F, KB::TBOX)
                              : !rt::make_tile_map(&p.a.map_out, false, ring, L2, L1, W, GA::F, KA::BOX_ROWS))
                        return false;
                    p.a.in = in + b0 * N;
                    p.a.out = ring;
                    p.a.tw = twa;
                    p.a.n_fft = nb << lg2;
                    p.a.lgN = lgN;
                    p.a.lg_other = lg2;
                    p.a.ring_w = W;
                    p.b.in = ring;
                    p.b.out = out + b0 * N;
                    p.b.tw = twb;
                    p.b.full_tw = full_tw;
                    p.b.tw_lo = tw_lo;
                    p.b.tw_hi = tw_hi;
                    p.b.n_fft = nb << lg1;
                    p.b.lgN = lgN;
                    p.b.lg_other = lg1;
                    p.b.discard = use_discard() ? 1u : 0u;
                    p.b.ring_w = W;
                    p.ctl = (uint32_t*)work;
                    p.flags = fused_flags();
                    if (fused_trace()) {  // the stamps live behind the ring (tools/fused_trace.py reads them back)
                        p.trace = (unsigned long long*)((char*)work + ctl_bytes + (uint64_t)W * N * sizeof(C));
                        if (!rt::memset_async(p.trace, 0, fused_trace_bytes(), s)) return false;
                    }
                    if (!make_flow_sched(p.sched, nb, TA, TB, W)) {
                        rt::g_err = "fused schedule overflow";
                        return false;
                    }
                    if (!rt::launch_fused<KA, KB, FUSED_NG, FUSED_NS>(p, ctl_bytes, s)) return false;
                }
                return true;
            };
            return true;
        } else {
            (void)pl; (void)lgN; (void)full_tw; (void)fn; (void)W_out;
            return false;
        }
    }
    template <int L1>
    static bool make_fused_l1(b200fft_plan& pl, uint32_t L2, uint32_t lgN, const C* tw, FusedFn& fn, uint32_t& W) {
        const bool sw = pl.direction != 0;
        if (L2 == (uint32_t)L1) return sw ? make_fused_t<L1, L1, true>(pl, lgN, tw, fn, W) : make_fused_t<L1, L1, false>(pl, lgN, tw, fn, W);
        if constexpr (2 * L1 <= (int)TILE_MAX) {
            if (L2 == 2u * L1)
                return sw ? make_fused_t<L1, 2 * L1, true>(pl, lgN, tw, fn, W) : make_fused_t<L1, 2 * L1, false>(pl, lgN, tw, fn, W);
        }
        return false;
    }
    static bool make_fused_rt(b200fft_plan& pl, uint32_t L1, uint32_t L2, uint32_t lgN, const C* tw, FusedFn& fn, uint32_t& W) {
        if constexpr (sizeof(T) == 4) return build_fused_f32(pl, L1, L2, lgN, tw, fn, W);
        return false;
    }
    static bool fused_build_here(b200fft_plan& pl, uint32_t L1, uint32_t L2, uint32_t lgN, const C* tw, FusedFn& fn, uint32_t& W) {
        switch (L1) {
            case 128: return make_fused_l1<128>(pl, L2, lgN, tw, fn, W);
            case 256: return make_fused_l1<256>(pl, L2, lgN, tw, fn, W);
            case 512: return make_fused_l1<512>(pl, L2, lgN, tw, fn, W);
            case 1024: return make_fused_l1<1024>(pl, L2, lgN, tw, fn, W);
            case 2048: return make_fused_l1<2048>(pl, L2, lgN, tw, fn, W);
            case 4096: return make_fused_l1<4096>(pl, L2, lgN, tw, fn, W);
        }
        return false;
    }
    // ---------------- single-pass four-step on a thread-block cluster (cluster.h), f32, N = 2^14 .. 2^17 ----------------
    // B200FFT_CLUSTER: bit mask of log2 N - 14 for which the cluster plan is the default (measured, profiles/); 0 = never
    static uint32_t cluster_mask() {
        static uint32_t v = [] {
            const char* e = std::getenv("B200FFT_CLUSTER");
            return e ? (uint32_t)std::atoi(e) : 0u;
        }();
        return v;
    }
    template <int L1, int L2, int CC, bool SW, bool HALF = false>
    static bool make_cluster_t(b200fft_plan& pl) {
        if constexpr (sizeof(T) == 4) {
            using GA = typename std::conditional<HALF, typename ClusterHalfGeo<L1>::type, typename FusedGeo<T, L1>::type>::type;
            using GB = typename std::conditional<HALF, typename ClusterHalfGeo<L2>::type, typename FusedGeo<T, L2>::type>::type;
            using KT = ClusterKernel<GA, GB, CC, SW>;
            static_assert(KT::SMEM_BYTES <= MAX_SMEM_PER_CTA, "cluster tile fits one CTA");
            const uint32_t lg1 = hm::ilog2(L1), lg2 = hm::ilog2(L2), lgN = lg1 + lg2;
            const C* twa = upload(pl, stage_twiddles<GA>());
            const C* twb = upload(pl, stage_twiddles<GB>());
            const C* full_tw = make_full_twiddles(pl, lg1, lg2);
            if (!twa || !twb || !full_tw) return false;
            if (rt::cluster_max_active<KT>() <= 0) return false;
            pl.exec = [=](const ExecCtx& c) {
                // (2^24 transforms per launch keep the CTA index inside 31 bits)
                const uint64_t seg = (1ull << 30) / CC;
                for (uint64_t b0 = 0; b0 < c.batch; b0 += seg) {
                    const uint64_t nb = std::min(seg, c.batch - b0);
                    typename KT::Params p;
                    p.load = LoadCols<T, SW>{(const C*)c.in + (b0 << lgN), lgN, lg2};
                    p.store = StoreTransposed<T, SW>{(C*)c.out + (b0 << lgN), lgN, lg1};
                    p.twa = twa;
                    p.twb = twb;
                    p.full_tw = full_tw;
                    p.n_transforms = nb;
                    if (!rt::launch_cluster<KT>(p, nb, c.stream)) return false;
                }
                return true;
            };
            pl.launches = [=](uint64_t batch) { return (batch + ((1ull << 30) / CC) - 1) / ((1ull << 30) / CC); };
            pl.desc = "ClusterFourStep{" + std::to_string(L1) + "x" + std::to_string(L2) + ",cluster=" + std::to_string(CC) + "}";
            set_recipe(pl, B200FFT_RECIPE_CLUSTER, HALF ? 1 : 0);
            return true;
        } else {
            (void)pl;
            return false;
        }
    }
    // Rader (n = M + 1 prime) / Bluestein (2n - 1 <= M) with the inner FFT of M = L * L points inside one cluster pass (cluster.h)
    template <int L, int CC, bool SW, int MODE>
    static bool make_cluster_conv_t(b200fft_plan& pl) {
        if constexpr (sizeof(T) == 4) {
            using G = typename FusedGeo<T, L>::type;
            using KT = ClusterConvKernel<G, CC, SW, MODE>;
            static_assert(KT::SMEM_BYTES <= MAX_SMEM_PER_CTA, "cluster tile fits one CTA");
            const uint64_t n = pl.len, M = (uint64_t)L * L;
            const uint32_t lg = hm::ilog2(L);
            std::vector<C> mult;
            uint64_t groot = 0;
            const uint32_t *d_g = nullptr, *d_s = nullptr;
            const C* d_chirp = nullptr;
            if (MODE == 0) {
                std::vector<uint32_t> gpow, ginv;
                rader_tables(n, M, gpow, ginv, mult, groot);
                d_g = upload(pl, gpow);
                d_s = upload(pl, ginv);
                if (!d_g || !d_s) return false;
            } else {
                std::vector<C> chirp;
                bluestein_tables(n, M, chirp, mult);
                d_chirp = upload(pl, chirp);
                if (!d_chirp) return false;
            }
            const C* d_mult = upload(pl, mult);
            const C* tw = upload(pl, stage_twiddles<G>());
            const C* full_tw = make_full_twiddles(pl, lg, lg);
            if (!d_mult || !tw || !full_tw) return false;
            if (rt::cluster_max_active<KT>() <= 0) return false;
            pl.exec = [=](const ExecCtx& c) {
                const uint64_t seg = (1ull << 30) / CC;
                for (uint64_t b0 = 0; b0 < c.batch; b0 += seg) {
                    const uint64_t nb = std::min(seg, c.batch - b0);
                    typename KT::Params p;
                    p.in = (const C*)c.in + b0 * n;
                    p.out = (C*)c.out + b0 * n;
                    p.gather = d_g;
                    p.scatter = d_s;
                    p.chirp = d_chirp;
                    p.mult = d_mult;
                    p.tw = tw;
                    p.full_tw = full_tw;
                    p.n = (uint32_t)n;
                    p.n_transforms = nb;
                    if (!rt::launch_cluster<KT>(p, nb, c.stream)) return false;
                }
                return true;
            };
            pl.launches = [=](uint64_t batch) { return (batch + ((1ull << 30) / CC) - 1) / ((1ull << 30) / CC); };
            const std::string inner = "ClusterFourStep{" + std::to_string(L) + "x" + std::to_string(L) + ",cluster=" + std::to_string(CC) + "}";
            pl.desc = MODE == 0 ? "Rader{n=" + std::to_string(n) + ",g=" + std::to_string(groot) + ",inner=" + inner + ",fused}"
                                : "Bluestein{n=" + std::to_string(n) + ",M=" + std::to_string(M) + ",inner=" + inner + ",fused}";
            set_recipe(pl, MODE == 0 ? B200FFT_RECIPE_RADER : B200FFT_RECIPE_BLUESTEIN, MODE == 0 ? 1 : 0, 0, B200FFT_RECIPE_CLUSTER, M);
            return true;
        } else {
            (void)pl;
            return false;
        }
    }
    // mode 0: Rader (pl.len = M + 1 prime), 1: Bluestein; M = 2^14 or 2^16
    static bool cluster_conv_build_here(b200fft_plan& pl, uint64_t M, int mode) {
        const bool sw = pl.direction != 0;
        if (M == (1u << 14)) {
            if (mode == 0) return sw ? make_cluster_conv_t<128, 2, true, 0>(pl) : make_cluster_conv_t<128, 2, false, 0>(pl);
            return sw ? make_cluster_conv_t<128, 2, true, 1>(pl) : make_cluster_conv_t<128, 2, false, 1>(pl);
        }
        if (M == (1u << 16)) {
            if (mode == 0) return sw ? make_cluster_conv_t<256, 8, true, 0>(pl) : make_cluster_conv_t<256, 8, false, 0>(pl);
            return sw ? make_cluster_conv_t<256, 8, true, 1>(pl) : make_cluster_conv_t<256, 8, false, 1>(pl);
        }
        return false;
    }
    static bool make_cluster_conv(b200fft_plan& pl, uint64_t M, int mode) {
        if constexpr (sizeof(T) == 4) return build_cluster_conv_f32(pl, M, mode);
        return false;
    }
    // B200FFT_CLUSTER_CONV=1: Rader 65537 and Bluestein with M = 2^14 / 2^16 (f32) run inside one cluster pass by default
    static bool use_cluster_conv() {
        static bool v = [] {
            const char* e = std::getenv("B200FFT_CLUSTER_CONV");
            return e && std::atoi(e) == 1;
        }();
        return v;
    }
    static bool cluster_build_here(b200fft_plan& pl, uint32_t lgN, bool half) {
        const bool sw = pl.direction != 0;
        if (half) {  // 4096-point tiles, 128 threads: four CTAs per SM, twice the cluster size
            switch (lgN) {
                case 14: return sw ? make_cluster_t<128, 128, 4, true, true>(pl) : make_cluster_t<128, 128, 4, false, true>(pl);
                case 15: return sw ? make_cluster_t<128, 256, 8, true, true>(pl) : make_cluster_t<128, 256, 8, false, true>(pl);
                case 16: return sw ? make_cluster_t<256, 256, 16, true, true>(pl) : make_cluster_t<256, 256, 16, false, true>(pl);
            }
            return false;
        }
        switch (lgN) {
            case 14: return sw ? make_cluster_t<128, 128, 2, true>(pl) : make_cluster_t<128, 128, 2, false>(pl);
            case 15: return sw ? make_cluster_t<128, 256, 4, true>(pl) : make_cluster_t<128, 256, 4, false>(pl);
            case 16: return sw ? make_cluster_t<256, 256, 8, true>(pl) : make_cluster_t<256, 256, 8, false>(pl);
            case 17: return sw ? make_cluster_t<256, 512, 16, true>(pl) : make_cluster_t<256, 512, 16, false>(pl);
        }
        return false;
    }
    static bool make_cluster(b200fft_plan& pl, uint32_t lgN, bool half = false) {
        if constexpr (sizeof(T) == 4) return build_cluster_f32(pl, lgN, half);
        return false;
    }

    static bool make_four_step(b200fft_plan& pl, uint32_t lgN) {
        const uint32_t lg1 = lgN / 2, lg2 = lgN - lg1;  // N1 <= N2
        const uint32_t N1 = 1u << lg1, N2 = 1u << lg2;
        if (N1 < TILE_MIN || N2 > TILE_MAX) return false;
        const C* full_tw = make_full_twiddles(pl, lg1, lg2);
        if (!full_tw) return false;
        if (use_flow() && N1 >= 128) {
            FlowFn fn;
            uint32_t W = 0;
            if (!make_flow_rt(pl, N1, N2, lgN, full_tw, fn, W)) return false;
            const uint64_t Nn = 1ull << lgN;
            uint64_t wbytes = flow_ctl_bytes(W) + (uint64_t)W * Nn * sizeof(C);
#if defined(B2_FLOW_TRACE)
            wbytes += (uint64_t)FLOW_TRACE_CTAS * FLOW_TRACE_WORDS * 8;
#endif
            pl.work_bytes = [=](uint64_t) { return wbytes; };
            pl.launches = [=](uint64_t batch) { return (batch + (1ull << 24) - 1) >> 24; };
            pl.exec = [=](const ExecCtx& c) {
                rt::set_l2_window(c.work, wbytes);
                const bool ok = fn((const C*)c.in, (C*)c.out, c.work, c.batch, c.stream);
                rt::set_l2_window(nullptr, 0);
                return ok;
            };
            pl.desc = "FourStep{" + std::to_string(N1) + "x" + std::to_string(N2) + ",flow,ring=" + std::to_string(W) + "}";
            set_recipe(pl, B200FFT_RECIPE_POW2);
            pl.chunk = W;
            return true;
        }
        PassFns fns;
        const bool sw = pl.direction != 0;
        const bool ok_a = sw ? make_pass_a_rt<true>(pl, N1, lgN, lg2, fns) : make_pass_a_rt<false>(pl, N1, lgN, lg2, fns);
        const bool ok_b = sw ? make_pass_b_rt<true>(pl, N2, lgN, lg1, full_tw, fns) : make_pass_b_rt<false>(pl, N2, lgN, lg1, full_tw, fns);
        if (!ok_a || !ok_b) return false;
        const uint64_t N = 1ull << lgN;
        // default for f32: one launch of the fused kernel; the chunked launch pairs below stay as the path for buffers
        // TMA cannot address (not 16-byte aligned) and for f64
        FusedFn fused;
        uint32_t fused_w = 0;
        uint64_t fused_bytes = 0;
        if (sizeof(T) == 4 && use_fused() && N1 >= 128) {
            if (!make_fused_rt(pl, N1, N2, lgN, full_tw, fused, fused_w)) return false;
            fused_bytes = flow_ctl_bytes(fused_w) + (uint64_t)fused_w * N * sizeof(C) + (fused_trace() ? fused_trace_bytes() : 0);
        }
        const int K = overlap_streams(lgN >= 20 ? 3 : 4);
        // K chunks in flight share the L2 budget
        const uint64_t chunk = std::max<uint64_t>(1, pick_chunk(N * sizeof(C), fns) / (uint64_t)K);
        pl.work_bytes = [=](uint64_t batch) {
            const uint64_t per = std::min(batch, chunk) * N * sizeof(C);
            const uint64_t nchunks = (batch + chunk - 1) / chunk;
            return std::max(fused_bytes, per * std::min<uint64_t>((uint64_t)K, std::max<uint64_t>(nchunks, 1)));
        };
        pl.launches = [=](uint64_t batch) { return fused ? (uint64_t)1 : 2 * ((batch + chunk - 1) / chunk); };
        b200fft_plan* self = &pl;
        pl.exec = [=](const ExecCtx& c) {
            const C* in = (const C*)c.in;
            C* out = (C*)c.out;
            C* work = (C*)c.work;
            if (fused && c.batch < (1ull << 31) &&
                ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 && (reinterpret_cast<uintptr_t>(work) & 127u) == 0)
                return fused((const void*)in, (void*)out, c.work, c.batch, c.stream);
            const uint64_t nchunks = (c.batch + chunk - 1) / chunk;
            const int ns = (int)std::min<uint64_t>((uint64_t)(c.max_streams > 0 ? std::min(c.max_streams, K) : K), nchunks);  // streams used
            rt::stream_t st[4] = {c.stream, nullptr, nullptr, nullptr};
            rt::event_t ev_fork = nullptr;
            if (ns > 1) {
                ev_fork = rt::event_create();
                if (!ev_fork || !rt::event_record(ev_fork, c.stream)) return false;
                for (int k = 1; k < ns; ++k) {
                    st[k] = self->aux_get();
                    if (!st[k] || !rt::stream_wait(st[k], ev_fork)) return false;
                }
            }
            bool ok = true;
            // TMA-tiled passes: tensor maps over the caller's buffers and the workspaces, built per call
            const bool tma = use_tma_tiles() && fns.a_tma && fns.b_tma && c.batch < (1ull << 31) &&
                             ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(work)) & 15u) == 0;
            TMap m_in, m_out, m_ws[4];
            if (tma) {
                const bool f64 = sizeof(T) == 8;
                ok = rt::make_tile_map(&m_in, f64, in, N2, N1, c.batch, fns.f_a, fns.box_a) &&
                     rt::make_tile_map(&m_out, f64, out, N1, N2, c.batch, fns.f_b, fns.box_b);
                for (int k = 0; k < ns && ok; ++k)
                    ok = rt::make_tile_map(&m_ws[k], f64, work + (uint64_t)k * chunk * N, N2, N1, std::min(chunk, c.batch), fns.f_a, fns.box_a);
            }
            rt::set_l2_window(work, (size_t)ns * std::min(chunk, c.batch) * N * sizeof(C));
            uint64_t idx = 0;
            for (uint64_t b0 = 0; b0 < c.batch && ok; b0 += chunk, ++idx) {
                const uint64_t nb = std::min(chunk, c.batch - b0);
                const int k = (int)(idx % (uint64_t)ns);
                C* w = work + (uint64_t)k * chunk * N;
                if (tma) {
                    // opt-in: pass B asks L2 for the input of this stream's NEXT chunk (it starts right after this pass)
                    const uint64_t nxt = b0 + (uint64_t)ns * chunk;
                    const void* pf = (use_prefetch() && nxt < c.batch) ? (const void*)(in + nxt * N) : nullptr;
                    const uint64_t pf_total = pf ? std::min(chunk, c.batch - nxt) * N * sizeof(C) : 0;
                    ok = fns.a_tma(m_in, m_ws[k], in, w, (uint32_t)b0, nb, st[k]) &&
                         fns.b_tma(m_out, w, out, (uint32_t)b0, nb, pf, pf_total, st[k]);
                }
                else
                    ok = fns.a(in + b0 * N, w, nb, st[k]) && fns.b(w, out + b0 * N, nb, st[k]);
            }
            rt::set_l2_window(nullptr, 0);
            if (ns > 1) {
                for (int k = 1; k < ns; ++k) {
                    rt::event_t ev = rt::event_create();
                    ok = ev && rt::event_record(ev, st[k]) && rt::stream_wait(c.stream, ev) && ok;
                    if (ev) rt::event_destroy(ev);
                    self->aux_put(st[k]);
                }
                rt::event_destroy(ev_fork);
            }
            return ok;
        };
        pl.desc = "FourStep{" + std::to_string(N1) + "x" + std::to_string(N2) +
                  (fused ? ",fused,ring=" + std::to_string(fused_w) + (fused_bdirect(lgN) && !fused_tiled() ? ",bdirect" : "") : std::string()) + "}";
        set_recipe(pl, B200FFT_RECIPE_POW2);
        pl.chunk = fused ? fused_w : chunk;
        return true;
    }

    // ---------------- large convolution plans (Rader / Bluestein over a four-step inner FFT) -------
    //   A1: gather|chirp-pad columns -> N1-pt FFT -> twiddle            in   -> w1
    //   B1: rows -> N2-pt FFT -> * mult, conj (+ Rader DC)               w1   -> w2 (natural order)
    //   A2: columns -> N1-pt FFT -> twiddle (plain pass A, in place)     w2   -> w2
    //   B2: rows -> N2-pt FFT -> conj + scatter | conj * chirp           w2   -> out
    struct ConvTables {
        const uint32_t* gather = nullptr;   // Rader: g^(i+1) mod n
        const uint32_t* scatter = nullptr;  // Rader: g^-(i+1) mod n
        const C* chirp = nullptr;           // Bluestein: W_2n^(i^2)
        const C* mult = nullptr;            // M entries
        const C* full_tw = nullptr;         // inner four-step twiddles [N1][N2]
        uint32_t n = 0, lgM = 0, lg1 = 0, lg2 = 0;
        bool rader = false;
    };
    struct ConvFns {
        std::function<bool(const C* in, C* w1, uint64_t nb, rt::stream_t)> a1;
        std::function<bool(const C* w1, C* w2, const C* in, C* out, uint64_t nb, rt::stream_t)> b1;
        std::function<bool(const C* w2, C* out, uint64_t nb, rt::stream_t)> b2;
    };
    template <int L1, bool SW, int V = 0>
    static bool make_conv_a1(b200fft_plan& pl, const ConvTables& t, ConvFns& f) {
        if constexpr (V == 0 && HasV1<T, L1>::tile) {
            if (use_radix32()) return make_conv_a1<L1, SW, 1>(pl, t, f);
        }
        using G = typename TileGeo<T, L1, V>::type;
        using KT = FftKernel<G, FF, FF, LoadColsConv<T, SW>, StoreCols<T>>;
        const C* tw = upload(pl, stage_twiddles<G>());
        if (!tw) return false;
        f.a1 = [=](const C* in, C* w1, uint64_t nb, rt::stream_t s) {
            typename KT::Params p;
            p.load = LoadColsConv<T, SW>{in, t.gather, t.chirp, t.n, t.lg2};
            p.store = StoreCols<T>{w1, t.lgM, t.lg2};
            p.tw = tw;
            p.n_fft = nb << t.lg2;
            return rt::launch<KT>(p, (p.n_fft + G::F - 1) / G::F, s);
        };
        return true;
    }
    template <int L2, bool SW, int V = 0>
    static bool make_conv_b(b200fft_plan& pl, const ConvTables& t, ConvFns& f) {
        if constexpr (V == 0 && HasV1<T, L2>::tile) {
            if (use_radix32()) return make_conv_b<L2, SW, 1>(pl, t, f);
        }
        using G = typename TileGeo<T, L2, V>::type;
        using K0 = FftKernel<G, JF, FF, LoadRowsTw<T>, StoreTransposedConv<T, SW, 0>>;
        const C* tw = upload(pl, stage_twiddles<G>());
        if (!tw) return false;
        f.b1 = [=](const C* w1, C* w2, const C* in, C* out, uint64_t nb, rt::stream_t s) {
            typename K0::Params p;
            p.load = LoadRowsTw<T>{w1, t.full_tw, (uint32_t)G::L, t.lg1};
            p.store = StoreTransposedConv<T, SW, 0>{w2, t.mult, nullptr, nullptr, t.rader ? in : nullptr, out, t.n, t.lgM, t.lg1};
            p.tw = tw;
            p.n_fft = nb << t.lg1;
            return rt::launch<K0>(p, (p.n_fft + G::F - 1) / G::F, s);
        };
        if (t.rader) {
            using K1 = FftKernel<G, JF, FF, LoadRowsTw<T>, StoreTransposedConv<T, SW, 1>>;
            f.b2 = [=](const C* w2, C* out, uint64_t nb, rt::stream_t s) {
                typename K1::Params p;
                p.load = LoadRowsTw<T>{w2, t.full_tw, (uint32_t)G::L, t.lg1};
                p.store = StoreTransposedConv<T, SW, 1>{out, nullptr, t.scatter, nullptr, nullptr, nullptr, t.n, t.lgM, t.lg1};
                p.tw = tw;
                p.n_fft = nb << t.lg1;
                return rt::launch<K1>(p, (p.n_fft + G::F - 1) / G::F, s);
            };
        } else {
            using K2 = FftKernel<G, JF, FF, LoadRowsTw<T>, StoreTransposedConv<T, SW, 2>>;
            f.b2 = [=](const C* w2, C* out, uint64_t nb, rt::stream_t s) {
                typename K2::Params p;
                p.load = LoadRowsTw<T>{w2, t.full_tw, (uint32_t)G::L, t.lg1};
                p.store = StoreTransposedConv<T, SW, 2>{out, nullptr, nullptr, t.chirp, nullptr, nullptr, t.n, t.lgM, t.lg1};
                p.tw = tw;
                p.n_fft = nb << t.lg1;
                return rt::launch<K2>(p, (p.n_fft + G::F - 1) / G::F, s);
            };
        }
        return true;
    }
    template <bool SW>
    static bool make_conv_rt(b200fft_plan& pl, const ConvTables& t, ConvFns& f, PassFns& plain) {
        bool ok = false;
        switch (1u << t.lg1) {
            case 64: ok = make_conv_a1<64, SW>(pl, t, f) && make_pass_a<64, false>(pl, t.lgM, t.lg2, plain); break;
            case 128: ok = make_conv_a1<128, SW>(pl, t, f) && make_pass_a<128, false>(pl, t.lgM, t.lg2, plain); break;
            case 256: ok = make_conv_a1<256, SW>(pl, t, f) && make_pass_a<256, false>(pl, t.lgM, t.lg2, plain); break;
            case 512: ok = make_conv_a1<512, SW>(pl, t, f) && make_pass_a<512, false>(pl, t.lgM, t.lg2, plain); break;
            case 1024: ok = make_conv_a1<1024, SW>(pl, t, f) && make_pass_a<1024, false>(pl, t.lgM, t.lg2, plain); break;
            case 2048: ok = make_conv_a1<2048, SW>(pl, t, f) && make_pass_a<2048, false>(pl, t.lgM, t.lg2, plain); break;
            case 4096: ok = make_conv_a1<4096, SW>(pl, t, f) && make_pass_a<4096, false>(pl, t.lgM, t.lg2, plain); break;
        }
        if (!ok) return false;
        switch (1u << t.lg2) {
            case 64: return make_conv_b<64, SW>(pl, t, f);
            case 128: return make_conv_b<128, SW>(pl, t, f);
            case 256: return make_conv_b<256, SW>(pl, t, f);
            case 512: return make_conv_b<512, SW>(pl, t, f);
            case 1024: return make_conv_b<1024, SW>(pl, t, f);
            case 2048: return make_conv_b<2048, SW>(pl, t, f);
            case 4096: return make_conv_b<4096, SW>(pl, t, f);
        }
        return false;
    }
    // rader = true: n prime, M = n - 1;  rader = false: Bluestein, M = next_pow2(2n - 1)
    static bool make_big_conv(b200fft_plan& pl, uint64_t M, bool rader) {
        const uint64_t n = pl.len;
        ConvTables t;
        t.n = (uint32_t)n;
        t.lgM = hm::ilog2(M);
        t.lg1 = t.lgM / 2;
        t.lg2 = t.lgM - t.lg1;
        t.rader = rader;
        if ((1u << t.lg1) < TILE_MIN || (1u << t.lg2) > TILE_MAX) return false;
        std::vector<C> mult;
        uint64_t groot = 0;
        if (rader) {
            std::vector<uint32_t> gpow, ginv;
            rader_tables(n, M, gpow, ginv, mult, groot);
            t.gather = upload(pl, gpow);
            t.scatter = upload(pl, ginv);
            if (!t.gather || !t.scatter) return false;
        } else {
            std::vector<C> chirp;
            bluestein_tables(n, M, chirp, mult);
            t.chirp = upload(pl, chirp);
            if (!t.chirp) return false;
        }
        t.mult = upload(pl, mult);
        if (!t.mult) return false;
        t.full_tw = make_full_twiddles(pl, t.lg1, t.lg2);
        if (!t.full_tw) return false;
        ConvFns f;
        PassFns plain;
        const bool ok = pl.direction ? make_conv_rt<true>(pl, t, f, plain) : make_conv_rt<false>(pl, t, f, plain);
        if (!ok) return false;
        const int K = overlap_streams();
        // two workspaces per chunk, K chunks in flight, all inside the L2 budget
        const uint64_t chunk = std::max<uint64_t>(1, chunk_bytes() / 2 / (uint64_t)K / (M * sizeof(C)));
        pl.work_bytes = [=](uint64_t batch) {
            const uint64_t nchunks = (batch + chunk - 1) / chunk;
            return 2 * std::min(batch, chunk) * M * sizeof(C) * std::min<uint64_t>((uint64_t)K, std::max<uint64_t>(nchunks, 1));
        };
        pl.launches = [=](uint64_t batch) { return 4 * ((batch + chunk - 1) / chunk); };
        b200fft_plan* self = &pl;
        pl.exec = [=](const ExecCtx& c) {
            const C* in = (const C*)c.in;
            C* out = (C*)c.out;
            const uint64_t nchunks = (c.batch + chunk - 1) / chunk;
            const int ns = (int)std::min<uint64_t>((uint64_t)(c.max_streams > 0 ? std::min(c.max_streams, K) : K), nchunks);
            const uint64_t per = std::min(c.batch, chunk) * M;  // elements of one workspace
            rt::stream_t st[4] = {c.stream, nullptr, nullptr, nullptr};
            rt::event_t ev_fork = nullptr;
            if (ns > 1) {
                ev_fork = rt::event_create();
                if (!ev_fork || !rt::event_record(ev_fork, c.stream)) return false;
                for (int k = 1; k < ns; ++k) {
                    st[k] = self->aux_get();
                    if (!st[k] || !rt::stream_wait(st[k], ev_fork)) return false;
                }
            }
            bool ok = true;
            uint64_t idx = 0;
            for (uint64_t b0 = 0; b0 < c.batch && ok; b0 += chunk, ++idx) {
                const uint64_t nb = std::min(chunk, c.batch - b0);
                const int k = (int)(idx % (uint64_t)ns);
                C* w1 = (C*)c.work + (uint64_t)k * 2 * per;
                C* w2 = w1 + per;
                ok = f.a1(in + b0 * n, w1, nb, st[k]) && f.b1(w1, w2, in + b0 * n, out + b0 * n, nb, st[k]) &&
                     plain.a(w2, w2, nb, st[k]) && f.b2(w2, out + b0 * n, nb, st[k]);
            }
            if (ns > 1) {
                for (int k = 1; k < ns; ++k) {
                    rt::event_t ev = rt::event_create();
                    ok = ev && rt::event_record(ev, st[k]) && rt::stream_wait(c.stream, ev) && ok;
                    if (ev) rt::event_destroy(ev);
                    self->aux_put(st[k]);
                }
                rt::event_destroy(ev_fork);
            }
            return ok;
        };
        const std::string inner = "FourStep{" + std::to_string(1u << t.lg1) + "x" + std::to_string(1u << t.lg2) + "}";
        pl.desc = rader ? "Rader{n=" + std::to_string(n) + ",g=" + std::to_string(groot) + ",inner=" + inner + "}"
                        : "Bluestein{n=" + std::to_string(n) + ",M=" + std::to_string(M) + ",inner=" + inner + "}";
        set_recipe(pl, rader ? B200FFT_RECIPE_RADER : B200FFT_RECIPE_BLUESTEIN, rader ? 1 : 0, 0, B200FFT_RECIPE_POW2, M);
        return true;
    }

    // ---------------- Smooth (7-smooth lengths without a compiled geometry) ----------------
    static constexpr uint32_t SMOOTH_MAX = sizeof(T) == 4 ? 4096 : 2048;  // 2 * n * sizeof(C) <= 64 KiB
    static bool smooth_factor(uint64_t n, std::vector<uint32_t>& radices) {
        uint32_t a = 0, b = 0, c = 0, d = 0;
        // the odd-prime butterflies the reference also hard-codes (src/plan.rs:609-634), largest first
        for (uint32_t p : {31u, 29u, 23u, 19u, 17u, 13u, 11u})
            while (n % p == 0) { n /= p; radices.push_back(p); }
        while (n % 2 == 0) { n /= 2; ++a; }
        while (n % 3 == 0) { n /= 3; ++b; }
        while (n % 5 == 0) { n /= 5; ++c; }
        while (n % 7 == 0) { n /= 7; ++d; }
        if (n != 1) return false;
        // odd radices FIRST: stage 0 scatters its outputs with a stride of R0 elements, and an odd stride is
        // free of shared-memory bank conflicts (a radix-16 first stage would be a 16-way conflict); then
        // the powers of two, largest first (fewest passes)
        for (uint32_t i = 0; i < d; ++i) radices.push_back(7);
        for (uint32_t i = 0; i < c; ++i) radices.push_back(5);
        for (uint32_t i = 0; i < b; ++i) radices.push_back(3);
        while (a >= 4) { radices.push_back(16); a -= 4; }
        if (a == 3) radices.push_back(8);
        if (a == 2) radices.push_back(4);
        if (a == 1) radices.push_back(2);
        return !radices.empty() && radices.size() <= 8;
    }
    static uint32_t max_radix(const std::vector<uint32_t>& r) {
        uint32_t m = 0;
        for (uint32_t v : r) m = std::max(m, v);
        return m;
    }
    template <bool SW, int RMAX = 31>
    static bool make_smooth_t(b200fft_plan& pl, const std::vector<uint32_t>& radices) {
        using KT = SmoothKernel<T, SW, RMAX>;
        const uint32_t n = (uint32_t)pl.len;
        typename KT::Params base;
        std::memset(&base, 0, sizeof(base));
        std::vector<C> tw;
        uint32_t p = 1;
        for (size_t s = 0; s < radices.size(); ++s) {
            const uint32_t R = radices[s];
            base.radix[s] = R;
            base.tw_off[s] = (uint32_t)tw.size();
            base.div_t[s] = make_fastdiv(n / R);
            base.div_p[s] = make_fastdiv(p);
            if (s >= 1)
                for (uint32_t r = 1; r < R; ++r)
                    for (uint32_t k = 0; k < p; ++k) tw.push_back(hm::twiddle<T>((uint64_t)k * r, (uint64_t)p * R));
            p *= R;
        }
        if (tw.empty()) tw.push_back(mk<T>(1, 0));
        const C* d_tw = upload(pl, tw);
        if (!d_tw) return false;
        base.tw = d_tw;
        base.n = n;
        base.n_stages = (uint32_t)radices.size();
        // transforms per CTA: fill 256 threads with butterflies of the smallest stage, stay inside 64 KiB
        uint32_t F = std::max<uint32_t>(1, SMOOTH_MAX / n);
        if (F > 64) F = 64;
        base.f_per_cta = F;
        base.smem_bytes = radices.size() > 1 ? (uint32_t)(2ull * F * n * sizeof(C)) : 0;
        const size_t max_smem = 2ull * SMOOTH_MAX * sizeof(C);
        pl.exec = [=](const ExecCtx& c) {
            typename KT::Params q = base;
            q.in = (const C*)c.in;
            q.out = (C*)c.out;
            q.n_fft = c.batch;
            return rt::launch_dyn<KT>(q, (c.batch + F - 1) / F, q.smem_bytes, max_smem, c.stream);
        };
        pl.launches = [](uint64_t) { return (uint64_t)1; };
        std::string rs;
        for (size_t s = 0; s < radices.size(); ++s) rs += (s ? "x" : "") + std::to_string(radices[s]);
        pl.desc = "Smooth{" + std::to_string(n) + "=" + rs + "}";
        set_recipe(pl, B200FFT_RECIPE_SMOOTH);
        return true;
    }

    // ---------------- SmoothFourStep: composite N = N1 * N2 above SMOOTH_MAX, every prime factor <= 31 ----------------
    // (the reference plans such lengths as MixedRadix / GoodThomas trees, src/plan.rs:508-607; here: two passes, the
    // intermediate in an L2-sized workspace, chunked over the batch like the power-of-two FourStep)
    static bool smooth_split(uint64_t n, uint32_t& n1, uint32_t& n2) {
        // N1 <= N2 <= SMOOTH_MAX, N1 as close to sqrt(N) as possible, both factorable into <= 8 stages
        uint64_t best = 0;
        for (uint64_t a = 2; a * a <= n; ++a) {
            if (n % a) continue;
            const uint64_t b = n / a;
            if (b > SMOOTH_MAX) continue;
            std::vector<uint32_t> ra, rb;
            if (smooth_factor(a, ra) && smooth_factor(b, rb)) best = a;
        }
        if (!best) return false;
        n1 = (uint32_t)best;
        n2 = (uint32_t)(n / best);
        return true;
    }
    template <class KT>
    static bool fill_smooth_pass(b200fft_plan& pl, typename KT::Params& base, uint32_t n, const std::vector<uint32_t>& radices) {
        std::memset(&base, 0, sizeof(base));
        std::vector<C> tw;
        uint32_t p = 1;
        for (size_t s = 0; s < radices.size(); ++s) {
            const uint32_t R = radices[s];
            base.radix[s] = R;
            base.tw_off[s] = (uint32_t)tw.size();
            base.div_t[s] = make_fastdiv(n / R);
            base.div_p[s] = make_fastdiv(p);
            if (s >= 1)
                for (uint32_t r = 1; r < R; ++r)
                    for (uint32_t k = 0; k < p; ++k) tw.push_back(hm::twiddle<T>((uint64_t)k * r, (uint64_t)p * R));
            p *= R;
        }
        if (tw.empty()) tw.push_back(mk<T>(1, 0));
        base.tw = upload(pl, tw);
        base.n = n;
        base.n_stages = (uint32_t)radices.size();
        return base.tw != nullptr;
    }
    // geometry of the two passes of a smooth N1 x N2 split (shared by SmoothFourStep, GoodThomas and the large convolution plans)
    template <class KA, class KB>
    static bool smooth_pass_pair(b200fft_plan& pl, uint32_t N1, uint32_t N2, typename KA::Params& pa, typename KB::Params& pb, uint32_t& FA,
                                 uint32_t& FB) {
        const uint64_t N = (uint64_t)N1 * N2;
        std::vector<uint32_t> ra, rb;
        if (!smooth_factor(N1, ra) || !smooth_factor(N2, rb)) return false;
        if (!fill_smooth_pass<KA>(pl, pa, N1, ra) || !fill_smooth_pass<KB>(pl, pb, N2, rb)) return false;
        pa.NN = pb.NN = N;
        pa.other = N2;
        pb.other = N1;
        pa.div_other = make_fastdiv(N2);
        pb.div_other = make_fastdiv(N1);
        // FFTs per CTA: both ping-pong buffers inside 2 * SMOOTH_MAX elements (+ the odd pitch of pass B)
        FA = std::max<uint32_t>(1, std::min<uint32_t>(64, SMOOTH_MAX / N1));
        const uint32_t pitch = N2 | 1u;
        FB = std::max<uint32_t>(1, std::min<uint32_t>(64, SMOOTH_MAX / pitch));
        pa.f_per_cta = FA;
        pa.pitch = N1;
        pa.div_f = make_fastdiv(FA);
        pa.smem_bytes = ra.size() > 1 ? (uint32_t)(2ull * FA * N1 * sizeof(C)) : 0;
        pb.f_per_cta = FB;
        pb.pitch = pitch;
        pb.div_f = make_fastdiv(FB);
        pb.smem_bytes = rb.size() > 1 ? (uint32_t)(2ull * FB * pitch * sizeof(C)) : 0;
        return true;
    }
    static const C* smooth_full_twiddles(b200fft_plan& pl, uint32_t N1, uint32_t N2) {
        // inter-pass twiddles W_N^(k1 n2), [k1][n2], each entry rounded once
        const uint64_t N = (uint64_t)N1 * N2;
        std::vector<C> t((size_t)N);
        for (uint64_t k1 = 0; k1 < N1; ++k1)
            for (uint64_t n2 = 0; n2 < N2; ++n2) t[(size_t)(k1 * N2 + n2)] = hm::twiddle<T>(k1 * n2, N);
        return upload(pl, t);
    }
    // variant 0: SmoothFourStep (the reference's MixedRadix);  1: GoodThomas (N1, N2 coprime: CRT / Ruritanian index maps, no twiddles)
    static bool small_radices_only(uint32_t N1, uint32_t N2) {
        std::vector<uint32_t> ra, rb;
        return smooth_factor(N1, ra) && smooth_factor(N2, rb) && max_radix(ra) <= 16 && max_radix(rb) <= 16;
    }
    template <bool SW, int RMAX = 31>
    static bool make_smooth_four_step_t(b200fft_plan& pl, uint32_t N1, uint32_t N2, int variant) {
        using KA = SmoothPassKernel<T, SW, 1, RMAX>;
        using KB = SmoothPassKernel<T, SW, 2, RMAX>;
        const uint64_t N = (uint64_t)N1 * N2;
        typename KA::Params pa;
        typename KB::Params pb;
        uint32_t FA = 0, FB = 0;
        if (!smooth_pass_pair<KA, KB>(pl, N1, N2, pa, pb, FA, FB)) return false;
        if (variant == 1) {
            // n = crt1[n1] + crt2[n2] mod N is the index with n mod N1 = n1 and n mod N2 = n2
            const uint64_t e1 = (uint64_t)N2 * hm::invmod(N2 % N1, N1) % N, e2 = (uint64_t)N1 * hm::invmod(N1 % N2, N2) % N;
            std::vector<uint32_t> c1(N1), c2(N2);
            for (uint64_t i = 0; i < N1; ++i) c1[(size_t)i] = (uint32_t)(hm::mulmod(i, e1, N));
            for (uint64_t i = 0; i < N2; ++i) c2[(size_t)i] = (uint32_t)(hm::mulmod(i, e2, N));
            pa.crt1 = upload(pl, c1);
            pa.crt2 = upload(pl, c2);
            if (!pa.crt1 || !pa.crt2) return false;
            pa.gt = pb.gt = 1;
        } else {
            pb.full_tw = smooth_full_twiddles(pl, N1, N2);
            if (!pb.full_tw) return false;
        }
        const size_t max_smem = 2ull * (SMOOTH_MAX + 64) * sizeof(C);
        // transforms per chunk: ~32 MiB of intermediate, and FFT indices of a launch must stay below 2^31
        uint64_t chunk = std::max<uint64_t>(1, (32ull << 20) / (N * sizeof(C)));
        chunk = std::min<uint64_t>(chunk, ((1ull << 31) - 1) / std::max(N1, N2));
        pl.chunk = chunk;
        pl.work_bytes = [=](uint64_t batch) { return std::min(batch, chunk) * N * sizeof(C); };
        pl.launches = [=](uint64_t batch) { return 2 * ((batch + chunk - 1) / chunk); };
        pl.exec = [=](const ExecCtx& c) {
            const C* in = (const C*)c.in;
            C* out = (C*)c.out;
            C* work = (C*)c.work;
            for (uint64_t b0 = 0; b0 < c.batch; b0 += chunk) {
                const uint64_t nb = std::min(chunk, c.batch - b0);
                typename KA::Params qa = pa;
                qa.in = in + b0 * N;
                qa.out = work;
                qa.n_fft = nb * N2;
                if (!rt::launch_dyn<KA>(qa, (qa.n_fft + FA - 1) / FA, qa.smem_bytes, max_smem, c.stream)) return false;
                typename KB::Params qb = pb;
                qb.in = work;
                qb.out = out + b0 * N;
                qb.n_fft = nb * N1;
                if (!rt::launch_dyn<KB>(qb, (qb.n_fft + FB - 1) / FB, qb.smem_bytes, max_smem, c.stream)) return false;
            }
            return true;
        };
        pl.desc = std::string(variant == 1 ? "GoodThomas{" : "SmoothFourStep{") + std::to_string(N1) + "x" + std::to_string(N2) + "}";
        set_recipe(pl, variant == 1 ? B200FFT_RECIPE_GOOD_THOMAS : B200FFT_RECIPE_MIXED_RADIX, N1, N2);
        return true;
    }
    // coprime split N = N1 * N2 for GoodThomas: both factors <= SMOOTH_MAX and smooth, as balanced as possible
    static bool coprime_split(uint64_t n, uint32_t& n1, uint32_t& n2) {
        uint64_t best = 0;
        for (uint64_t a = 2; a * a <= n; ++a) {
            if (n % a) continue;
            const uint64_t b = n / a;
            if (b > SMOOTH_MAX || hm::gcd(a, b) != 1) continue;
            std::vector<uint32_t> ra, rb;
            if (smooth_factor(a, ra) && smooth_factor(b, rb)) best = a;
        }
        if (!best) return false;
        n1 = (uint32_t)best;
        n2 = (uint32_t)(n / best);
        return true;
    }

    // ---------------- compiled two-pass plans of composite lengths (f32; SmoothTileGeo) ----------------
    // B200FFT_SMOOTH_COMPILED=0: the run-time-radix passes for these lengths too (A/B measurements)
    static bool use_compiled_smooth() {
        static bool v = [] {
            const char* e = std::getenv("B200FFT_SMOOTH_COMPILED");
            return !(e && std::atoi(e) == 0);
        }();
        return v;
    }
    static bool compiled_len(uint64_t x) {
        for (uint32_t v : COMPILED_TILE_LENGTHS)
            if (v == x) return true;
        return false;
    }
    static bool compiled_pair(uint64_t n, uint32_t& a, uint32_t& b) {
        if (sizeof(T) != 4 || !use_compiled_smooth() || hm::is_pow2(n)) return false;
        // the most balanced product of two compiled tile lengths
        bool found = false;
        for (uint32_t x : COMPILED_TILE_LENGTHS)
            for (uint32_t y : COMPILED_TILE_LENGTHS)
                if (x <= y && (uint64_t)x * y == n && (!found || x > a)) {
                    a = x;
                    b = y;
                    found = true;
                }
        return found;
    }
    // chunks of a multi-pass plan rotate over up to K streams (the caller's + auxiliary ones, fork / join with events); body(b0, nb, k,
    // stream) issues the launches of one chunk, k = which workspace
    template <class Body>
    static bool run_chunks(b200fft_plan* self, const ExecCtx& c, uint64_t chunk, int K, Body body) {
        const uint64_t nchunks = (c.batch + chunk - 1) / chunk;
        const int ns = (int)std::min<uint64_t>((uint64_t)(c.max_streams > 0 ? std::min(c.max_streams, K) : K), nchunks);
        rt::stream_t st[4] = {c.stream, nullptr, nullptr, nullptr};
        rt::event_t ev_fork = nullptr;
        if (ns > 1) {
            ev_fork = rt::event_create();
            if (!ev_fork || !rt::event_record(ev_fork, c.stream)) return false;
            for (int k = 1; k < ns; ++k) {
                st[k] = self->aux_get();
                if (!st[k] || !rt::stream_wait(st[k], ev_fork)) return false;
            }
        }
        bool ok = true;
        uint64_t idx = 0;
        for (uint64_t b0 = 0; b0 < c.batch && ok; b0 += chunk, ++idx) {
            const int k = (int)(idx % (uint64_t)ns);
            ok = body(b0, std::min(chunk, c.batch - b0), k, st[k]);
        }
        if (ns > 1) {
            for (int k = 1; k < ns; ++k) {
                rt::event_t ev = rt::event_create();
                ok = ev && rt::event_record(ev, st[k]) && rt::stream_wait(c.stream, ev) && ok;
                if (ev) rt::event_destroy(ev);
                self->aux_put(st[k]);
            }
            rt::event_destroy(ev_fork);
        }
        return ok;
    }
    // the two passes are built per LENGTH (pass A depends on N1 only, pass B on N2 only): any pair of compiled lengths combines
    typedef std::function<bool(const C* in, C* work, uint64_t N, uint32_t N2, uint64_t nb, rt::stream_t)> CPassA;
    typedef std::function<bool(const C* work, C* out, const C* full_tw, uint64_t N, uint32_t N1, uint64_t nb, rt::stream_t)> CPassB;
    template <int L, bool SW>
    static bool make_cpass_a(b200fft_plan& pl, CPassA& fn) {
        if constexpr (sizeof(T) == 4) {
            using G = typename SmoothTileGeo<L>::type;
            using KA = FftKernel<G, FF, FF, LoadColsG<T, SW>, StoreColsG<T>>;
            const C* tw = upload(pl, stage_twiddles<G>());
            if (!tw) return false;
            fn = [=](const C* in, C* work, uint64_t N, uint32_t N2, uint64_t nb, rt::stream_t s) {
                const FastDiv d2 = make_fastdiv(N2);
                typename KA::Params a;
                a.load = LoadColsG<T, SW>{in, N, N2, d2};
                a.store = StoreColsG<T>{work, N, N2, d2};
                a.tw = tw;
                a.n_fft = nb * N2;
                return rt::launch<KA>(a, (a.n_fft + G::F - 1) / G::F, s);
            };
            return true;
        } else {
            (void)pl; (void)fn;
            return false;
        }
    }
    template <int L, bool SW>
    static bool make_cpass_b(b200fft_plan& pl, CPassB& fn) {
        if constexpr (sizeof(T) == 4) {
            using G = typename SmoothTileGeo<L>::type;
            using KB = FftKernel<G, JF, FF, LoadRowsTwG<T>, StoreTransposedG<T, SW>>;
            const C* tw = upload(pl, stage_twiddles<G>());
            if (!tw) return false;
            fn = [=](const C* work, C* out, const C* full_tw, uint64_t N, uint32_t N1, uint64_t nb, rt::stream_t s) {
                const FastDiv d1 = make_fastdiv(N1);
                typename KB::Params b;
                b.load = LoadRowsTwG<T>{work, full_tw, (uint32_t)L, N1, d1, use_discard() ? 1u : 0u};
                b.store = StoreTransposedG<T, SW>{out, N, N1, d1};
                b.tw = tw;
                b.n_fft = nb * N1;
                return rt::launch<KB>(b, (b.n_fft + G::F - 1) / G::F, s);
            };
            return true;
        } else {
            (void)pl; (void)fn;
            return false;
        }
    }
    template <bool SW>
    static bool make_cpasses(b200fft_plan& pl, uint32_t a, uint32_t b, CPassA& fa, CPassB& fb) {
        bool ok = false;
#define B2_CL(L) case L: ok = make_cpass_a<L, SW>(pl, fa); break;
        switch (a) { B2_CL(64) B2_CL(100) B2_CL(125) B2_CL(128) B2_CL(196) B2_CL(200) B2_CL(225) B2_CL(250) B2_CL(256) B2_CL(375) B2_CL(400) B2_CL(500) B2_CL(512) B2_CL(625) B2_CL(1000) B2_CL(1024) }
#undef B2_CL
        if (!ok) return false;
        ok = false;
#define B2_CL(L) case L: ok = make_cpass_b<L, SW>(pl, fb); break;
        switch (b) { B2_CL(64) B2_CL(100) B2_CL(125) B2_CL(128) B2_CL(196) B2_CL(200) B2_CL(225) B2_CL(250) B2_CL(256) B2_CL(375) B2_CL(400) B2_CL(500) B2_CL(512) B2_CL(625) B2_CL(1000) B2_CL(1024) }
#undef B2_CL
        return ok;
    }
    static bool compiled_smooth_build_here(b200fft_plan& pl, uint32_t L1, uint32_t L2) {
        CPassA fa;
        CPassB fb;
        if (!(pl.direction ? make_cpasses<true>(pl, L1, L2, fa, fb) : make_cpasses<false>(pl, L1, L2, fa, fb))) return false;
        const uint64_t N = (uint64_t)L1 * L2;
        const C* full_tw = smooth_full_twiddles(pl, L1, L2);
        if (!full_tw) return false;
        // chunks in flight share ~64 MiB of L2; a chunk should still fill the device about twice (a launch of 125 CTAs on 148 SMs is
        // all ramp and tail), so long transforms use fewer streams; FFT indices of a launch stay below 2^31
        const uint64_t fit = std::max<uint64_t>(1, (64ull << 20) / (N * sizeof(C)));
        const int K = overlap_streams(fit >= 16 ? 4 : (fit >= 6 ? 2 : 1));
        uint64_t chunk = std::max<uint64_t>(1, fit / (uint64_t)K);
        chunk = std::min<uint64_t>(chunk, ((1ull << 31) - 1) / std::max(L1, L2));
        pl.chunk = chunk;
        pl.work_bytes = [=](uint64_t batch) {
            const uint64_t nchunks = (batch + chunk - 1) / chunk;
            return std::min(batch, chunk) * N * sizeof(C) * std::min<uint64_t>((uint64_t)K, std::max<uint64_t>(nchunks, 1));
        };
        pl.launches = [=](uint64_t batch) { return 2 * ((batch + chunk - 1) / chunk); };
        b200fft_plan* self = &pl;
        pl.exec = [=](const ExecCtx& c) {
            const C* in = (const C*)c.in;
            C* out = (C*)c.out;
            C* work = (C*)c.work;
            return run_chunks(self, c, chunk, K, [&](uint64_t b0, uint64_t nb, int k, rt::stream_t s) {
                C* w = work + (uint64_t)k * chunk * N;
                return fa(in + b0 * N, w, N, L2, nb, s) && fb(w, out + b0 * N, full_tw, N, L1, nb, s);
            });
        };
        pl.desc = "SmoothFourStep{" + std::to_string(L1) + "x" + std::to_string(L2) + ",compiled}";
        set_recipe(pl, B200FFT_RECIPE_MIXED_RADIX, L1, L2);
        return true;
    }
    static bool make_compiled_smooth(b200fft_plan& pl, uint32_t a, uint32_t b) {
        if constexpr (sizeof(T) == 4) return build_compiled_smooth_f32(pl, a, b);
        return false;
    }

    // ---------------- SmoothConv: Rader / Bluestein in ONE CTA pass over a smooth inner length (kernels.h) ----------------
    static constexpr uint32_t CONV_SMOOTH_MAX = sizeof(T) == 4 ? 6144 : 3072;  // 2 M sizeof(C) <= 96 KiB: two CTAs per SM
    static std::string radix_string(const std::vector<uint32_t>& r) {
        std::string rs;
        for (size_t s = 0; s < r.size(); ++s) rs += (s ? "x" : "") + std::to_string(r[s]);
        return rs;
    }
    // mode 0: Rader, pl.len = r0 * p, inner M = p - 1;  mode 1: Bluestein, inner M = pM >= 2 len - 1
    template <bool SW, int RMAX = 31>
    static bool make_smooth_conv_t(b200fft_plan& pl, int mode, uint32_t r0, uint32_t pM) {
        using KT = SmoothConvKernel<T, SW, RMAX>;
        const uint32_t n = (uint32_t)pl.len;
        const uint32_t M = mode == 0 ? pM - 1 : pM;
        std::vector<uint32_t> radices;
        if (!smooth_factor(M, radices) || (uint64_t)r0 * M > CONV_SMOOTH_MAX) return false;
        typename KT::Params base;
        if (!fill_smooth_pass<KT>(pl, base, M, radices)) return false;
        base.n = n;
        base.M = M;
        base.r0 = r0;
        base.div_r0 = make_fastdiv(r0);
        base.mode = mode == 0 ? KT::MODE_RADER : KT::MODE_BLUESTEIN;
        uint64_t groot = 0;
        std::vector<C> mult;
        if (mode == 0) {
            base.p = pM;
            std::vector<uint32_t> gpow, ginv;
            rader_tables(pM, M, gpow, ginv, mult, groot);
            base.gpow = upload(pl, gpow);
            base.ginv = upload(pl, ginv);
            if (!base.gpow || !base.ginv) return false;
            if (r0 > 1) {
                std::vector<C> otw((size_t)r0 * pM), wr(r0);
                for (uint64_t k1 = 0; k1 < r0; ++k1)
                    for (uint64_t n2 = 0; n2 < pM; ++n2) otw[(size_t)(k1 * pM + n2)] = hm::twiddle<T>(k1 * n2, n);
                for (uint64_t j = 0; j < r0; ++j) wr[(size_t)j] = hm::twiddle<T>(j, r0);
                base.otw = upload(pl, otw);
                base.w_r0 = upload(pl, wr);
                if (!base.otw || !base.w_r0) return false;
            }
        } else {
            base.p = n;
            std::vector<C> chirp;
            bluestein_tables(n, M, chirp, mult);
            base.chirp = upload(pl, chirp);
            if (!base.chirp) return false;
        }
        base.mult = upload(pl, mult);
        if (!base.mult) return false;
        // virtual transforms per CTA: a multiple of r0, both ping-pong buffers inside the budget, at most 64
        // (the instantiation without the prime butterflies fits three CTAs per SM when a CTA stays inside 64 KiB: measured +11..16 % at
        //  97 / 1009; the 128-register one runs two CTAs either way and is faster with more transforms per CTA: 2053 0.077 vs 0.060)
        const uint32_t budget = (RMAX <= 16 && (uint64_t)r0 * M <= SMOOTH_MAX) ? SMOOTH_MAX : CONV_SMOOTH_MAX;
        uint32_t F = std::max<uint32_t>(1, budget / M / r0) * r0;
        while (F > r0 && F > 64) F -= r0;
        base.f_per_cta = F;
        base.smem_bytes = (uint32_t)(((2ull * F * M + F) * sizeof(C) + 15) / 16 * 16);
        const size_t max_smem = (2ull * CONV_SMOOTH_MAX + 64) * sizeof(C) + 16;
        const uint32_t n_steps = 2 * (uint32_t)radices.size();
        // virtual transform indices of a launch stay below 2^31 (the kernel divides them by r0 in 32 bits); whole CTAs per segment
        const uint64_t seg = std::max<uint64_t>(F / r0, (((1ull << 31) - 1) / r0) / (F / r0) * (F / r0));
        pl.exec = [=](const ExecCtx& c) {
            for (uint64_t b0 = 0; b0 < c.batch; b0 += seg) {
                const uint64_t nb = std::min(seg, c.batch - b0);
                typename KT::Params q = base;
                q.in = (const C*)c.in + b0 * n;
                q.out = (C*)c.out + b0 * n;
                q.n_fft = nb;
                if (!rt::launch_loop<KT>(q, (nb * r0 + F - 1) / F, n_steps, q.smem_bytes, max_smem, c.stream)) return false;
            }
            return true;
        };
        pl.launches = [](uint64_t) { return (uint64_t)1; };
        const std::string inner = "Smooth{" + std::to_string(M) + "=" + radix_string(radices) + "}";
        if (mode == 0) {
            const std::string r = "Rader{n=" + std::to_string(pM) + ",g=" + std::to_string(groot) + ",inner=" + inner + ",fused}";
            pl.desc = r0 > 1 ? "MixedRadix{" + std::to_string(r0) + "x" + r + ",fused}" : r;
        } else {
            pl.desc = "Bluestein{n=" + std::to_string(n) + ",M=" + std::to_string(M) + ",inner=" + inner + ",fused}";
        }
        set_recipe(pl, mode == 0 ? B200FFT_RECIPE_RADER : B200FFT_RECIPE_BLUESTEIN, mode == 0 ? r0 : 0, 0, B200FFT_RECIPE_SMOOTH, M);
        return true;
    }

    // ---------------- large Rader / Bluestein over a smooth two-pass inner FFT of M = N1 * N2 ----------------
    // the four launches of make_big_conv (A1 gather|chirp-pad, B1 x mult + conj (+ DC), A2 plain, B2 conj + scatter | conj x chirp)
    // with the run-time-radix passes: any "easy" prime (p - 1 smooth) above the one-pass limit, Bluestein with the smallest smooth M
    template <bool SW, int RMAX = 31>
    static bool make_smooth_big_conv_t(b200fft_plan& pl, uint32_t N1, uint32_t N2, bool rader) {
        using KA = SmoothPassKernel<T, SW, 1, RMAX>;
        using KA2 = SmoothPassKernel<T, false, 1, RMAX>;
        using KB = SmoothPassKernel<T, SW, 2, RMAX>;
        const uint64_t n = pl.len, M = (uint64_t)N1 * N2;
        typename KA::Params pa;
        typename KB::Params pb;
        uint32_t FA = 0, FB = 0;
        if (!smooth_pass_pair<KA, KB>(pl, N1, N2, pa, pb, FA, FB)) return false;
        pb.full_tw = smooth_full_twiddles(pl, N1, N2);
        if (!pb.full_tw) return false;
        std::vector<C> mult;
        uint64_t groot = 0;
        pa.n_outer = pb.n_outer = (uint32_t)n;
        if (rader) {
            std::vector<uint32_t> gpow, ginv;
            rader_tables(n, M, gpow, ginv, mult, groot);
            pa.gather = upload(pl, gpow);
            pb.scatter = upload(pl, ginv);
            if (!pa.gather || !pb.scatter) return false;
        } else {
            std::vector<C> chirp;
            bluestein_tables(n, M, chirp, mult);
            pa.chirp = pb.chirp = upload(pl, chirp);
            if (!pa.chirp) return false;
        }
        pb.mult = upload(pl, mult);
        if (!pb.mult) return false;
        const size_t max_smem = 2ull * (SMOOTH_MAX + 64) * sizeof(C);
        // two workspaces per chunk inside ~32 MiB; FFT indices of a launch stay below 2^31
        uint64_t chunk = std::max<uint64_t>(1, (16ull << 20) / (M * sizeof(C)));
        chunk = std::min<uint64_t>(chunk, ((1ull << 31) - 1) / std::max(N1, N2));
        pl.chunk = chunk;
        pl.work_bytes = [=](uint64_t batch) { return 2 * std::min(batch, chunk) * M * sizeof(C); };
        pl.launches = [=](uint64_t batch) { return 4 * ((batch + chunk - 1) / chunk); };
        pl.exec = [=](const ExecCtx& c) {
            const C* in = (const C*)c.in;
            C* out = (C*)c.out;
            const uint64_t per = std::min(c.batch, chunk) * M;
            C* w1 = (C*)c.work;
            C* w2 = w1 + per;
            for (uint64_t b0 = 0; b0 < c.batch; b0 += chunk) {
                const uint64_t nb = std::min(chunk, c.batch - b0);
                typename KA::Params a1 = pa;  // gather | chirp-pad columns -> N1-point FFTs -> w1
                a1.conv = rader ? 1u : 2u;
                a1.in = in + b0 * n;
                a1.out = w1;
                a1.n_fft = nb * N2;
                if (!rt::launch_dyn<KA>(a1, (a1.n_fft + FA - 1) / FA, a1.smem_bytes, max_smem, c.stream)) return false;
                typename KB::Params b1 = pb;  // rows x twiddle -> N2-point FFTs -> x mult, conj (+ DC) -> w2, natural order
                b1.conv = 1;
                b1.in = w1;
                b1.out = w2;
                b1.x_in = rader ? in + b0 * n : nullptr;
                b1.x_out = out + b0 * n;
                b1.n_fft = nb * N1;
                if (!rt::launch_dyn<KB>(b1, (b1.n_fft + FB - 1) / FB, b1.smem_bytes, max_smem, c.stream)) return false;
                typename KA2::Params a2;  // plain pass A, in place
                std::memcpy(&a2, &pa, sizeof(a2));
                a2.conv = 0;
                a2.in = w2;
                a2.out = w2;
                a2.n_fft = nb * N2;
                if (!rt::launch_dyn<KA2>(a2, (a2.n_fft + FA - 1) / FA, a2.smem_bytes, max_smem, c.stream)) return false;
                typename KB::Params b2 = pb;  // rows x twiddle -> N2-point FFTs -> conj + scatter | conj x chirp -> out
                b2.conv = rader ? 2u : 3u;
                b2.in = w2;
                b2.out = out + b0 * n;
                b2.n_fft = nb * N1;
                if (!rt::launch_dyn<KB>(b2, (b2.n_fft + FB - 1) / FB, b2.smem_bytes, max_smem, c.stream)) return false;
            }
            return true;
        };
        const std::string inner = "SmoothFourStep{" + std::to_string(N1) + "x" + std::to_string(N2) + "}";
        pl.desc = rader ? "Rader{n=" + std::to_string(n) + ",g=" + std::to_string(groot) + ",inner=" + inner + "}"
                        : "Bluestein{n=" + std::to_string(n) + ",M=" + std::to_string(M) + ",inner=" + inner + "}";
        set_recipe(pl, rader ? B200FFT_RECIPE_RADER : B200FFT_RECIPE_BLUESTEIN, rader ? 1 : 0, 0, B200FFT_RECIPE_MIXED_RADIX, M, N1, N2);
        return true;
    }

    // ---------------- Bluestein (fused) ----------------
    static void bluestein_tables(uint64_t n, uint64_t M, std::vector<C>& chirp, std::vector<C>& mult) {
        chirp.resize((size_t)n);
        std::vector<hm::cld> c((size_t)M, hm::cld{0, 0});
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t m = (uint64_t)(((unsigned __int128)i * i) % (2 * n));
            const hm::cld w = hm::twiddle_ld(m, 2 * n);  // forward chirp W_2n^(i^2)
            chirp[(size_t)i] = mk<T>((T)w.x, (T)w.y);
            const hm::cld cj{w.x / (hm::ld)M, -w.y / (hm::ld)M};  // conj(w)/M
            c[(size_t)i] = cj;
            if (i) c[(size_t)(M - i)] = cj;
        }
        if (hm::is_pow2(M)) hm::fft_pow2_ld(c); else hm::fft_any_ld(c);
        mult.resize((size_t)M);
        for (uint64_t i = 0; i < M; ++i) mult[(size_t)i] = mk<T>((T)c[(size_t)i].x, (T)c[(size_t)i].y);
    }
    template <int M, bool SW>
    static bool make_bluestein_t(b200fft_plan& pl) {
        using G = typename DirectGeo<T, M>::type;
        using KT = BluesteinKernel<G, SW>;
        const uint32_t n = (uint32_t)pl.len;
        std::vector<C> chirp, mult;
        bluestein_tables(n, M, chirp, mult);
        const C* d_chirp = upload(pl, chirp);
        const C* d_mult = upload(pl, mult);
        const C* tw = G::TW_ELEMS ? upload(pl, stage_twiddles<G>()) : nullptr;
        if (!d_chirp || !d_mult || (G::TW_ELEMS && !tw)) return false;
        pl.exec = [=](const ExecCtx& c) {
            typename KT::Params p;
            p.in = (const C*)c.in;
            p.out = (C*)c.out;
            p.chirp = d_chirp;
            p.mult = d_mult;
            p.tw = tw;
            p.n = n;
            p.n_fft = c.batch;
            return rt::launch<KT>(p, (c.batch + G::F - 1) / G::F, c.stream);
        };
        pl.launches = [](uint64_t) { return (uint64_t)1; };
        pl.desc = "Bluestein{n=" + std::to_string(n) + ",M=" + std::to_string(M) + ",fused}";
        set_recipe(pl, B200FFT_RECIPE_BLUESTEIN, 0, 0, B200FFT_RECIPE_POW2, M);
        return true;
    }
    template <int M>
    static bool make_bluestein(b200fft_plan& pl) {
        return pl.direction ? make_bluestein_t<M, true>(pl) : make_bluestein_t<M, false>(pl);
    }
    static bool make_bluestein_rt(b200fft_plan& pl, uint32_t M) {
        switch (M) {
            case 8: return make_bluestein<8>(pl);
            case 16: return make_bluestein<16>(pl);
            case 32: return make_bluestein<32>(pl);
            case 64: return make_bluestein<64>(pl);
            case 128: return make_bluestein<128>(pl);
            case 256: return make_bluestein<256>(pl);
            case 512: return make_bluestein<512>(pl);
            case 1024: return make_bluestein<1024>(pl);
            case 2048: return make_bluestein<2048>(pl);
            case 4096: return make_bluestein<4096>(pl);
        }
        return false;
    }

    // ---------------- Rader (fused) ----------------
    // gpow[i] = g^(i+1) mod n, ginv[i] = g^-(i+1) mod n, mult = FFT_M( twiddle(g^-i mod n, n) / M )
    static void rader_tables(uint64_t n, uint64_t M, std::vector<uint32_t>& gpow, std::vector<uint32_t>& ginv,
                             std::vector<C>& mult, uint64_t& g) {
        g = hm::primitive_root(n);
        const uint64_t gi = hm::powmod(g, n - 2, n);
        gpow.resize((size_t)M);
        ginv.resize((size_t)M);
        std::vector<hm::cld> d((size_t)M);
        uint64_t a = 1, b = 1;
        for (uint64_t i = 0; i < M; ++i) {
            const hm::cld w = hm::twiddle_ld(b, n);  // b = g^-i
            d[(size_t)i] = hm::cld{w.x / (hm::ld)M, w.y / (hm::ld)M};
            a = hm::mulmod(a, g, n);
            b = hm::mulmod(b, gi, n);
            gpow[(size_t)i] = (uint32_t)a;
            ginv[(size_t)i] = (uint32_t)b;
        }
        if (hm::is_pow2(M)) hm::fft_pow2_ld(d); else hm::fft_any_ld(d);
        mult.resize((size_t)M);
        for (size_t i = 0; i < (size_t)M; ++i) mult[i] = mk<T>((T)d[i].x, (T)d[i].y);
    }
    template <int M, bool SW>
    static bool make_rader_t(b200fft_plan& pl) {
        using G = typename DirectGeo<T, M>::type;
        using KT = RaderKernel<G, SW>;
        const uint64_t n = pl.len;
        std::vector<uint32_t> gpow, ginv;
        std::vector<C> mult;
        uint64_t g = 0;
        rader_tables(n, (uint64_t)M, gpow, ginv, mult, g);
        const uint32_t* d_gpow = upload(pl, gpow);
        const uint32_t* d_ginv = upload(pl, ginv);
        const C* d_mult = upload(pl, mult);
        const C* tw = G::TW_ELEMS ? upload(pl, stage_twiddles<G>()) : nullptr;
        if (!d_gpow || !d_ginv || !d_mult || (G::TW_ELEMS && !tw)) return false;
        pl.exec = [=](const ExecCtx& c) {
            typename KT::Params p;
            p.in = (const C*)c.in;
            p.out = (C*)c.out;
            p.gpow = d_gpow;
            p.ginv = d_ginv;
            p.mult = d_mult;
            p.tw = tw;
            p.n = (uint32_t)n;
            p.n_fft = c.batch;
            return rt::launch<KT>(p, (c.batch + G::F - 1) / G::F, c.stream);
        };
        pl.launches = [](uint64_t) { return (uint64_t)1; };
        pl.desc = "Rader{n=" + std::to_string(n) + ",g=" + std::to_string(g) + ",fused}";
        set_recipe(pl, B200FFT_RECIPE_RADER, 1, 0, B200FFT_RECIPE_POW2, M);
        return true;
    }
    template <int M>
    static bool make_rader(b200fft_plan& pl) {
        return pl.direction ? make_rader_t<M, true>(pl) : make_rader_t<M, false>(pl);
    }
    static bool make_rader_rt(b200fft_plan& pl, uint32_t M) {
        switch (M) {
            case 2: return make_rader<2>(pl);
            case 4: return make_rader<4>(pl);
            case 16: return make_rader<16>(pl);
            case 256: return make_rader<256>(pl);
        }
        return false;
    }

    static bool smooth_factor_any(uint64_t n) {
        std::vector<uint32_t> r;
        return smooth_factor(n, r);
    }
    // smallest M = 2^a 3^b 5^c 7^d >= lo that the smooth kernels can run (one pass, or a two-pass split); 0 = none
    static uint64_t bluestein_smooth_len(uint64_t lo) {
        std::vector<uint64_t> cand;
        const uint64_t hi = 2 * lo;  // (the next power of two is below this)
        for (uint64_t a = 1; a < hi; a *= 7)
            for (uint64_t b = a; b < hi; b *= 5)
                for (uint64_t c = b; c < hi; c *= 3)
                    for (uint64_t d = c; d < hi; d *= 2)
                        if (d >= lo) cand.push_back(d);
        std::sort(cand.begin(), cand.end());
        for (uint64_t d : cand) {
            uint32_t s1 = 0, s2 = 0;
            if (d <= CONV_SMOOTH_MAX ? smooth_factor_any(d) : smooth_split(d, s1, s2)) return d;
        }
        return 0;
    }
    // B200FFT_GENERAL_RADER=0: primes other than the Fermat ones go through Bluestein, as in round 1 (A/B measurements)
    static bool use_general_rader() {
        static bool v = [] {
            const char* e = std::getenv("B200FFT_GENERAL_RADER");
            return !(e && std::atoi(e) == 0);
        }();
        return v;
    }
    // B200FFT_GOOD_THOMAS=1: coprime two-pass splits run as GoodThomas (index maps, no twiddle table) instead of SmoothFourStep
    static bool use_good_thomas() {
        static bool v = [] {
            const char* e = std::getenv("B200FFT_GOOD_THOMAS");
            return e && std::atoi(e) == 1;
        }();
        return v;
    }
    // B200FFT_BLUESTEIN_SMOOTH_BIG=1: also above the one-pass limit (four run-time-radix passes over the smooth length instead of
    // four compiled passes over the next power of two); otherwise reachable through a recipe only
    static bool rader_smooth_big() {
        static bool v = [] {
            const char* e = std::getenv("B200FFT_RADER_SMOOTH_BIG");
            return e && std::atoi(e) == 1;
        }();
        return v;
    }
    static bool bluestein_smooth_big() {
        static bool v = [] {
            const char* e = std::getenv("B200FFT_BLUESTEIN_SMOOTH_BIG");
            return e && std::atoi(e) == 1;
        }();
        return v;
    }
    static uint64_t bluestein_smooth_pct() {
        static uint64_t v = [] {
            const char* e = std::getenv("B200FFT_BLUESTEIN_SMOOTH");
            return e ? (uint64_t)std::atoi(e) : (uint64_t)0;  // measured: the smooth inner FFT loses to the next power of two at every ratio tried
        }();
        return v;
    }

    // ---------------- top level ----------------
    // a caller-supplied recipe (include/b200fft.h): node 0 is the root; every structural claim of the recipe is checked
    static int build_from_recipe(b200fft_plan& pl) {
        const std::vector<b200fft_recipe_node>& rc = pl.recipe;
        const b200fft_recipe_node& r = rc[0];
        const uint64_t n = pl.len;
        auto unsupported = [](const std::string& m) { return fail(B200FFT_ERR_UNSUPPORTED, "recipe: " + m); };
        auto child_of = [&](const b200fft_recipe_node& nd) -> const b200fft_recipe_node* {
            return (nd.child > 0 && nd.child < rc.size()) ? &rc[nd.child] : nullptr;
        };
        bool ok = false;
        switch (r.kind) {
            case B200FFT_RECIPE_POW2:
                if (!hm::is_pow2(n)) return unsupported("POW2 needs a power-of-two length");
                if (n <= DirectMax<T>::v) ok = make_direct_rt(pl, (uint32_t)n);
                else if (n <= (uint64_t)TILE_MAX * TILE_MAX) ok = make_four_step(pl, hm::ilog2(n));
                else return unsupported("power-of-two lengths above 2^24");
                break;
            case B200FFT_RECIPE_SMOOTH: {
                std::vector<uint32_t> radices;
                if (n > SMOOTH_MAX || !smooth_factor(n, radices)) return unsupported("SMOOTH needs prime factors <= 31 and a one-pass length");
                ok = smooth_dispatch(pl, 0, 0, 0);
                break;
            }
            case B200FFT_RECIPE_MIXED_RADIX:
            case B200FFT_RECIPE_GOOD_THOMAS: {
                uint64_t a = std::min(r.a, r.b), b = std::max(r.a, r.b);
                if (a < 2 || a * b != n) return unsupported("the split must multiply to the length");
                if (r.kind == B200FFT_RECIPE_GOOD_THOMAS && hm::gcd(a, b) != 1) return unsupported("GOOD_THOMAS needs a coprime split");
                if (hm::is_pow2(n) && r.kind == B200FFT_RECIPE_MIXED_RADIX) {
                    const uint32_t lg = hm::ilog2(n);
                    if (a != (1ull << (lg / 2)) || n <= DirectMax<T>::v) return unsupported("power-of-two MIXED_RADIX must be the balanced two-pass split");
                    ok = make_four_step(pl, lg);
                    break;
                }
                if (b > SMOOTH_MAX || !smooth_factor_any(a) || !smooth_factor_any(b)) return unsupported("both factors must be smooth one-pass lengths");
                if (r.kind == B200FFT_RECIPE_MIXED_RADIX && sizeof(T) == 4 && use_compiled_smooth() && compiled_len(a) && compiled_len(b)) {
                    ok = make_compiled_smooth(pl, (uint32_t)a, (uint32_t)b);
                    break;
                }
                ok = smooth_dispatch(pl, r.kind == B200FFT_RECIPE_GOOD_THOMAS ? 4 : 1, (uint32_t)a, (uint32_t)b);
                break;
            }
            case B200FFT_RECIPE_RADER: {
                const uint64_t r0 = r.a > 1 ? r.a : 1;
                if (n % r0 || !hm::is_prime(n / r0) || n / r0 < 3) return unsupported("RADER needs len = a * prime");
                const uint64_t p = n / r0, M = p - 1;
                const b200fft_recipe_node* in = child_of(r);
                if (in && in->len != M) return unsupported("the inner FFT of RADER has length p - 1");
                const uint32_t ik = in ? in->kind : (uint32_t)B200FFT_RECIPE_AUTO;
                uint32_t s1 = 0, s2 = 0;
                if (ik == B200FFT_RECIPE_CLUSTER) {
                    if (sizeof(T) != 4 || r0 != 1 || (M != (1u << 14) && M != (1u << 16))) return unsupported("RADER over a CLUSTER inner FFT: f32, p - 1 = 2^14 or 2^16");
                    ok = make_cluster_conv(pl, M, 0);
                } else if (hm::is_pow2(M) && r0 == 1 && (ik == B200FFT_RECIPE_AUTO || ik == B200FFT_RECIPE_POW2 || ik == B200FFT_RECIPE_MIXED_RADIX)) {
                    if (M <= 256) ok = make_rader_rt(pl, (uint32_t)M);
                    else if (M <= (uint64_t)TILE_MAX * TILE_MAX && M >= (uint64_t)TILE_MIN * TILE_MIN) ok = make_big_conv(pl, M, true);
                    else return unsupported("RADER over this power-of-two length");
                } else if ((ik == B200FFT_RECIPE_AUTO || ik == B200FFT_RECIPE_SMOOTH) && r0 <= 8 && r0 * M <= CONV_SMOOTH_MAX && smooth_factor_any(M)) {
                    ok = smooth_dispatch(pl, 2, (uint32_t)r0, (uint32_t)p);
                } else if (r0 == 1 && (ik == B200FFT_RECIPE_AUTO || ik == B200FFT_RECIPE_MIXED_RADIX) && smooth_factor_any(M)) {
                    if (in && ik == B200FFT_RECIPE_MIXED_RADIX) {
                        s1 = (uint32_t)std::min(in->a, in->b);
                        s2 = (uint32_t)std::max(in->a, in->b);
                        if ((uint64_t)s1 * s2 != M || s2 > SMOOTH_MAX || !smooth_factor_any(s1) || !smooth_factor_any(s2))
                            return unsupported("inner split of RADER");
                    } else if (!smooth_split(M, s1, s2)) {
                        return unsupported("RADER: p - 1 has no two-pass split");
                    }
                    ok = smooth_dispatch(pl, 5, s1, s2);
                } else {
                    return unsupported("RADER: p - 1 must factor into primes <= 31");
                }
                break;
            }
            case B200FFT_RECIPE_BLUESTEIN: {
                const b200fft_recipe_node* in = child_of(r);
                const uint64_t M = in ? in->len : hm::next_pow2(2 * n - 1);
                if (n < 2 || M < 2 * n - 1) return unsupported("the inner FFT of BLUESTEIN needs length >= 2 len - 1");
                const uint32_t ik = in ? in->kind : (uint32_t)B200FFT_RECIPE_AUTO;
                uint32_t s1 = 0, s2 = 0;
                if (ik == B200FFT_RECIPE_CLUSTER) {
                    if (sizeof(T) != 4 || (M != (1u << 14) && M != (1u << 16))) return unsupported("BLUESTEIN over a CLUSTER inner FFT: f32, M = 2^14 or 2^16");
                    ok = make_cluster_conv(pl, M, 1);
                } else if (hm::is_pow2(M) && ik != B200FFT_RECIPE_SMOOTH) {
                    if (M <= FUSED_CONV_MAX) ok = make_bluestein_rt(pl, (uint32_t)std::max<uint64_t>(M, 8));
                    else if (M <= (uint64_t)TILE_MAX * TILE_MAX) ok = make_big_conv(pl, M, false);
                    else return unsupported("BLUESTEIN over this power-of-two length");
                } else if (M <= CONV_SMOOTH_MAX && smooth_factor_any(M) && ik != B200FFT_RECIPE_MIXED_RADIX) {
                    ok = smooth_dispatch(pl, 3, (uint32_t)M, 0);
                } else if (smooth_factor_any(M)) {
                    if (in && ik == B200FFT_RECIPE_MIXED_RADIX && in->a * in->b == M) {
                        s1 = (uint32_t)std::min(in->a, in->b);
                        s2 = (uint32_t)std::max(in->a, in->b);
                        if (s2 > SMOOTH_MAX || !smooth_factor_any(s1) || !smooth_factor_any(s2)) return unsupported("inner split of BLUESTEIN");
                    } else if (!smooth_split(M, s1, s2)) {
                        return unsupported("BLUESTEIN: the inner length has no two-pass split");
                    }
                    ok = smooth_dispatch(pl, 6, s1, s2);
                } else {
                    return unsupported("BLUESTEIN: the inner length must be a power of two or factor into primes <= 31");
                }
                break;
            }
            case B200FFT_RECIPE_COLUMNS: {
                if (r.a < 2 || r.b < 1 || r.a * r.b != n || r.a > SMOOTH_MAX || !smooth_factor_any(r.a) || r.b >= (1ull << 31))
                    return unsupported("COLUMNS: len = a * b, a smooth and at most the one-pass limit");
                ok = smooth_dispatch(pl, 7, (uint32_t)r.a, (uint32_t)r.b);
                break;
            }
            case B200FFT_RECIPE_CLUSTER: {
                if (sizeof(T) != 4 || !hm::is_pow2(n) || n < (1u << 14) || n > (1u << 17)) return unsupported("CLUSTER plans exist for f32, 2^14 .. 2^17");
                if (r.a == 1 && n > (1u << 16)) return unsupported("half-tile CLUSTER plans exist for 2^14 .. 2^16");
                ok = make_cluster(pl, hm::ilog2(n), r.a == 1);
                break;
            }
            default:
                return fail(B200FFT_ERR_INVALID_ARG, "recipe: unknown node kind");
        }
        if (!ok) return fail(B200FFT_ERR_CUDA, "plan construction failed: " + rt::last_error());
        return B200FFT_OK;
    }

    static int build(b200fft_plan& pl) {
        const uint64_t n = pl.len;
        if (!pl.recipe.empty() && pl.recipe[0].kind != B200FFT_RECIPE_AUTO && n > 1) return build_from_recipe(pl);
        if (n <= 1) {
            pl.exec = [n](const ExecCtx& c) {
                if (n == 0 || c.in == c.out || c.batch == 0) return true;
                return rt::d2d_async(c.out, c.in, c.batch * n * sizeof(C), c.stream);
            };
            pl.desc = "Identity{" + std::to_string(n) + "}";
            set_recipe(pl, B200FFT_RECIPE_AUTO);
            return B200FFT_OK;
        }
        // lengths this build cannot plan are rejected BEFORE any factoring / primality work (a prime near 2^62 would otherwise
        // spin in trial division, and 2 n - 1 wraps above 2^63)
        if (!hm::is_pow2(n) && n > (1ull << 23))
            return fail(B200FFT_ERR_UNSUPPORTED, "non-power-of-two lengths above 2^23 are not planned by this build");
        bool ok = false;
        if (hm::is_pow2(n)) {
            const uint32_t lgn = hm::ilog2(n);
            if (sizeof(T) == 4 && lgn >= 14 && lgn <= 17 && ((cluster_mask() >> (lgn - 14)) & 1u))
                ok = make_cluster(pl, lgn);  // one pass over HBM, transposed through distributed shared memory
            else if (n <= DirectMax<T>::v)
                ok = make_direct_rt(pl, (uint32_t)n);
            else if (n <= (uint64_t)TILE_MAX * TILE_MAX)
                ok = make_four_step(pl, hm::ilog2(n));
            else
                return fail(B200FFT_ERR_UNSUPPORTED, "power-of-two lengths above 2^24 are not planned by this build");
        } else if (std::vector<uint32_t> radices; n <= SMOOTH_MAX && smooth_factor(n, radices)) {
            ok = smooth_dispatch(pl, 0, 0, 0);  // every prime factor <= 31
        } else if (uint32_t c1 = 0, c2 = 0; n > SMOOTH_MAX && compiled_pair(n, c1, c2)) {
            ok = make_compiled_smooth(pl, c1, c2);  // two passes through compiled composite tiles
        } else if (uint32_t g1 = 0, g2 = 0; use_good_thomas() && n > SMOOTH_MAX && coprime_split(n, g1, g2)) {
            ok = smooth_dispatch(pl, 4, g1, g2);  // opt-in (B200FFT_GOOD_THOMAS=1): coprime split, no inter-pass twiddles
        } else if (uint32_t s1 = 0, s2 = 0; n > SMOOTH_MAX && n <= (1ull << 23) && smooth_split(n, s1, s2)) {
            ok = smooth_dispatch(pl, 1, s1, s2);  // composite of small primes: two passes instead of Bluestein's four
        } else if (hm::is_prime(n) && hm::is_pow2(n - 1) && n - 1 <= 256) {
            ok = make_rader_rt(pl, (uint32_t)(n - 1));
        } else if (sizeof(T) == 4 && n == 65537 && use_cluster_conv()) {
            ok = make_cluster_conv(pl, 65536, 0);  // the whole Rader algorithm inside one cluster pass
        } else if (hm::is_prime(n) && hm::is_pow2(n - 1) && n - 1 <= (uint64_t)TILE_MAX * TILE_MAX) {
            ok = make_big_conv(pl, n - 1, true);  // 65537
        } else {
            // what is left has a prime factor p > 31.  The reference's rule (src/plan.rs:636-664): Rader when p - 1 factors into small
            // primes ("easy" prime), else Bluestein -- with the freedom to pick any inner length >= 2n - 1.
            const uint64_t p = hm::largest_prime_factor(n), r0 = n / p;
            std::vector<uint32_t> rr;
            uint32_t b1 = 0, b2 = 0;
            // Measured on B200 (profiles/r2d_ab_plans.txt): the one-pass Rader over a smooth p - 1 beats the fused Bluestein it replaces
            // by 1.1-1.8x in f64 at every size and in f32 from the point where Bluestein needs M >= 2048 (617: 1.7x, 2053: 1.3x, 1009 and
            // 1234: a tie); below that the compiled power-of-two Bluestein kernels win (37, 97: 0.75x) and keep the length.  Above
            // the one-pass limit Rader runs four run-time-radix passes, ~2x SLOWER than Bluestein's four compiled power-of-two passes
            // over 2-4x the data (7681, 112501): recipe / B200FFT_RADER_SMOOTH_BIG=1 only.
            const bool rader_one_pass_wins = sizeof(T) == 8 || hm::next_pow2(2 * n - 1) >= 2048;
            if (use_general_rader() && rader_one_pass_wins && r0 <= 8 && r0 * (p - 1) <= CONV_SMOOTH_MAX && smooth_factor(p - 1, rr)) {
                ok = smooth_dispatch(pl, 2, (uint32_t)r0, (uint32_t)p);  // one CTA pass: MixedRadix{r0 x Rader(p)} fused
            } else if (use_general_rader() && rader_smooth_big() && r0 == 1 && smooth_factor_any(p - 1) && smooth_split(p - 1, b1, b2)) {
                ok = smooth_dispatch(pl, 5, b1, b2);  // easy prime above the one-pass limit: four passes over n - 1
            } else {
                const uint64_t M2 = hm::next_pow2(2 * n - 1);
                const uint64_t M3 = bluestein_smooth_len(2 * n - 1);
                // the run-time-radix kernels cost more per point than the compiled power-of-two ones: the smooth length must
                // be clearly shorter to win (B200FFT_BLUESTEIN_SMOOTH = percentage of M2 it must stay below; 0 = never)
                const bool smooth_wins = M3 != 0 && M3 * 100 <= M2 * bluestein_smooth_pct();
                if (smooth_wins && M3 <= CONV_SMOOTH_MAX)
                    ok = smooth_dispatch(pl, 3, (uint32_t)M3, 0);
                else if (sizeof(T) == 4 && use_cluster_conv() && (M2 == (1u << 14) || M2 == (1u << 16)))
                    ok = make_cluster_conv(pl, M2, 1);
                else if (M2 <= FUSED_CONV_MAX)
                    ok = make_bluestein_rt(pl, (uint32_t)std::max<uint64_t>(M2, 8));
                else if (smooth_wins && M3 > CONV_SMOOTH_MAX && bluestein_smooth_big() && smooth_split(M3, b1, b2))
                    ok = smooth_dispatch(pl, 6, b1, b2);  // opt-in: the run-time-radix passes are slower per byte than the compiled ones
                else if (M2 <= (uint64_t)TILE_MAX * TILE_MAX)
                    ok = make_big_conv(pl, M2, false);
                else
                    return fail(B200FFT_ERR_UNSUPPORTED,
                                "non-power-of-two lengths above 2^23 are not planned by this build");
            }
        }
        if (!ok) return fail(B200FFT_ERR_CUDA, "plan construction failed: " + rt::last_error());
        return B200FFT_OK;
    }
};

// per-precision entry points of the planner, one per translation unit
int build_plan_f32(b200fft_plan& pl);
int build_plan_f64(b200fft_plan& pl);
#if defined(B2_PART_F32)
int build_plan_f32(b200fft_plan& pl) { return Builder<float>::build(pl); }
#endif
#if defined(B2_PART_F64)
int build_plan_f64(b200fft_plan& pl) { return Builder<double>::build(pl); }
#endif
#if defined(B2_PART_SMOOTH32)
bool build_smooth_f32(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) { return Builder<float>::smooth_build_onepass<31>(pl, kind, a, b); }
#endif
#if defined(B2_PART_SMOOTH32P)
bool build_smooth_passes_f32(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) { return Builder<float>::smooth_build_twopass<31>(pl, kind, a, b); }
#endif
#if defined(B2_PART_SMOOTH32S)
bool build_smooth_small_f32(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) {
    return (kind == 0 || kind == 2 || kind == 3) ? Builder<float>::smooth_build_onepass<16>(pl, kind, a, b) : Builder<float>::smooth_build_twopass<16>(pl, kind, a, b);
}
#endif
#if defined(B2_PART_FUSED32)
bool build_cluster_f32(b200fft_plan& pl, uint32_t lgN, bool half) { return Builder<float>::cluster_build_here(pl, lgN, half); }
bool build_cluster_conv_f32(b200fft_plan& pl, uint64_t M, int mode) { return Builder<float>::cluster_conv_build_here(pl, M, mode); }
bool build_fused_f32(b200fft_plan& pl, uint32_t L1, uint32_t L2, uint32_t lgN, const void* full_tw, FusedFn& fn, uint32_t& W) {
    return Builder<float>::fused_build_here(pl, L1, L2, lgN, (const cx<float>*)full_tw, fn, W);
}
#endif
#if defined(B2_PART_CTILE32)
bool build_compiled_smooth_f32(b200fft_plan& pl, uint32_t a, uint32_t b) { return Builder<float>::compiled_smooth_build_here(pl, a, b); }
#endif
#if defined(B2_PART_SMOOTH64)
bool build_smooth_f64(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) { return Builder<double>::smooth_build_onepass<31>(pl, kind, a, b); }
#endif
#if defined(B2_PART_SMOOTH64P)
bool build_smooth_passes_f64(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) { return Builder<double>::smooth_build_twopass<31>(pl, kind, a, b); }
#endif
#if defined(B2_PART_SMOOTH64S)
bool build_smooth_small_f64(b200fft_plan& pl, int kind, uint32_t a, uint32_t b) {
    return (kind == 0 || kind == 2 || kind == 3) ? Builder<double>::smooth_build_onepass<16>(pl, kind, a, b) : Builder<double>::smooth_build_twopass<16>(pl, kind, a, b);
}
#endif

}  // namespace b2

#if defined(B2_PART_CABI)
namespace b2 {
static int validate_len(const b200fft_plan* pl, uint64_t n_in, uint64_t n_out, bool two) {
    // src/common.rs:13-104 (messages kept verbatim; the Rust shim panics with them)
    const uint64_t len = pl->len;
    if (two && n_in != n_out)
        return fail(B200FFT_ERR_LEN_MISMATCH,
                    "Provided FFT input buffer and output buffer must have the same length. Got input.len() = " +
                        std::to_string(n_in) + ", output.len() = " + std::to_string(n_out));
    if (n_in == 0) return B200FFT_OK;  // zero chunks validate fine (src/array_utils.rs:151-177)
    if (n_in < len)
        return fail(B200FFT_ERR_BUFFER_TOO_SMALL, "Provided FFT buffer was too small. Expected len = " +
                                                      std::to_string(len) + ", got len = " + std::to_string(n_in));
    if (n_in % len != 0)
        return fail(B200FFT_ERR_NOT_MULTIPLE, "Input FFT buffer must be a multiple of FFT length. Expected multiple of " +
                                                  std::to_string(len) + ", got len = " + std::to_string(n_in));
    return B200FFT_OK;
}

static int exec_device_impl(const b200fft_plan* pl, const void* d_in, void* d_out, uint64_t batch, rt::stream_t stream,
                            void* ws, uint64_t ws_bytes, bool ws_given, int max_streams = 0) {
    if (!pl || (!d_in && batch && pl->len) || (!d_out && batch && pl->len)) return fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    if (pl->len == 0 || batch == 0) return B200FFT_OK;  // src/fft_helper.rs:16-18
    rt::DeviceGuard guard(pl->device);  // (restores the caller's current device on every exit path)
    if (!guard.ok) return fail(B200FFT_ERR_CUDA, rt::last_error());
    const uint64_t need = pl->work_bytes(batch);
    void* work = ws;
    bool own = false;
    if (need) {
        if (ws_given) {
            if (ws_bytes < need || !ws)
                return fail(B200FFT_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
            // the second pass drops consumed workspace lines with discard.global.L2 (128-byte aligned by definition) and
            // the tiled passes address the workspace through TMA
            if (reinterpret_cast<uintptr_t>(ws) & 127u)
                return fail(B200FFT_ERR_INVALID_ARG, "workspace must be 128-byte aligned");
        } else {
            work = rt::malloc_async(need, stream);
            if (!work) return fail(B200FFT_ERR_CUDA, "workspace allocation failed: " + rt::last_error());
            own = true;
        }
    }
    ExecCtx c{d_in, d_out, work, batch, stream, max_streams};
    const bool ok = pl->exec(c);
    if (own) rt::free_async(work, stream);
    if (!ok) return fail(B200FFT_ERR_CUDA, "kernel launch failed: " + rt::last_error());
    return B200FFT_OK;
}

// Host-slice path: chunks of the batch flow H2D -> kernels -> D2H on two streams with two device
// buffers, so copies in both directions overlap compute when the caller's memory is pinned.
static uint64_t host_chunk_bytes_cfg() {
    static const uint64_t v = [] {
        const char* e = std::getenv("B200FFT_HOST_CHUNK_MB");  // staging granularity (tests shrink it)
        const uint64_t mb = e ? std::strtoull(e, nullptr, 10) : 64;
        return (mb < 1 ? 1 : mb) << 20;
    }();
    return v;
}

static int exec_host_impl_2stream(const b200fft_plan* pl, const void* in, void* out, uint64_t n_complex) {
    if (!pl) return fail(B200FFT_ERR_INVALID_ARG, "null plan");
    if (pl->len == 0 || n_complex == 0) return B200FFT_OK;
    if (!in || !out) return fail(B200FFT_ERR_INVALID_ARG, "null buffer");
    if (!rt::set_device(pl->device)) return fail(B200FFT_ERR_CUDA, rt::last_error());
    const uint64_t esz = pl->precision == B200FFT_F32 ? 8 : 16;
    const uint64_t batch = n_complex / pl->len;
    const uint64_t tbytes = pl->len * esz;
    uint64_t chunk = std::max<uint64_t>(1, host_chunk_bytes_cfg() / tbytes);
    if (chunk > batch) chunk = batch;
    const int NBUF = 2;  // (three 32 MiB buffers measured slower: 2.50 s vs 1.86 s per sweep step, round 1)
    void* dbuf[NBUF] = {nullptr, nullptr};
    void* wbuf[NBUF] = {nullptr, nullptr};
    rt::stream_t st[NBUF] = {nullptr, nullptr};
    const uint64_t wbytes = pl->work_bytes(chunk);
    int rc = B200FFT_OK;
    for (int i = 0; i < NBUF && rc == B200FFT_OK; ++i) {
        st[i] = rt::stream_create();
        dbuf[i] = rt::dmalloc(chunk * tbytes);
        if (wbytes) wbuf[i] = rt::dmalloc(wbytes);
        if (!st[i] || !dbuf[i] || (wbytes && !wbuf[i])) rc = fail(B200FFT_ERR_CUDA, "staging allocation failed: " + rt::last_error());
    }
    uint64_t idx = 0;
    for (uint64_t b0 = 0; b0 < batch && rc == B200FFT_OK; b0 += chunk, ++idx) {
        const int s = (int)(idx % NBUF);
        const uint64_t nb = std::min(chunk, batch - b0);
        const char* src = (const char*)in + b0 * tbytes;
        char* dst = (char*)out + b0 * tbytes;
        if (!rt::h2d_async(dbuf[s], src, nb * tbytes, st[s])) { rc = fail(B200FFT_ERR_CUDA, rt::last_error()); break; }
        rc = exec_device_impl(pl, dbuf[s], dbuf[s], nb, st[s], wbuf[s], wbytes, wbytes != 0, 1);
        if (rc != B200FFT_OK) break;
        if (!rt::d2h_async(dst, dbuf[s], nb * tbytes, st[s])) { rc = fail(B200FFT_ERR_CUDA, rt::last_error()); break; }
    }
    for (int i = 0; i < NBUF; ++i) {
        if (st[i] && !rt::stream_sync(st[i]) && rc == B200FFT_OK) rc = fail(B200FFT_ERR_CUDA, rt::last_error());
    }
    for (int i = 0; i < NBUF; ++i) {
        if (dbuf[i]) rt::dfree(dbuf[i]);
        if (wbuf[i]) rt::dfree(wbuf[i]);
        if (st[i]) rt::stream_destroy(st[i]);
    }
    return rc;
}

// ---- copy pool: a pageable caller's slices are copied to / from the pinned staging ring by several threads ------------------
// (one thread moves ~10 GB/s; the link wants ~50 GB/s in each direction at once)
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool p;
        return p;
    }
    void copy(void* dst, const void* src, size_t bytes) {
        const size_t MIN_PART = 4u << 20;
        if (bytes < 2 * MIN_PART || workers_.empty()) {
            std::memcpy(dst, src, bytes);
            return;
        }
        std::unique_lock<std::mutex> job_lock(job_mutex_);  // one job at a time
        const size_t parts = std::min<size_t>(workers_.size() + 1, (bytes + MIN_PART - 1) / MIN_PART);
        const size_t per = ((bytes + parts - 1) / parts + 63) / 64 * 64;
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = (char*)dst;
            src_ = (const char*)src;
            bytes_ = bytes;
            per_ = per;
            next_ = 1;  // part 0 is the caller's
            parts_ = parts;
            done_ = 0;
            ++gen_;
        }
        cv_.notify_all();
        std::memcpy(dst, src, std::min(per, bytes));
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return done_ == parts_ - 1; });
        parts_ = 0;
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        const char* e = std::getenv("B200FFT_COPY_THREADS");
        unsigned n = e ? (unsigned)std::atoi(e) : std::min(12u, hw > 2 ? hw / 2 : 1u);
        for (unsigned i = 1; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            size_t part;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || (gen_ != seen && next_ < parts_) ; });
                if (stop_) return;
                part = next_++;
                if (next_ >= parts_) seen = gen_;
            }
            const size_t off = part * per_;
            if (off < bytes_) std::memcpy(dst_ + off, src_ + off, std::min(per_, bytes_ - off));
            {
                std::lock_guard<std::mutex> g(m_);
                ++done_;
            }
            cv_done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, job_mutex_;
    std::condition_variable cv_, cv_done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, per_ = 0, next_ = 0, parts_ = 0, done_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// Host-slice path (default; B200FFT_HOST_PIPE=2 selects the two-stream version above): a three-stage pipeline over a ring of four
// device buffers -- one stream only copies in, one only computes, one only copies out, ordered per chunk by events -- so both PCIe
// directions and the SMs are busy at once.  The pipeline's resources (streams, events, device ring, workspace, pinned staging ring)
// belong to the plan: created on first use, reused by every later call (a batch-1 `process()` pays no allocation).  A caller's
// pageable memory (what a Rust Vec is) cannot be the source of an asynchronous copy, so it is staged: copy-pool threads move slice c
// into pinned slot c mod 4 while the device works on the slices before it, and finished slices back out; pinned / registered
// buffers are used in place.  B200FFT_HOST_STAGE=0/1 forces the choice.
static int exec_host_impl(const b200fft_plan* pl_c, const void* in, void* out, uint64_t n_complex) {
    static const bool three_stage = [] {
        const char* e = std::getenv("B200FFT_HOST_PIPE");
        return !(e && std::atoi(e) == 2);
    }();
    if (!three_stage) return exec_host_impl_2stream(pl_c, in, out, n_complex);
    if (!pl_c) return fail(B200FFT_ERR_INVALID_ARG, "null plan");
    if (pl_c->len == 0 || n_complex == 0) return B200FFT_OK;
    if (!in || !out) return fail(B200FFT_ERR_INVALID_ARG, "null buffer");
    b200fft_plan* pl = const_cast<b200fft_plan*>(pl_c);  // (the pipeline pool is the plan's only mutable state, under its mutex)
    rt::DeviceGuard guard(pl->device);
    if (!guard.ok) return fail(B200FFT_ERR_CUDA, rt::last_error());
    const uint64_t esz = pl->precision == B200FFT_F32 ? 8 : 16;
    const uint64_t batch = n_complex / pl->len;
    const uint64_t tbytes = pl->len * esz;
    uint64_t chunk = std::max<uint64_t>(1, host_chunk_bytes_cfg() / tbytes);
    if (chunk > batch) chunk = batch;
    const uint64_t nchunks = (batch + chunk - 1) / chunk;
    static const int force_stage = [] {
        const char* e = std::getenv("B200FFT_HOST_STAGE");
        return e ? std::atoi(e) : -1;
    }();
    const bool stage = force_stage >= 0 ? force_stage != 0 : !(rt::host_is_pinned(in) && rt::host_is_pinned(out));
    constexpr int NB = HostPipe::NB;

    HostPipe* hp = nullptr;
    {
        std::lock_guard<std::mutex> g(pl->pipe_mutex);
        if (!pl->pipes.empty()) {
            hp = pl->pipes.back();
            pl->pipes.pop_back();
        }
    }
    if (!hp) hp = new HostPipe();
    int rc = B200FFT_OK;
    auto cuda_fail = [&](const char* what) { return fail(B200FFT_ERR_CUDA, std::string(what) + ": " + rt::last_error()); };
    // (re)size the resources
    const uint64_t need_buf = chunk * tbytes, need_w = pl->work_bytes(chunk);
    const int slots = (int)std::min<uint64_t>(NB, nchunks);
    if (!hp->s_in) {
        hp->s_in = rt::stream_create();
        hp->s_k = rt::stream_create();
        hp->s_out = rt::stream_create();
        for (int i = 0; i < NB; ++i) {
            hp->ev_in[i] = rt::event_create();
            hp->ev_k[i] = rt::event_create();
            hp->ev_out[i] = rt::event_create();
        }
        if (!hp->s_in || !hp->s_k || !hp->s_out || !hp->ev_out[NB - 1]) rc = cuda_fail("pipeline creation failed");
    }
    if (rc == B200FFT_OK && (hp->dbuf_bytes < need_buf || !hp->dbuf[slots - 1])) {
        const uint64_t sz = std::max(hp->dbuf_bytes, need_buf);
        for (int i = 0; i < NB; ++i) {
            if (hp->dbuf[i] && hp->dbuf_bytes < sz) {
                rt::dfree(hp->dbuf[i]);
                hp->dbuf[i] = nullptr;
            }
            if (i < slots && !hp->dbuf[i]) {
                hp->dbuf[i] = rt::dmalloc(sz);
                if (!hp->dbuf[i]) {
                    rc = cuda_fail("device ring allocation failed");
                    break;
                }
            }
        }
        hp->dbuf_bytes = sz;
    }
    if (rc == B200FFT_OK && stage && (hp->pin_bytes < need_buf || !hp->pin_in[slots - 1])) {
        const uint64_t sz = std::max(hp->pin_bytes, need_buf);
        for (int i = 0; i < NB; ++i) {
            if (hp->pin_in[i] && hp->pin_bytes < sz) {
                rt::host_free_pinned(hp->pin_in[i]);
                rt::host_free_pinned(hp->pin_out[i]);
                hp->pin_in[i] = hp->pin_out[i] = nullptr;
            }
            if (i < slots && !hp->pin_in[i]) {
                hp->pin_in[i] = rt::host_alloc_pinned(sz);
                hp->pin_out[i] = rt::host_alloc_pinned(sz);
                if (!hp->pin_in[i] || !hp->pin_out[i]) {
                    rc = cuda_fail("pinned staging allocation failed");
                    break;
                }
            }
        }
        hp->pin_bytes = sz;
    }
    if (rc == B200FFT_OK && need_w > hp->wbytes) {
        if (hp->wbuf) rt::dfree(hp->wbuf);
        hp->wbuf = rt::dmalloc(need_w);
        hp->wbytes = hp->wbuf ? need_w : 0;
        if (!hp->wbuf) rc = cuda_fail("workspace allocation failed");
    }

    CopyPool& pool = CopyPool::get();
    auto drain = [&](uint64_t idx) {  // staged: slice idx has been copied out to its pinned slot -> the caller's memory
        const int b = (int)(idx % (uint64_t)NB);
        const uint64_t b0 = idx * chunk, nb = std::min(chunk, batch - b0);
        if (!rt::event_sync(hp->ev_out[b])) return false;
        pool.copy((char*)out + b0 * tbytes, hp->pin_out[b], nb * tbytes);
        return true;
    };
    uint64_t idx = 0, drained = 0;
    for (uint64_t b0 = 0; b0 < batch && rc == B200FFT_OK; b0 += chunk, ++idx) {
        const int b = (int)(idx % (uint64_t)NB);
        const uint64_t nb = std::min(chunk, batch - b0);
        const char* src = (const char*)in + b0 * tbytes;
        char* dst = (char*)out + b0 * tbytes;
        bool ok = true;
        if (stage) {
            if (idx >= (uint64_t)NB) {  // the slot's previous slice must be out of the staging buffers before they are refilled
                ok = drain(drained);
                ++drained;
            }
            if (ok) pool.copy(hp->pin_in[b], src, nb * tbytes);
            src = (const char*)hp->pin_in[b];
            dst = (char*)hp->pin_out[b];
        } else if (idx >= (uint64_t)NB) {
            ok = rt::stream_wait(hp->s_in, hp->ev_out[b]);  // ring slot drained
        }
        ok = ok && rt::h2d_async(hp->dbuf[b], src, nb * tbytes, hp->s_in) && rt::event_record(hp->ev_in[b], hp->s_in) &&
             rt::stream_wait(hp->s_k, hp->ev_in[b]);
        if (!ok) {
            rc = cuda_fail("host pipeline (copy in)");
            break;
        }
        rc = exec_device_impl(pl, hp->dbuf[b], hp->dbuf[b], nb, hp->s_k, hp->wbuf, hp->wbytes, hp->wbytes != 0, 1);
        if (rc != B200FFT_OK) break;
        ok = rt::event_record(hp->ev_k[b], hp->s_k) && rt::stream_wait(hp->s_out, hp->ev_k[b]) && rt::d2h_async(dst, hp->dbuf[b], nb * tbytes, hp->s_out) &&
             rt::event_record(hp->ev_out[b], hp->s_out);
        if (!ok) {
            rc = cuda_fail("host pipeline (copy out)");
            break;
        }
    }
    if (stage && rc == B200FFT_OK)
        for (; drained < idx; ++drained)
            if (!drain(drained)) {
                rc = cuda_fail("host pipeline (drain)");
                break;
            }
    rt::stream_t all[3] = {hp->s_in, hp->s_k, hp->s_out};
    for (rt::stream_t st : all)
        if (st && !rt::stream_sync(st) && rc == B200FFT_OK) rc = cuda_fail("host pipeline (sync)");
    {
        std::lock_guard<std::mutex> g(pl->pipe_mutex);
        pl->pipes.push_back(hp);
    }
    return rc;
}

}  // namespace b2

extern "C" {

int b200fft_device_count(int* n) {
    if (!n) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    *n = b2::rt::device_count();
    return B200FFT_OK;
}

int b200fft_plan_create(b200fft_plan** out, uint64_t len, int direction, int precision, int device) {
    if (!out) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if ((direction != B200FFT_FORWARD && direction != B200FFT_INVERSE) || (precision != B200FFT_F32 && precision != B200FFT_F64))
        return b2::fail(B200FFT_ERR_INVALID_ARG, "unknown direction or precision");
    const int ndev = b2::rt::device_count();
    if (ndev <= 0) return b2::fail(B200FFT_ERR_NO_DEVICE, "no sm_100 CUDA device is visible (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return b2::fail(B200FFT_ERR_INVALID_ARG, "device index out of range");
    if (!b2::rt::set_device(device)) return b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    std::unique_ptr<b200fft_plan> pl(new b200fft_plan());
    pl->len = len;
    pl->direction = direction;
    pl->precision = precision;
    pl->device = device;
    const int rc = precision == B200FFT_F32 ? b2::build_plan_f32(*pl) : b2::build_plan_f64(*pl);
    if (rc != B200FFT_OK) return rc;
    *out = pl.release();
    return B200FFT_OK;
}

int b200fft_plan_create_from_recipe(b200fft_plan** out, const b200fft_recipe_node* nodes, uint32_t n_nodes, int direction, int precision,
                                    int device) {
    if (!out) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (!nodes || n_nodes == 0 || n_nodes > 64) return b2::fail(B200FFT_ERR_INVALID_ARG, "recipe: 1..64 nodes expected");
    if ((direction != B200FFT_FORWARD && direction != B200FFT_INVERSE) || (precision != B200FFT_F32 && precision != B200FFT_F64))
        return b2::fail(B200FFT_ERR_INVALID_ARG, "unknown direction or precision");
    for (uint32_t i = 0; i < n_nodes; ++i)
        if (nodes[i].child >= n_nodes || (nodes[i].child != 0 && nodes[i].child <= i))
            return b2::fail(B200FFT_ERR_INVALID_ARG, "recipe: child indices must point forward inside the node array");
    const int ndev = b2::rt::device_count();
    if (ndev <= 0) return b2::fail(B200FFT_ERR_NO_DEVICE, "no sm_100 CUDA device is visible (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return b2::fail(B200FFT_ERR_INVALID_ARG, "device index out of range");
    if (!b2::rt::set_device(device)) return b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    std::unique_ptr<b200fft_plan> pl(new b200fft_plan());
    pl->len = nodes[0].len;
    pl->direction = direction;
    pl->precision = precision;
    pl->device = device;
    pl->recipe.assign(nodes, nodes + n_nodes);
    if (!b2::hm::is_pow2(pl->len) && pl->len > (1ull << 23) && nodes[0].kind != B200FFT_RECIPE_COLUMNS)
        return b2::fail(B200FFT_ERR_UNSUPPORTED, "non-power-of-two lengths above 2^23 are not planned by this build");
    const int rc = precision == B200FFT_F32 ? b2::build_plan_f32(*pl) : b2::build_plan_f64(*pl);
    if (rc != B200FFT_OK) return rc;
    *out = pl.release();
    return B200FFT_OK;
}

int b200fft_plan_destroy(b200fft_plan* plan) {
    if (!plan) return B200FFT_OK;
    b2::rt::set_device(plan->device);
    delete plan;
    return B200FFT_OK;
}

int b200fft_plan_recipe(const b200fft_plan* plan, b200fft_recipe_node* nodes, uint32_t cap) {
    if (!plan) return b2::fail(B200FFT_ERR_INVALID_ARG, "null plan");
    const uint32_t n = (uint32_t)plan->chosen.size();
    if (nodes)
        for (uint32_t i = 0; i < n && i < cap; ++i) nodes[i] = plan->chosen[i];
    return (int)n;
}

uint64_t b200fft_plan_len(const b200fft_plan* plan) { return plan ? plan->len : 0; }
int b200fft_plan_direction(const b200fft_plan* plan) { return plan ? plan->direction : -1; }
int b200fft_plan_precision(const b200fft_plan* plan) { return plan ? plan->precision : -1; }
uint64_t b200fft_plan_scratch_len(const b200fft_plan*, int) { return 0; }
uint64_t b200fft_plan_launches(const b200fft_plan* plan, uint64_t batch) { return plan ? plan->launches(batch) : 0; }

int b200fft_plan_describe(const b200fft_plan* plan, char* buf, uint64_t cap) {
    if (!plan || !buf || cap == 0) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    if (plan->desc.size() + 1 > cap) return b2::fail(B200FFT_ERR_INVALID_ARG, "buffer too small");
    std::memcpy(buf, plan->desc.c_str(), plan->desc.size() + 1);
    return (int)plan->desc.size();
}

int b200fft_exec_host_inplace(const b200fft_plan* plan, void* buffer, uint64_t n_complex) {
    if (!plan) return b2::fail(B200FFT_ERR_INVALID_ARG, "null plan");
    if (plan->len == 0) return B200FFT_OK;
    const int v = b2::validate_len(plan, n_complex, n_complex, false);
    if (v != B200FFT_OK) return v;
    return b2::exec_host_impl(plan, buffer, buffer, n_complex);
}

int b200fft_exec_host_outofplace(const b200fft_plan* plan, const void* input, void* output, uint64_t n_complex) {
    if (!plan) return b2::fail(B200FFT_ERR_INVALID_ARG, "null plan");
    if (plan->len == 0) return B200FFT_OK;
    const int v = b2::validate_len(plan, n_complex, n_complex, true);
    if (v != B200FFT_OK) return v;
    return b2::exec_host_impl(plan, input, output, n_complex);
}

int b200fft_exec_device(const b200fft_plan* plan, const void* d_in, void* d_out, uint64_t batch, void* cuda_stream) {
    return b2::exec_device_impl(plan, d_in, d_out, batch, (b2::rt::stream_t)cuda_stream, nullptr, 0, false);
}

uint64_t b200fft_workspace_bytes(const b200fft_plan* plan, uint64_t batch) { return plan ? plan->work_bytes(batch) : 0; }

int b200fft_exec_device_ws(const b200fft_plan* plan, const void* d_in, void* d_out, uint64_t batch, void* cuda_stream,
                           void* d_workspace, uint64_t workspace_bytes) {
    return b2::exec_device_impl(plan, d_in, d_out, batch, (b2::rt::stream_t)cuda_stream, d_workspace, workspace_bytes, true);
}

}  // extern "C"

// ---- real-input / real-output wrappers (real.h) ----------------------------------------------------------------------------
struct b200fft_real_plan {
    uint64_t len = 0, M = 0;
    int precision = 0, device = 0;
    b200fft_plan* fwd = nullptr;  // M-point complex plans
    b200fft_plan* inv = nullptr;
    void* tw = nullptr;           // W_len^k, k = 0 .. M/2
    ~b200fft_real_plan() {
        if (fwd) b200fft_plan_destroy(fwd);
        if (inv) b200fft_plan_destroy(inv);
        if (tw) b2::rt::dfree(tw);
    }
};

namespace b2 {
template <typename T>
static bool real_pack_launch(const b200fft_real_plan* rp, int dir, const void* in, void* out, uint64_t batch, rt::stream_t s) {
    const uint32_t h = (uint32_t)(rp->M / 2 + 1);
    if (dir == 0) {
        typename RealPackKernel<T, 0>::Params p{(const cx<T>*)in, (cx<T>*)out, (const cx<T>*)rp->tw, batch * h, (uint32_t)rp->M, make_fastdiv(h)};
        return rt::launch<RealPackKernel<T, 0>>(p, (p.n_pairs + 255) / 256, s);
    }
    typename RealPackKernel<T, 1>::Params p{(const cx<T>*)in, (cx<T>*)out, (const cx<T>*)rp->tw, batch * h, (uint32_t)rp->M, make_fastdiv(h)};
    return rt::launch<RealPackKernel<T, 1>>(p, (p.n_pairs + 255) / 256, s);
}
// forward: real (as M complex) --FFT_M--> work --unpack--> out;   inverse: in --pack--> work --IFFT_M--> real out (as M complex)
static int real_exec_device(const b200fft_real_plan* rp, bool inverse, const void* d_in, void* d_out, uint64_t batch, rt::stream_t stream) {
    if (!rp || !d_in || !d_out) return fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    if (batch == 0) return B200FFT_OK;
    if (batch * (rp->M / 2 + 1) >= (1ull << 31)) return fail(B200FFT_ERR_UNSUPPORTED, "real transforms: batch * (len/4 + 1) must stay below 2^31 per call");
    rt::DeviceGuard guard(rp->device);
    if (!guard.ok) return fail(B200FFT_ERR_CUDA, rt::last_error());
    const uint64_t esz = rp->precision == B200FFT_F32 ? 8 : 16;
    const uint64_t zbytes = (batch * rp->M * esz + 255) / 256 * 256;
    const b200fft_plan* cp = inverse ? rp->inv : rp->fwd;
    const uint64_t inner = cp->work_bytes(batch);
    char* work = (char*)rt::malloc_async(zbytes + inner + 256, stream);
    if (!work) return fail(B200FFT_ERR_CUDA, "workspace allocation failed: " + rt::last_error());
    void* z = work;
    void* iw = inner ? work + zbytes : nullptr;
    int rc = B200FFT_OK;
    bool ok = true;
    if (!inverse) {
        rc = exec_device_impl(cp, d_in, z, batch, stream, iw, inner, inner != 0);
        if (rc == B200FFT_OK) ok = rp->precision == B200FFT_F32 ? real_pack_launch<float>(rp, 0, z, d_out, batch, stream) : real_pack_launch<double>(rp, 0, z, d_out, batch, stream);
    } else {
        ok = rp->precision == B200FFT_F32 ? real_pack_launch<float>(rp, 1, d_in, z, batch, stream) : real_pack_launch<double>(rp, 1, d_in, z, batch, stream);
        if (ok) rc = exec_device_impl(cp, z, d_out, batch, stream, iw, inner, inner != 0);
    }
    rt::free_async(work, stream);
    if (rc != B200FFT_OK) return rc;
    if (!ok) return fail(B200FFT_ERR_CUDA, "kernel launch failed: " + rt::last_error());
    return B200FFT_OK;
}
static int real_exec_host(const b200fft_real_plan* rp, bool inverse, const void* in, void* out, uint64_t batch) {
    if (!rp || !in || !out) return fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    if (batch == 0) return B200FFT_OK;
    rt::DeviceGuard guard(rp->device);
    if (!guard.ok) return fail(B200FFT_ERR_CUDA, rt::last_error());
    const uint64_t esz = rp->precision == B200FFT_F32 ? 8 : 16;
    const uint64_t rbytes = batch * rp->M * esz, cbytes = batch * (rp->M + 1) * esz;
    const uint64_t ib = inverse ? cbytes : rbytes, ob = inverse ? rbytes : cbytes;
    void* d_in = rt::dmalloc(ib);
    void* d_out = rt::dmalloc(ob);
    rt::stream_t s = rt::stream_create();
    int rc = (d_in && d_out && s) ? B200FFT_OK : fail(B200FFT_ERR_CUDA, "staging allocation failed: " + rt::last_error());
    if (rc == B200FFT_OK && !rt::h2d_async(d_in, in, ib, s)) rc = fail(B200FFT_ERR_CUDA, rt::last_error());
    if (rc == B200FFT_OK) rc = real_exec_device(rp, inverse, d_in, d_out, batch, s);
    if (rc == B200FFT_OK && !(rt::d2h_async(out, d_out, ob, s) && rt::stream_sync(s))) rc = fail(B200FFT_ERR_CUDA, rt::last_error());
    if (s) {
        rt::stream_sync(s);
        rt::stream_destroy(s);
    }
    if (d_in) rt::dfree(d_in);
    if (d_out) rt::dfree(d_out);
    return rc;
}
}  // namespace b2

extern "C" {

int b200fft_real_plan_create(b200fft_real_plan** out, uint64_t len, int precision, int device) {
    if (!out) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (len < 2 || (len & 1)) return b2::fail(B200FFT_ERR_UNSUPPORTED, "real transforms need an even length >= 2");
    if (precision != B200FFT_F32 && precision != B200FFT_F64) return b2::fail(B200FFT_ERR_INVALID_ARG, "unknown precision");
    std::unique_ptr<b200fft_real_plan> rp(new b200fft_real_plan());
    rp->len = len;
    rp->M = len / 2;
    rp->precision = precision;
    rp->device = device;
    int rc = b200fft_plan_create(&rp->fwd, rp->M, B200FFT_FORWARD, precision, device);
    if (rc == B200FFT_OK) rc = b200fft_plan_create(&rp->inv, rp->M, B200FFT_INVERSE, precision, device);
    if (rc != B200FFT_OK) return rc;
    const uint64_t h = rp->M / 2 + 1;
    if (precision == B200FFT_F32) {
        std::vector<b2::cx<float>> t((size_t)h);
        for (uint64_t k = 0; k < h; ++k) t[(size_t)k] = b2::hm::twiddle<float>(k, len);
        rp->tw = b2::rt::dmalloc(h * 8);
        if (!rp->tw || !b2::rt::h2d_sync(rp->tw, t.data(), h * 8)) return b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    } else {
        std::vector<b2::cx<double>> t((size_t)h);
        for (uint64_t k = 0; k < h; ++k) t[(size_t)k] = b2::hm::twiddle<double>(k, len);
        rp->tw = b2::rt::dmalloc(h * 16);
        if (!rp->tw || !b2::rt::h2d_sync(rp->tw, t.data(), h * 16)) return b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    }
    *out = rp.release();
    return B200FFT_OK;
}
int b200fft_real_plan_destroy(b200fft_real_plan* plan) {
    if (!plan) return B200FFT_OK;
    b2::rt::set_device(plan->device);
    delete plan;
    return B200FFT_OK;
}
uint64_t b200fft_real_workspace_bytes(const b200fft_real_plan* plan, uint64_t batch) {
    if (!plan) return 0;
    const uint64_t esz = plan->precision == B200FFT_F32 ? 8 : 16;
    return (batch * plan->M * esz + 255) / 256 * 256 + std::max(plan->fwd->work_bytes(batch), plan->inv->work_bytes(batch)) + 256;
}
int b200fft_real_forward_device(const b200fft_real_plan* plan, const void* d_real_in, void* d_complex_out, uint64_t batch, void* cuda_stream) {
    return b2::real_exec_device(plan, false, d_real_in, d_complex_out, batch, (b2::rt::stream_t)cuda_stream);
}
int b200fft_real_inverse_device(const b200fft_real_plan* plan, const void* d_complex_in, void* d_real_out, uint64_t batch, void* cuda_stream) {
    return b2::real_exec_device(plan, true, d_complex_in, d_real_out, batch, (b2::rt::stream_t)cuda_stream);
}
int b200fft_real_forward_host(const b200fft_real_plan* plan, const void* real_in, void* complex_out, uint64_t batch) {
    return b2::real_exec_host(plan, false, real_in, complex_out, batch);
}
int b200fft_real_inverse_host(const b200fft_real_plan* plan, const void* complex_in, void* real_out, uint64_t batch) {
    return b2::real_exec_host(plan, true, complex_in, real_out, batch);
}

}  // extern "C"

// ---- 2-D plans: rows through the width-point plan, then one strided column pass ---------------------------------------------
struct b200fft_plan2d {
    uint64_t H = 0, W = 0;
    int precision = 0, device = 0;
    b200fft_plan* rows = nullptr;
    b200fft_plan* cols = nullptr;
    ~b200fft_plan2d() {
        if (rows) b200fft_plan_destroy(rows);
        if (cols) b200fft_plan_destroy(cols);
    }
};

extern "C" {

int b200fft_plan2d_create(b200fft_plan2d** out, uint64_t height, uint64_t width, int direction, int precision, int device) {
    if (!out) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (height < 1 || width < 1) return b2::fail(B200FFT_ERR_INVALID_ARG, "2-D plans need height >= 1 and width >= 1");
    std::unique_ptr<b200fft_plan2d> p2(new b200fft_plan2d());
    p2->H = height;
    p2->W = width;
    p2->precision = precision;
    p2->device = device;
    int rc = b200fft_plan_create(&p2->rows, width, direction, precision, device);
    if (rc != B200FFT_OK) return rc;
    if (height > 1) {
        const b200fft_recipe_node node{B200FFT_RECIPE_COLUMNS, 0u, height * width, height, width};
        rc = b200fft_plan_create_from_recipe(&p2->cols, &node, 1, direction, precision, device);
        if (rc != B200FFT_OK) return rc;
    }
    *out = p2.release();
    return B200FFT_OK;
}
int b200fft_plan2d_destroy(b200fft_plan2d* plan) {
    delete plan;
    return B200FFT_OK;
}
int b200fft_exec2d_device(const b200fft_plan2d* plan, const void* d_in, void* d_out, uint64_t batch, void* cuda_stream) {
    if (!plan) return b2::fail(B200FFT_ERR_INVALID_ARG, "null plan");
    int rc = b200fft_exec_device(plan->rows, d_in, d_out, batch * plan->H, cuda_stream);
    if (rc == B200FFT_OK && plan->cols) rc = b200fft_exec_device(plan->cols, d_out, d_out, batch, cuda_stream);
    return rc;
}
int b200fft_exec2d_host(const b200fft_plan2d* plan, const void* in, void* out, uint64_t batch) {
    if (!plan || !in || !out) return b2::fail(B200FFT_ERR_INVALID_ARG, "null pointer");
    if (batch == 0) return B200FFT_OK;
    b2::rt::DeviceGuard guard(plan->device);
    if (!guard.ok) return b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    const uint64_t bytes = batch * plan->H * plan->W * (plan->precision == B200FFT_F32 ? 8 : 16);
    void* d = b2::rt::dmalloc(bytes);
    b2::rt::stream_t s = b2::rt::stream_create();
    int rc = (d && s) ? B200FFT_OK : b2::fail(B200FFT_ERR_CUDA, "staging allocation failed: " + b2::rt::last_error());
    if (rc == B200FFT_OK && !b2::rt::h2d_async(d, in, bytes, s)) rc = b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    if (rc == B200FFT_OK) rc = b200fft_exec2d_device(plan, d, d, batch, s);
    if (rc == B200FFT_OK && !(b2::rt::d2h_async(out, d, bytes, s) && b2::rt::stream_sync(s))) rc = b2::fail(B200FFT_ERR_CUDA, b2::rt::last_error());
    if (s) {
        b2::rt::stream_sync(s);
        b2::rt::stream_destroy(s);
    }
    if (d) b2::rt::dfree(d);
    return rc;
}

const char* b200fft_last_error(void) { return b2::g_last_error.c_str(); }
const char* b200fft_version(void) { return "b200fft 0.1 sm_100a"; }

}  // extern "C"
#endif  // B2_PART_CABI
