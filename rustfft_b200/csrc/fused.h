// Fused single-launch four-step: ONE persistent, warp-specialised kernel runs both passes of every transform of a batch.
//
// Why: the chunked path (one launch pair per L2 chunk, kernels.h run_kernel_tma) is latency bound -- a tile's load, its
// butterflies and its store are strictly serial inside a CTA, an SM holds two such tiles, and every ~10 us launch pays a
// ramp and a tail (round-1 ncu: 17 % occupancy, DRAM ~50 % busy, 0.48-0.57 of the HBM roofline).  Here every SM runs one
// resident CTA made of
//     1 producer thread   draws tile tickets from a global counter, waits for the ticket's dependency (always a
//                         SMALLER ticket), and queues the tile's TMA load into the next free shared-memory stage;
//     NG consumer groups  (NTG threads each, one tile at a time): wait for the stage's mbarrier, run the tile's FFTs in
//                         registers + in place in the stage buffer (the phases of TmaTileKernel), leave the finished
//                         dense tile in the same buffer;
//     1 storer thread     TMA-stores finished tiles, frees the stage once the store has read it, publishes the
//                         "pass-A tile landed" counters;
// so the load of tile i+2, the butterflies of tiles i and i+1 and the store of tile i-1 overlap inside one SM, there
// are no launches between tiles, and the device mixes HBM reads (pass-A tiles) with HBM writes (pass-B tiles) at tile
// granularity.  Ticket order = FlowSched (kernels.h): round r holds the pass-A tiles of transform r interleaved with
// the pass-B tiles of transform r - D; the intermediate lives in a ring of W = 2 D transform slots that stays in L2
// (pass B drops the lines it has consumed with discard.global.L2, so they are never written back to HBM).
//   B(t) may start when all TA tiles of A(t) have landed    (ready[slot] >= (gen + 1) * TA)
//   A(t) may start when all TB tiles of B(t - W) were read  (freed[slot] >= gen * TB)
// Both counters only grow; a dependency always points to a smaller ticket and a CTA's consumers / storer never wait
// for its producer, so the smallest unfinished ticket can always proceed: no deadlock whatever the co-residency.
// The reference's shape for the same job: MixedRadix's six steps (src/algorithm/mixed_radix.rs:128-158).
#pragma once
#include "kernels.h"

namespace b2 {

template <class KA, class KB, int NG_, int NS_>
struct FusedKernel {
    using T = typename KA::T;
    static constexpr int NG = NG_;      // consumer groups
    static constexpr int NSTAGE = NS_;  // shared-memory stages (tiles in flight per SM)
    static constexpr int NTG = KA::NT;  // threads per consumer group
    static_assert(KA::NT == KB::NT, "both passes use the same consumer-group size");
    static_assert(KA::TILE_BYTES == KB::TILE_BYTES, "both passes move tiles of the same size");
    static_assert(NTG % 32 == 0 && NSTAGE <= 8 && NG <= 8, "geometry");
    static constexpr int NT = NG * NTG + 64;  // + producer warp + storer warp
    static constexpr size_t STAGE_BYTES = ((KA::SMEM_BYTES > KB::SMEM_BYTES ? KA::SMEM_BYTES : KB::SMEM_BYTES) + 127) / 128 * 128;
    static constexpr size_t CTRL_BYTES = 512;  // 4 * NSTAGE mbarriers, NSTAGE x {kind, tile, slot, -}
    static constexpr size_t SMEM_BYTES = (size_t)NSTAGE * STAGE_BYTES + CTRL_BYTES + 128;  // + alignment slack
    struct Params {
        typename KA::Params a;
        typename KB::Params b;
        FlowSched sched;
        uint32_t* ctl;    // control block (FLOW_CTL_HEAD layout of kernels.h), zeroed before the launch
        uint32_t flags;   // bit 0: skip the butterflies (memory-pipeline ceiling measurement, results are garbage)
    };
};

#if defined(__CUDACC__)
B2_D uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *ctr >= target (bounded: a scheduling bug raises the error flag instead of hanging the GPU)
B2_D void fused_spin(uint32_t* ctl, const uint32_t* ctr, uint32_t target) {
    uint32_t spins = 0;
    while (ld_acquire_u32(ctr) < target) {
        __nanosleep(spins < 64 ? 20 : 200);
        if (++spins > (1u << 22) || (spins > 4096 && ld_relaxed_u32(ctl + 1) != 0)) {
            atomicExch(ctl + 1, 1u);
            break;
        }
    }
}

// the phases of one tile inside a consumer group (named barrier `bar_id` among KT::NT threads); the last phase of
// TmaTileKernel (the store) belongs to the storer thread.  `freed`: pass-B tiles publish "ring slot rows consumed" once
// every thread holds its inputs and has issued its discards (the barrier after phase 1).
template <class KT, int P>
struct GroupPhases {
    static B2_D void run(const typename KT::Params& p, uint32_t bid, int ltid, typename KT::Regs& r, cx<typename KT::T>* buf, int bar_id,
                         uint32_t* freed) {
        KT::template phase<P>(p, bid, ltid, r, buf);
        if constexpr (P + 2 < KT::NPHASE) {
            tma::named_bar_sync(bar_id, KT::NT);
            if constexpr (P == 1) {
                if (freed != nullptr && ltid == 0) {
                    __threadfence();
                    atomicAdd(freed, 1u);
                }
            }
            GroupPhases<KT, P + 1>::run(p, bid, ltid, r, buf, bar_id, freed);
        }
    }
};

template <class KA, class KB, int NG, int NS>
__global__ void __launch_bounds__(FusedKernel<KA, KB, NG, NS>::NT, 1)
run_fused(const __grid_constant__ typename FusedKernel<KA, KB, NG, NS>::Params p) {
    using FK = FusedKernel<KA, KB, NG, NS>;
    using C = cx<typename FK::T>;
    constexpr int NTG = FK::NTG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)NS * FK::STAGE_BYTES);
    uint64_t* full = bars;            // tile landed (producer arrival + TMA bytes)
    uint64_t* meta = bars + NS;       // tile description written (consumers may prefetch their tables)
    uint64_t* outf = bars + 2 * NS;   // finished tile in the buffer (one arrival per consumer warp)
    uint64_t* empty = bars + 3 * NS;  // the store has read the buffer (storer)
    volatile uint32_t* info = reinterpret_cast<volatile uint32_t*>(bars + 4 * NS);  // [stage][4]: kind (0 A, 1 B, 2 end), tile, slot
    const int tid = (int)threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const FlowSched& sc = p.sched;
    uint32_t* ready = p.ctl + FLOW_CTL_HEAD;
    uint32_t* freed = ready + sc.ring_w;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            tma::mbar_init(&full[s], 1);
            tma::mbar_init(&meta[s], 1);
            tma::mbar_init(&outf[s], NTG / 32);
            tma::mbar_init(&empty[s], 1);
        }
        tma::fence_mbar_init();
    }
    __syncthreads();
    auto stage_buf = [&](uint32_t s) { return reinterpret_cast<C*>(base + (size_t)s * FK::STAGE_BYTES); };

    if (warp == NG * NTG / 32) {
        // ---------------- producer ----------------
        if (lane == 0) {
            uint32_t i = 0;
            uint32_t ticket = atomicAdd(p.ctl, 1u);
            while (ticket < sc.total) {
                int kind;
                uint32_t t, tile;
                bool valid;
                sc.decode(ticket, kind, t, tile, valid);
                if (!valid) {
                    ticket = atomicAdd(p.ctl, 1u);
                    continue;
                }
                const uint32_t s = i % NS, ph = (i / NS) & 1u;
                tma::mbar_wait(&empty[s], ph ^ 1u);  // (passes at once for the first NS tiles)
                const uint32_t slot = t % sc.ring_w;
                const uint32_t bid = kind == 0 ? t * sc.TA + tile : t * sc.TB + tile;
                info[4 * s + 0] = (uint32_t)kind;
                info[4 * s + 1] = bid;
                info[4 * s + 2] = slot;
                tma::mbar_arrive(&meta[s]);
                const FlowDep d = flow_dep(sc, p.ctl, ticket);
                if (d.ptr != nullptr) fused_spin(p.ctl, d.ptr, d.target);
                tma::fence_proxy_async_all();  // the acquire above -> the TMA reads / (later) writes of the slot
                if (kind == 0)
                    KA::issue_load(p.a, bid, stage_buf(s), &full[s]);
                else
                    KB::issue_load(p.b, bid, stage_buf(s), &full[s]);
                ticket = atomicAdd(p.ctl, 1u);  // its round trip overlaps the tile's flight
                ++i;
            }
            for (int g = 0; g < NG; ++g, ++i) {  // one end marker per consumer group
                const uint32_t s = i % NS, ph = (i / NS) & 1u;
                tma::mbar_wait(&empty[s], ph ^ 1u);
                info[4 * s + 0] = 2u;
                tma::mbar_arrive(&meta[s]);
            }
        }
    } else if (warp == NG * NTG / 32 + 1) {
        // ---------------- storer ----------------
        if (lane == 0) {
            uint32_t* pending = nullptr;  // ready counter of the last pass-A tile stored, not yet published
            for (uint32_t i = 0;; ++i) {
                const uint32_t s = i % NS, ph = (i / NS) & 1u;
                if (pending != nullptr && !tma::mbar_test(&outf[s], ph)) {
                    // about to idle: other CTAs (or this CTA's own producer) may be waiting for that tile
                    tma::bulk_wait<0>();
                    tma::fence_proxy_async_all();
                    __threadfence();
                    atomicAdd(pending, 1u);
                    pending = nullptr;
                }
                tma::mbar_wait(&outf[s], ph);
                const uint32_t kind = info[4 * s + 0], bid = info[4 * s + 1], slot = info[4 * s + 2];
                if (kind == 2u) break;
                if (kind == 0u)
                    KA::issue_store(p.a, bid, stage_buf(s));
                else
                    KB::issue_store(p.b, bid, stage_buf(s));
                tma::bulk_commit();
                if (pending != nullptr) {  // every group but the one just committed has completed
                    tma::bulk_wait<1>();
                    tma::fence_proxy_async_all();
                    __threadfence();
                    atomicAdd(pending, 1u);
                    pending = nullptr;
                }
                tma::bulk_wait_read<0>();  // the buffer may be refilled
                tma::mbar_arrive(&empty[s]);
                if (kind == 0u) pending = ready + slot;
            }
            tma::bulk_wait<0>();
            if (pending != nullptr) {
                tma::fence_proxy_async_all();
                __threadfence();
                atomicAdd(pending, 1u);
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int g = warp / (NTG / 32);
        const int ltid = tid - g * NTG;
        const int bar_id = 1 + g;
        for (uint32_t i = (uint32_t)g;; i += NG) {
            const uint32_t s = i % NS, ph = (i / NS) & 1u;
            tma::mbar_wait(&meta[s], ph);
            const uint32_t kind = info[4 * s + 0], bid = info[4 * s + 1], slot = info[4 * s + 2];
            C* buf = stage_buf(s);
            if (kind == 0u) {
                typename KA::Regs r;
                KA::prefetch(p.a, bid, ltid, r);  // table loads overlap the tile's flight
                tma::mbar_wait(&full[s], ph);
                if (!(p.flags & 1u)) GroupPhases<KA, 0>::run(p.a, bid, ltid, r, buf, bar_id, nullptr);
                else tma::fence_proxy_async();
            } else if (kind == 1u) {
                typename KB::Regs r;
                KB::prefetch(p.b, bid, ltid, r);
                tma::mbar_wait(&full[s], ph);
                if (!(p.flags & 1u)) GroupPhases<KB, 0>::run(p.b, bid, ltid, r, buf, bar_id, freed + slot);
                else {
                    tma::fence_proxy_async();
                    if (ltid == 0) atomicAdd(freed + slot, 1u);
                }
            }
            // every thread has written its share of the dense output tile and fenced it towards the async proxy
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&outf[s]);
            if (kind == 2u) break;
        }
    }
}
#endif

}  // namespace b2
