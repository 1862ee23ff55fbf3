// Single-pass four-step on a thread-block CLUSTER: one transform of N = C * 8192 points (f32) lives in the shared memory and
// registers of C CTAs; the transpose between the column pass and the row pass goes through DISTRIBUTED shared memory (DSMEM)
// instead of through L2, so the signal crosses HBM once in and once out and touches L2 only on the way.
//
// Why: the two-pass plans (chunked launch pairs, or the fused persistent kernel of fused.h) write the intermediate to L2 and read
// it back.  Measured with membw / tmabw (profiles/r2_membw.txt, r2_tmabw.txt): a kernel that streams D bytes HBM -> SM -> L2 and
// D bytes L2 -> SM -> HBM at the same time tops out at ~0.71 of the HBM copy roofline -- the L2 slices carry 4 D of SM-side traffic
// plus 2 D of DRAM fills / write-backs -- and the real kernels reach 0.55-0.60.  Keeping the intermediate on chip removes the cap.
//
//   CTA c of a cluster (rank c of C) owns          columns n2 in [c W, (c+1) W), W = N2 / C      during the column pass
//                                                   rows    k1 in [c H, (c+1) H), H = N1 / C      during the row pass
//   phase A   tile [N1 rows x W columns] -> registers (coalesced 8-byte loads, "f fastest") -> N1-point FFTs (CTA engine, 32
//             elements per thread, one exchange through the CTA's own shared memory)
//   barrier.cluster #1: every CTA of the cluster has finished READING its exchange buffer
//   exchange  thread (column f, slot q) holds Y[k1 = j + TP q][n2 = c W + f]; it stores it into the row-pass tile of the CTA that owns
//             row k1 -- rank k1 / H, element (k1 mod H) * N2 + n2 -- with a plain store through the cluster's shared-memory window
//             (mapa / st.shared::cluster, here via cooperative_groups::cluster_group::map_shared_rank).  One warp instruction writes 256
//             contiguous bytes of ONE remote CTA.
//   barrier.cluster #2: all remote stores have landed
//   phase B   tile [H rows x N2 columns] x W_N^(k1 n2) -> N2-point FFTs -> X[k1 + N1 k2] (runs of H consecutive elements)
// Same maths as the two passes of FourStep (kernels.h: LoadCols / LoadRowsTw / StoreTransposed), the reference's six-step
// MixedRadix (src/algorithm/mixed_radix.rs:128-158) with all three transposes folded into loads, DSMEM stores and stores.
#pragma once
#include "kernels.h"

#if defined(__CUDACC__)
#include <cooperative_groups.h>
#endif

namespace b2 {

template <class GA_, class GB_, int C_, bool SW>
struct ClusterKernel {
    using GA = GA_;
    using GB = GB_;
    using T = typename GA::T;
    using EA = Engine<GA, FF, FF>;
    using EB = Engine<GB, JF, FF>;
    static constexpr int C = C_;
    static constexpr int N1 = GA::L, N2 = GB::L;
    static constexpr int W = GA::F;  // columns per CTA
    static constexpr int H = GB::F;  // rows per CTA
    static constexpr int NT = GA::NT;
    static_assert(GA::NT == GB::NT, "both passes use the CTA's threads");
    static_assert(W * C == N2 && H * C == N1, "the cluster covers the transform");
    static_assert(GA::E == GB::E && H % GA::TP == 0, "a thread's slots of one exchange store go to one remote CTA");
    static constexpr int MIN_BLOCKS = 512 / GA::NT;  // 128 registers per thread
    static constexpr int NPA = EA::NPHASE, NPB = EB::NPHASE;
    static constexpr int NPHASE = NPA + 2 + NPB;
    // CTA-local buffer: the engines' padded exchange buffers and the dense row-pass tile share it
    static constexpr size_t TILE_ELEMS = (size_t)H * N2;
    static constexpr size_t BUF_ELEMS = (((size_t)GA::SMEM_ELEMS > (size_t)GB::SMEM_ELEMS ? (size_t)GA::SMEM_ELEMS : (size_t)GB::SMEM_ELEMS) > TILE_ELEMS
                                             ? ((size_t)GA::SMEM_ELEMS > (size_t)GB::SMEM_ELEMS ? (size_t)GA::SMEM_ELEMS : (size_t)GB::SMEM_ELEMS)
                                             : TILE_ELEMS);
    static constexpr size_t SMEM_BYTES = (BUF_ELEMS * sizeof(cx<T>) + 15) / 16 * 16;
    // barrier that follows phase P: 1 = CTA, 2 = cluster
    static constexpr int barrier_after(int P) { return (P == NPA - 1 || P == NPA) ? 2 : 1; }

    struct Params {
        LoadCols<T, SW> load;            // in[b N + e N2 + c]
        StoreTransposed<T, SW> store;    // out[b N + k1 + N1 e]
        const cx<T>* twa;                // stage twiddles of the N1-point FFT
        const cx<T>* twb;                // stage twiddles of the N2-point FFT
        const cx<T>* full_tw;            // W_N^(k1 n2), [k1][n2]
        uint64_t n_transforms;
    };
    struct Regs { cx<T> v[GA::E]; };

    // P < NPA: column pass;  P == NPA: exchange (remote[r] = the row-pass tile of rank r);  P > NPA: row pass
    template <int P>
    static B2_HD void phase(const Params& p, uint32_t bid, int tid, Regs& r, cx<T>* buf, cx<T>* const* remote) {
        const uint32_t rank = bid % C;
        if constexpr (P < NPA) {
            if constexpr (P == 0) {
                int f, j;
                EA::template owner<0>(tid, f, j);
                const auto st = p.load.prep((uint64_t)bid * W + f, true);  // column (bid / C) * N2 + rank * W + f
                B2_UNROLL
                for (int q = 0; q < GA::E; ++q) r.v[q] = p.load.get(st, j + GA::TP * q);
            }
            EA::template phase<P>(tid, r.v, buf, p.twa);
        } else if constexpr (P == NPA) {
            int f, j;
            EA::out_owner(tid, f, j);
            constexpr int QPR = H / GA::TP;  // slots per destination rank
            B2_UNROLL
            for (int q = 0; q < GA::E; ++q) {
                cx<T>* dst = remote[q / QPR] + (size_t)(GA::TP * (q % QPR) + j) * N2 + rank * W + f;
                *dst = r.v[q];
            }
        } else if constexpr (P == NPA + 1) {
            // dense row-pass tile -> registers, times the inter-pass twiddles (1 + log2 E table entries per thread, the rest as products)
            int f, j;
            EB::template owner<0>(tid, f, j);
            const cx<T>* src = buf + (size_t)f * N2 + j;
            const cx<T>* t = p.full_tw + ((size_t)rank * H + f) * N2;
            cx<T> wq[GB::E];
            const cx<T> a = ldg_stream(t + j);
            B2_UNROLL
            for (int q = 1; q < GB::E; q <<= 1) wq[q] = ldg_stream(t + GB::TP * q);
            B2_UNROLL
            for (int q = 3; q < GB::E; ++q)
                if (q & (q - 1)) wq[q] = cmul(wq[hibit(q)], wq[q - hibit(q)]);
            r.v[0] = cmul(src[0], a);
            B2_UNROLL
            for (int q = 1; q < GB::E; ++q) r.v[q] = cmul(src[GB::TP * q], cmul(a, wq[q]));
        } else {
            constexpr int PB = P - NPA - 2;
            EB::template phase<PB>(tid, r.v, buf, p.twb);
            if constexpr (PB == NPB - 1) {
                int f, j;
                EB::out_owner(tid, f, j);
                const auto st = p.store.prep((uint64_t)bid * H + f, true);  // row (bid / C) * N1 + rank * H + f
                B2_UNROLL
                for (int q = 0; q < GB::E; ++q) p.store.put(st, j + GB::TP * q, r.v[q]);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Convolution plans inside ONE cluster pass: Rader (prime n = M + 1) and Bluestein (2n - 1 <= M) over an inner FFT of M = L * L
// points held by the cluster (L = 128: C = 2, M = 2^14;  L = 256: C = 8, M = 2^16 -- BASELINE config 4, n = 65537).
//   gather | chirp-pad load -> FFT_M (column pass, DSMEM exchange, row pass) -> x mult, conj (+ Rader DC) IN REGISTERS
//   -> FFT_M again -> conj + scatter | conj x chirp store
// The second FFT needs no exchange on the way in: with N1 = N2 the row pass leaves thread (row k1, slot q <-> k2 = j + TP q) holding
// inner index k1 + L k2 = element (row k2, column k1) of the next column pass -- the registers of the SAME thread in the same
// slots.  So the whole algorithm of src/algorithm/raders_algorithm.rs:235-283 / bluesteins_algorithm.rs:100-136 is one read and one
// write of the signal (plus L2-resident tables), against four passes over M-sized intermediates in make_big_conv.
// ------------------------------------------------------------------------------------------
template <class G_, int C_, bool SW, int MODE>  // MODE 0 = Rader, 1 = Bluestein
struct ClusterConvKernel {
    using G = G_;
    using T = typename G::T;
    using EA = Engine<G, FF, FF>;
    using EB = Engine<G, JF, FF>;
    static constexpr int C = C_;
    static constexpr int L = G::L;   // N1 = N2
    static constexpr int W = G::F;   // columns (then rows) per CTA
    static constexpr int NT = G::NT;
    static_assert(W * C == L, "the cluster covers the inner FFT");
    static_assert(W % G::TP == 0, "a thread's slots of one exchange store go to one remote CTA");
    static constexpr int MIN_BLOCKS = 2;
    static constexpr int NPA = EA::NPHASE, NPB = EB::NPHASE;
    static constexpr int NP1 = NPA + 2 + NPB;  // phases of one inner FFT
    static constexpr int NPHASE = 2 * NP1;
    static constexpr size_t TILE_ELEMS = (size_t)W * L;
    static constexpr size_t BUF_ELEMS = (size_t)G::SMEM_ELEMS > TILE_ELEMS ? (size_t)G::SMEM_ELEMS : TILE_ELEMS;
    static constexpr size_t SMEM_BYTES = (BUF_ELEMS * sizeof(cx<T>) + 15) / 16 * 16;
    static constexpr int barrier_after(int P) {
        const int q = P % NP1;
        return (q == NPA - 1 || q == NPA) ? 2 : 1;
    }
    struct Params {
        const cx<T>* in;
        cx<T>* out;
        const uint32_t* gather;   // Rader: g^(i+1) mod n
        const uint32_t* scatter;  // Rader: g^-(i+1) mod n
        const cx<T>* chirp;       // Bluestein: n entries
        const cx<T>* mult;        // M entries
        const cx<T>* tw;          // stage twiddles of the L-point FFT
        const cx<T>* full_tw;     // W_M^(k1 n2), [k1][n2]
        uint32_t n;               // outer length = stride between transforms
        uint64_t n_transforms;
    };
    struct Regs { cx<T> v[G::E]; };

    template <int P>
    static B2_HD void phase(const Params& p, uint32_t bid, int tid, Regs& r, cx<T>* buf, cx<T>* const* remote) {
        const uint32_t rank = bid % C;
        const uint64_t b = bid / C;
        constexpr int Q = P % NP1;
        constexpr bool second = P >= NP1;
        if constexpr (Q < NPA) {
            if constexpr (Q == 0 && !second) {
                // inner element i = e L + c (row e = j + TP q, column c = rank W + f) of the first FFT's input
                int f, j;
                EA::template owner<0>(tid, f, j);
                const cx<T>* src = p.in + b * (uint64_t)p.n;
                const uint32_t c = rank * W + f;
                B2_UNROLL
                for (int q = 0; q < G::E; ++q) {
                    const uint32_t i = (uint32_t)(j + G::TP * q) * L + c;
                    cx<T> v = mk<T>(0, 0);
                    if (MODE == 0) {
                        v = src[ldg_u32(p.gather + i)];
                        if (SW) v = swap_ri(v);
                    } else if (i < p.n) {
                        v = ld_stream(src + i);
                        if (SW) v = swap_ri(v);
                        v = cmul(v, ldg(p.chirp + i));
                    }
                    r.v[q] = v;
                }
            }
            EA::template phase<Q>(tid, r.v, buf, p.tw);
        } else if constexpr (Q == NPA) {
            int f, j;
            EA::out_owner(tid, f, j);
            constexpr int QPR = W / G::TP;  // slots per destination rank
            B2_UNROLL
            for (int q = 0; q < G::E; ++q) {
                cx<T>* dst = remote[q / QPR] + (size_t)(G::TP * (q % QPR) + j) * L + rank * W + f;
                *dst = r.v[q];
            }
        } else if constexpr (Q == NPA + 1) {
            int f, j;
            EB::template owner<0>(tid, f, j);
            const cx<T>* src = buf + (size_t)f * L + j;
            const cx<T>* t = p.full_tw + ((size_t)rank * W + f) * L;
            cx<T> wq[G::E];
            const cx<T> a = ldg_stream(t + j);
            B2_UNROLL
            for (int q = 1; q < G::E; q <<= 1) wq[q] = ldg_stream(t + G::TP * q);
            B2_UNROLL
            for (int q = 3; q < G::E; ++q)
                if (q & (q - 1)) wq[q] = cmul(wq[hibit(q)], wq[q - hibit(q)]);
            r.v[0] = cmul(src[0], a);
            B2_UNROLL
            for (int q = 1; q < G::E; ++q) r.v[q] = cmul(src[G::TP * q], cmul(a, wq[q]));
        } else {
            constexpr int PB = Q - NPA - 2;
            EB::template phase<PB>(tid, r.v, buf, p.tw);
            if constexpr (PB == NPB - 1) {
                // natural-order result k = k1 + L k2 of the inner FFT: row k1 = rank W + f, k2 = j + TP q
                int f, j;
                EB::out_owner(tid, f, j);
                const uint32_t k1 = rank * W + f;
                if constexpr (!second) {
                    // end of FFT #1: x mult, conjugate (+ Rader's DC bookkeeping at k = 0); the values stay in the registers and are
                    // the input of FFT #2's column pass
                    B2_UNROLL
                    for (int q = 0; q < G::E; ++q) {
                        const uint32_t k = k1 + (uint32_t)(j + G::TP * q) * L;
                        cx<T> w = conj(cmul(r.v[q], ldg(p.mult + k)));
                        if (MODE == 0 && k == 0) {
                            cx<T> x0 = p.in[b * (uint64_t)p.n];
                            if (SW) x0 = swap_ri(x0);
                            const cx<T> dc = x0 + r.v[q];
                            p.out[b * (uint64_t)p.n] = SW ? swap_ri(dc) : dc;
                            w = w + conj(x0);
                        }
                        r.v[q] = w;
                    }
                } else {
                    cx<T>* dst = p.out + b * (uint64_t)p.n;
                    B2_UNROLL
                    for (int q = 0; q < G::E; ++q) {
                        const uint32_t k = k1 + (uint32_t)(j + G::TP * q) * L;
                        if (MODE == 0) {
                            const cx<T> w = conj(r.v[q]);
                            dst[ldg_u32(p.scatter + k)] = SW ? swap_ri(w) : w;
                        } else if (k < p.n) {
                            const cx<T> w = cmul(conj(r.v[q]), ldg(p.chirp + k));
                            st_stream(dst + k, SW ? swap_ri(w) : w);
                        }
                    }
                }
            }
        }
    }
};

#if defined(__CUDACC__)
template <class KT, int P>
struct ClusterPhases {
    static B2_D void run(const typename KT::Params& p, uint32_t bid, int tid, typename KT::Regs& r, cx<typename KT::T>* buf,
                         cx<typename KT::T>* const* remote) {
        KT::template phase<P>(p, bid, tid, r, buf, remote);
        if constexpr (P + 1 < KT::NPHASE) {
            if constexpr (KT::barrier_after(P) == 2)
                cooperative_groups::this_cluster().sync();  // barrier.cluster.arrive.release + wait.acquire: orders the DSMEM stores
            else
                __syncthreads();
            ClusterPhases<KT, P + 1>::run(p, bid, tid, r, buf, remote);
        }
    }
};

template <class KT>
__global__ void __launch_bounds__(KT::NT, KT::MIN_BLOCKS) run_cluster(const __grid_constant__ typename KT::Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = cx<typename KT::T>;
    C* buf = reinterpret_cast<C*>(smem_raw);
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    C* remote[KT::C];
#pragma unroll
    for (int r = 0; r < KT::C; ++r) remote[r] = cl.map_shared_rank(buf, r);
    typename KT::Regs regs;
    ClusterPhases<KT, 0>::run(p, blockIdx.x, (int)threadIdx.x, regs, buf, remote);
}
#endif

}  // namespace b2
