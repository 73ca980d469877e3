// fft_tma.cuh -- strided-axis c2c Stockham kernel with TMA-staged tiles (sm_100a).
//
// The register-path kernel (fft_pow2.cuh) keeps a tile's points in registers
// while its loads are in flight, so a CTA cannot compute and load at once and
// big tiles (N x 128 B = 64..128 KiB) leave room for one or two CTAs per SM:
// the strided axes ran latency-bound at 55-75 % of the HBM roofline.  Here a
// persistent CTA owns a ring of shared-memory stages that the TMA engine fills
// (cp.async.bulk.tensor, one box of P contiguous elements x <=256 rows per
// request, completion on an mbarrier) while the threads transform the previous
// tile:
//
//   TMA  : HBM --box--> stage[s]                      (async, no registers, no LSU)
//   pass0: stage[s] -> registers -> R0-point DFTs     (stage s is free again:
//          the next tile's box loads are issued right here)
//   mid  : registers <-> exchange buffer (padded), twiddle, DFT
//   last : registers -> HBM, normalisation fused      (coalesced 128 B runs)
//
// With SPLIT the exchange buffer holds one real component at a time (re, then
// im), halving its footprint so that a 128 KiB tile (N = 1024, complex128, 128 B
// rows) still fits next to its stage.  Backward = forward with re/im swapped on
// the way in and out, as in fft_pow2.cuh.
//
// Replaces fftw_execute_dft on plans with is/os > 1
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:52-56, fftw_xfftn.pyx:29-30).
#pragma once
#include <type_traits>
#if defined(__CUDACC__)
#include <cuda.h>
#endif
#include "fft_core.cuh"

namespace b2f {

struct TmaParams {
    const void* in;          // used by the cp.async loader only (TMA reads through its descriptor)
    long long in_ostride, in_nstride;
    void* out;
    const void* tw;
    long long out_ostride;   // elements between consecutive outer indices
    long long out_nstride;   // elements between consecutive points of a pencil
    long long inner;         // extent of the contiguous inner index
    long long tiles_per_outer;
    long long ntiles;
    double scale;
    int swap;
    PeerStore peer;          // peer.p > 0: fused redistribution (fft_core.cuh)
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// LOADER 1: the stage is filled with cp.async (LDGSTS) by all threads instead of
// TMA boxes -- same asynchrony, goes through the LSU address path.  Serves the
// layouts a TMA descriptor cannot express (8-byte row pitch) and rows that sit
// in different 2 MiB pages, where the LSU path sustains more translations per
// second than the TMA engine (tools/probe/stride_probe.cu).
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* dst, const void* src, bool valid) {
    const int sz = valid ? BYTES : 0;   // src-size 0 -> zero fill
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}

#endif  // __CUDACC__

// exchange-buffer element: the complex value, or one real component (SPLIT)
template <class TF, bool SPLIT>
struct Exchange {
    using C = typename TF::C;
    using T = typename TF::Real;
    using SI = typename TF::SI;
    using X = typename std::conditional<SPLIT, T, C>::type;
    static constexpr size_t bytes = sizeof(X) * (size_t)SI::tile_elems;

    // registers (after pass S) -> buffer, component c (ignored unless SPLIT)
    template <int S>
    static B2F_HD void put(const C* v, int p, int q, X* buf, int c) {
        constexpr int R = TF::RADS::get(S);
        constexpr int Ns = TF::RADS::before(S);
        constexpr int NB = TF::EPT / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = q + b * TF::TP;
            const int base = (j / Ns) * (Ns * R) + (j % Ns);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if constexpr (SPLIT) buf[SI::at(p, base + r * Ns)] = c ? v[b * R + r].y : v[b * R + r].x;
                else buf[SI::at(p, base + r * Ns)] = v[b * R + r];
            }
        }
    }
    // buffer -> registers (before pass S)
    template <int S>
    static B2F_HD void get(C* v, int p, int q, const X* buf, int c) {
        constexpr int R = TF::RADS::get(S);
        constexpr int NB = TF::EPT / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = q + b * TF::TP + r * (TF::LEN / R);
                if constexpr (SPLIT) {
                    if (c) v[b * R + r].y = buf[SI::at(p, i)];
                    else v[b * R + r].x = buf[SI::at(p, i)];
                } else {
                    v[b * R + r] = buf[SI::at(p, i)];
                }
            }
        }
    }
};

// pass-0 read of a dense [N][P] stage (what a TMA box load leaves behind)
template <class TF>
static B2F_HD void load_stage(typename TF::C* v, int p, int q, const typename TF::C* stage, bool swap) {
    using C = typename TF::C;
    constexpr int R = TF::RADS::get(0);
    constexpr int NB = TF::EPT / R;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = q + b * TF::TP + r * (TF::LEN / R);
            C a = stage[n * TF::PEN + p];
            if (swap) { auto t = a.x; a.x = a.y; a.y = t; }
            v[b * R + r] = a;
        }
    }
}

// Output staging of the TMA-store flavour (OPT bit 3): the last pass writes its
// results into the (then idle) exchange buffer as dense rows [n][P] and the TMA
// engine stores them (cp.async.bulk.tensor, boxes of P elements x <= 256 rows)
// while the CTA already works on the next tile.  A buffer smaller than the tile
// takes the rows in COUNT chunks of ROWS consecutive rows.
template <class TF, class EX>
struct OutChunks {
    static constexpr int fit = (int)(EX::bytes / (sizeof(typename TF::C) * TF::PEN));
    static constexpr int count = fit >= TF::LEN ? 1 : fit >= TF::LEN / 2 ? 2 : fit >= TF::LEN / 4 ? 4 : 8;
    static constexpr int rows = TF::LEN / count;
    static constexpr int box_rows = rows > 256 ? 256 : rows;
    static_assert(TF::RADS::get(TF::NPASS - 1) % count == 0, "last radix must be a multiple of the chunk count");
    static_assert(fit >= TF::LEN / 8, "exchange buffer too small to stage the output");
};

// results of the last pass that fall into chunk c -> staging rows (scaled, re/im swapped back)
template <class TF, int COUNT>
static B2F_HD void stage_out_rows(const typename TF::C* v, int p, int q, typename TF::C* xo, int c, bool swap,
                                  typename TF::Real scale) {
    using C = typename TF::C;
    constexpr int R = TF::RADS::get(TF::NPASS - 1);
    constexpr int NB = TF::EPT / R;
    constexpr int ROWS = TF::LEN / COUNT;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if ((r * COUNT) / R == c) {
                const int n = q + b * TF::TP + r * (TF::LEN / R);
                C a = v[b * R + r];
                a.x *= scale;
                a.y *= scale;
                if (swap) { auto t = a.x; a.x = a.y; a.y = t; }
                xo[(n - c * ROWS) * TF::PEN + p] = a;
            }
        }
    }
}

#if defined(__CUDACC__)

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every committed bulk store has READ its shared-memory source (the buffer may be rewritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#endif  // __CUDACC__

#if defined(__CUDACC__)

template <class TF, class EX, int S>
struct TmaMid {
    using C = typename TF::C;
    // hook(k): called at two points of the first exchange (k = 2, 3) so that the
    // loader can spread its requests over the tile's lifetime
    template <class HOOK>
    static __device__ __forceinline__ void run(C* v, int p, int q, typename EX::X* xbuf, const C* __restrict__ tw,
                                               bool split, HOOK&& hook) {
        if constexpr (S < TF::NPASS) {
            // exchange between pass S-1 and pass S (the caller's values are post pass S-1)
            EX::template put<S - 1>(v, p, q, xbuf, 0);
            __syncthreads();
            EX::template get<S>(v, p, q, xbuf, 0);
            if constexpr (S == 1) hook(2);
            if (split) {
                __syncthreads();
                EX::template put<S - 1>(v, p, q, xbuf, 1);
                __syncthreads();
                EX::template get<S>(v, p, q, xbuf, 1);
            }
            if constexpr (S == 1) hook(3);
            TF::template twiddle_dft<S>(v, q, tw);
            if constexpr (S + 1 < TF::NPASS) __syncthreads();   // buffer is rewritten by the next exchange
            TmaMid<TF, EX, S + 1>::run(v, p, q, xbuf, tw, split, hook);
        }
    }
};

// OPT bit 0: pass twiddles are copied to shared memory once per CTA (the L1 left
// beside a ~200 KiB carve-out does not keep them); bit 1: the cp.async loader
// spreads the next tile's requests over the current tile's phases instead of
// issuing them in one burst (LSU queue pressure)
// bit 3: results leave through the exchange buffer and TMA tensor stores (UTMASTG) instead of per-thread
// STG: the store burst of a tile no longer stalls every warp on the LSU queue, it drains while the next
// tile is computed -- and goes through the TMA unit's address path while the loads use the LSU's
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int LOADER, bool SWAP, bool PEER, int OPT>
__device__ __forceinline__ void fft_tma_body(const CUtensorMap* map_in, const TmaParams& prm,
                                             const CUtensorMap* map_out = nullptr) {
    constexpr bool TSTORE = (OPT & 8) != 0 && !PEER;
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    using C = cplx<T>;
    constexpr uint32_t TILE_BYTES = (uint32_t)(sizeof(C) * N * P);
    constexpr int BR = N > 256 ? 256 : N;          // rows per TMA box (box dims are capped at 256)
    extern __shared__ __align__(1024) unsigned char b2f_tma_smem[];
    __shared__ uint64_t full[STAGES];
    C* stages = reinterpret_cast<C*>(b2f_tma_smem);
    typename EX::X* xbuf = reinterpret_cast<typename EX::X*>(b2f_tma_smem + (size_t)STAGES * TILE_BYTES);

    const int tid = threadIdx.x;
    const int p = TF::pencil_of(tid);
    const int q = TF::slot_of(tid);
    const C* __restrict__ tw = reinterpret_cast<const C*>(prm.tw);
    const long long first = blockIdx.x, step = gridDim.x;
    if constexpr ((OPT & 1) != 0) {
        C* tws = reinterpret_cast<C*>(b2f_tma_smem + (size_t)STAGES * TILE_BYTES + EX::bytes);
        for (int k = tid; k < RAD::tw_total(); k += TF::THREADS) tws[k] = tw[k];
        tw = tws;
        __syncthreads();
    }

    if (LOADER == 0) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    auto issue = [&](long long t, int s) {
        const long long o = t / prm.tiles_per_outer;
        const long long i0 = (t - o * prm.tiles_per_outer) * P;
        mbar_expect_tx(&full[s], TILE_BYTES);
        unsigned char* dst = b2f_tma_smem + (size_t)s * TILE_BYTES;
#pragma unroll
        for (int r = 0; r < N; r += BR)
            tma_load_3d(dst + (size_t)r * P * sizeof(C), map_in, (int)(2 * i0), r, (int)o, &full[s]);
    };
    // cp.async flavour: every thread copies the E elements it will read back
    // (row q + e*TP, column p); one commit group per tile slot, empty or not
    // part k of 4 (or everything when part < 0); the commit closes the tile's group
    auto issue_cpa_part = [&](long long t, int s, int part) {
        if (t < prm.ntiles) {
            const long long o = t / prm.tiles_per_outer;
            const long long i = (t - o * prm.tiles_per_outer) * P + p;
            const bool ok = i < prm.inner;
            const C* src = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + (ok ? i : 0);
            C* dst = stages + (size_t)s * N * P + p;
            constexpr int Q4 = E >= 4 ? E / 4 : E;
            const int e0 = part < 0 ? 0 : part * Q4, e1 = part < 0 ? E : (part == 3 || E < 4 ? E : e0 + Q4);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (e >= e0 && e < e1) {
                    const int row = q + e * TF::TP;
                    cp_async_elem<(int)sizeof(C)>(dst + row * P, src + (long long)row * prm.in_nstride, ok);
                }
            }
        }
        if (part < 0 || part == 3) asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto issue_cpa = [&](long long t, int s) { issue_cpa_part(t, s, -1); };
    // OPT bit 2: pull the tile after next into L2 ahead of time.  A prefetch needs no
    // landing buffer, so it takes the address-translation latency of rows that sit in
    // different pages off the critical path and doubles the bytes in flight per SM.
    auto prefetch_tile = [&](long long t) {
        if (t < prm.ntiles) {
            const long long o = t / prm.tiles_per_outer;
            const long long i0 = (t - o * prm.tiles_per_outer) * P;
            const char* base = reinterpret_cast<const char*>(reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + i0);
            constexpr int SEG = (int)(P * sizeof(C) + 127) / 128;     // 128-byte lines per row
            for (int r = tid; r < N * SEG; r += TF::THREADS) {
                const int row = r / SEG, seg = r - row * SEG;
                const char* ptr = base + (long long)row * prm.in_nstride * (long long)sizeof(C) + seg * 128;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
        }
    };
    constexpr bool SPREAD = LOADER == 1 && (OPT & 2) != 0 && E >= 4 && TF::NPASS > 1;
    if (LOADER == 0) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (first + s * step < prm.ntiles) issue(first + s * step, s);
        }
    } else {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) issue_cpa(first + s * step, s);
    }

    if constexpr ((OPT & 4) != 0) prefetch_tile(first + (long long)STAGES * step);
    int s = 0;
    uint32_t parity = 0;
    for (long long t = first; t < prm.ntiles; t += step) {
        const long long o = t / prm.tiles_per_outer;
        const long long i = (t - o * prm.tiles_per_outer) * P + p;
        const bool valid = i < prm.inner;
        C* gout = reinterpret_cast<C*>(prm.out) + o * prm.out_ostride + i;

        C v[E];
        if (LOADER == 0) {
            mbar_wait(&full[s], parity);
        } else {
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
            __syncthreads();   // everybody's copies of this tile have landed
        }
        load_stage<TF>(v, p, q, stages + (size_t)s * N * P, SWAP);
        if (TSTORE && tid == 0) bulk_wait_read();   // the previous tile's TMA stores have read the exchange buffer
        // TMA loader: one thread refills the stage for everybody, so everybody must have read it.  cp.async
        // loader: a thread refills exactly the elements it has just read itself, and the barrier after
        // wait_group above already separates the previous tile's exchange reads from this tile's writes.
        if (LOADER == 0 || TSTORE) __syncthreads();
        const long long tn = t + (long long)STAGES * step;
        if (LOADER == 0) {
            if (tid == 0 && tn < prm.ntiles) issue(tn, s);
        } else if (SPREAD) {
            issue_cpa_part(tn, s, 0);
        } else {
            issue_cpa(tn, s);
        }
        if constexpr ((OPT & 4) != 0) prefetch_tile(tn + step);
        TF::template twiddle_dft<0>(v, q, tw);
        if (SPREAD) issue_cpa_part(tn, s, 1);
        TmaMid<TF, EX, 1>::run(v, p, q, xbuf, tw, SPLIT, [&](int k) {
            if (SPREAD) issue_cpa_part(tn, s, k);
        });
        if constexpr (PEER) {
            long long part = 0, rest = 0;
            if (valid) prm.peer.locate(o, i, &part, &rest);
            TF::store_peer(v, q, prm.peer, part, rest, valid, SWAP, (T)prm.scale);
        } else if constexpr (TSTORE) {
            using OC = OutChunks<TF, EX>;
            C* xo = reinterpret_cast<C*>(xbuf);
            const int i0 = (int)((t - o * prm.tiles_per_outer) * P);
#pragma unroll
            for (int c = 0; c < OC::count; ++c) {
                if (c > 0 && tid == 0) bulk_wait_read();
                __syncthreads();   // exchange reads (c = 0) / the TMA reads of the previous chunk are done
                stage_out_rows<TF, OC::count>(v, p, q, xo, c, SWAP, (T)prm.scale);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
#pragma unroll
                    for (int k = 0; k < OC::rows; k += OC::box_rows)
                        tma_store_4d(map_out, xo + (size_t)k * P, 2 * i0, c * OC::rows + k, (int)o, 0);
                    bulk_commit();
                }
            }
        } else {
            TF::store_global(v, q, gout, prm.out_nstride, valid, SWAP, (T)prm.scale);
        }
        if (++s == STAGES) {
            s = 0;
            parity ^= 1;
        }
    }
    if (TSTORE && tid == 0) bulk_wait_all();   // shared memory stays alive until the last store has left
}

template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MO>
__global__ void __launch_bounds__((N / E) * P, (MO & 15))
fft_tma_kernel(const __grid_constant__ CUtensorMap map_in, const TmaParams prm) {
    if (prm.swap) fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 0, true, false, (MO >> 4)>(&map_in, prm);
    else fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 0, false, false, (MO >> 4)>(&map_in, prm);
}
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MO>
__global__ void __launch_bounds__((N / E) * P, (MO & 15))
fft_tma_peer_kernel(const __grid_constant__ CUtensorMap map_in, const TmaParams prm) {
    if (prm.swap) fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 0, true, true, (MO >> 4)>(&map_in, prm);
    else fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 0, false, true, (MO >> 4)>(&map_in, prm);
}

// cp.async loader + TMA tensor stores (OPT bit 3 of MO): the descriptor describes the OUTPUT
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MO>
__global__ void __launch_bounds__((N / E) * P, (MO & 15))
fft_cpa_ts_kernel(const __grid_constant__ CUtensorMap map_out, const TmaParams prm) {
    if (prm.swap) fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 1, true, false, (MO >> 4)>(nullptr, prm, &map_out);
    else fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 1, false, false, (MO >> 4)>(nullptr, prm, &map_out);
}

// the same pipeline with the cp.async loader (no descriptor)
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MO>
__global__ void __launch_bounds__((N / E) * P, (MO & 15))
fft_cpa_kernel(const TmaParams prm) {
    if (prm.swap) fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 1, true, false, (MO >> 4)>(nullptr, prm);
    else fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 1, false, false, (MO >> 4)>(nullptr, prm);
}
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MO>
__global__ void __launch_bounds__((N / E) * P, (MO & 15))
fft_cpa_peer_kernel(const TmaParams prm) {
    if (prm.swap) fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 1, true, true, (MO >> 4)>(nullptr, prm);
    else fft_tma_body<T, N, E, RAD, P, PS, STAGES, SPLIT, 1, false, true, (MO >> 4)>(nullptr, prm);
}

#endif  // __CUDACC__

}  // namespace b2f
