// fft_rot.cuh -- "rotating" c2c Stockham kernel (sm_100a): transform the
// CONTIGUOUS axis of a block and store the result with the axes rotated,
//
//     in  [A][B][C]  (C contiguous, transformed)   ->   out [B][C][A]
//
// so that a multi-axis stage never reads or writes short rows that lie in
// different 2 MiB pages.  Three such steps bring a 3-D block back to its natural
// layout with every axis transformed:
//     [x][y][z] -z-> [y][z][x] -x-> [z][x][y] -y-> [x][y][z].
// The plain strided kernels (fft_tma.cuh) read AND write N rows of P*16 bytes
// per tile; along the slowest axis of a 1024^3 complex128 block those rows are
// 16 MiB apart and the step runs at the address-translation rate of 128-byte
// requests (0.69 of the HBM roofline, profiles/r1e_ncu_full_1024.txt).  Here a
// tile is P whole pencils:
//
//   load : P bulk copies of N*16 bytes each (cp.async.bulk global -> shared,
//          the TMA engine's 1-D form, completion on an mbarrier) -- the largest
//          requests the memory system takes, no registers, no LSU slots;
//   math : the same register/shared-memory Stockham passes as every other kernel
//          of the library (TileFFT, Exchange);
//   store: rows of P contiguous elements at pitch A -- rows of one tile are
//          A*16 bytes apart (KiBs, the same pages), the friendly strided pattern.
//
// One read and one write of the block per axis, as before; what changes is that
// every request is either large or page-local.  Replaces fftw_execute_dft of a
// multi-axis plan (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:52-56).
//
// Status: OPT-IN (option "rotate" / B2F_ROTATE=1).  Measured on 1024^3 complex128 each step runs at
// 0.86 of the HBM roofline whatever the layout, which ties with the per-axis schedule (1.03 + 0.92 +
// 0.69) in short runs and loses to it under the power cap (profiles/r2_rot_sweeps.txt, DESIGN.md
// section 3).  The kernel also serves the transposing transform of the four-step split (lengths.h).
#pragma once
#include "fft_tma.cuh"

namespace b2f {

struct RotParams {
    const void* in;
    void* out;
    const void* tw;
    long long in_istride;    // elements between pencils i and i+1 (the axis that becomes contiguous)
    long long in_ostride;    // elements between pencils o and o+1
    long long in_bstride;    // elements between batches
    long long out_ostride;   // out[b][o][k][i] = out + b*out_bstride + o*out_ostride + k*out_nstride + i
    long long out_nstride;
    long long out_bstride;
    long long I, O;          // extents of i and o
    long long tiles_per_o;   // ceil(I / P)
    long long tiles_per_b;   // O * tiles_per_o
    long long ntiles;        // batches * tiles_per_b
    double scale;
    int swap;
};

// padding of a pencil row in the stage (row pitch N + PAD elements, a multiple of 16 bytes): a
// shared-memory wavefront serves 128 bytes, i.e. G = 128 / sizeof(element) lanes, p = lane % P fastest
// and q consecutive; their addresses p * PITCH + q are conflict free when PITCH = G / P (mod G), P <= G.
template <class T, int P> struct RotPad {
    static constexpr int G = (int)(128 / (2 * sizeof(T)));
    static constexpr int unit = (int)(16 / (2 * sizeof(T)));       // elements per 16 bytes
    static constexpr int want = P >= G ? 1 : G / P;
    static constexpr int value = want < unit ? unit : want;
};

// pass-0 read of a [P][N + PAD] stage (what the bulk copies leave behind)
template <class TF, int PITCH>
static B2F_HD void load_rows(typename TF::C* v, int p, int q, const typename TF::C* stage, bool swap) {
    using C = typename TF::C;
    constexpr int R = TF::RADS::get(0);
    constexpr int NB = TF::EPT / R;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = q + b * TF::TP + r * (TF::LEN / R);
            C a = stage[p * PITCH + n];
            if (swap) { auto t = a.x; a.x = a.y; a.y = t; }
            v[b * R + r] = a;
        }
    }
}

#if defined(__CUDACC__)

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the exchanges and passes after pass 0, as TmaMid (fft_tma.cuh) but with the barrier
// handed in: a CTA may hold several independent warp groups (named barriers)
template <class TF, class EX, int S>
struct RotMid {
    using C = typename TF::C;
    template <class SYNC>
    static __device__ __forceinline__ void run(C* v, int p, int q, typename EX::X* xbuf, const C* __restrict__ tw,
                                               bool split, SYNC&& sync) {
        if constexpr (S < TF::NPASS) {
            EX::template put<S - 1>(v, p, q, xbuf, 0);
            sync();
            EX::template get<S>(v, p, q, xbuf, 0);
            if (split) {
                sync();
                EX::template put<S - 1>(v, p, q, xbuf, 1);
                sync();
                EX::template get<S>(v, p, q, xbuf, 1);
            }
            TF::template twiddle_dft<S>(v, q, tw);
            if constexpr (S + 1 < TF::NPASS) sync();   // buffer is rewritten by the next exchange
            RotMid<TF, EX, S + 1>::run(v, p, q, xbuf, tw, split, sync);
        }
    }
};

// OPT bit 0: pass twiddles in shared memory (as fft_tma.cuh); bit 3: results leave through
// shared memory and TMA stores instead of per-thread global stores; bits 4-5: log2 of the
// number of independent warp GROUPS in the CTA.  A tile of 8 pencils of 1024 complex128
// points is 128 KiB -- half the register file, more than half the shared memory -- so only
// one CTA fits an SM and all its warps move through the load / fp64 / exchange / store phases
// in lockstep: the fp64 pipe idles while shared memory or the store queue is busy and vice
// versa.  With GROUPS > 1 the CTA runs GROUPS narrower tiles at once, each group with its own
// stage, exchange buffer, mbarriers and named barrier (the twiddle table is shared), so that
// one group's fp64 phase overlaps another's memory phases.
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, bool SWAP, int OPT>
__device__ __forceinline__ void fft_rot_body(const CUtensorMap* map_out, const RotParams& prm) {
    constexpr bool TSTORE = (OPT & 8) != 0;
    constexpr int GROUPS = 1 << ((OPT >> 4) & 3);
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    using C = cplx<T>;
    constexpr int GT = TF::THREADS;
    constexpr int PITCH = N + RotPad<T, P>::value;
    constexpr uint32_t ROW_BYTES = (uint32_t)(sizeof(C) * N);
    constexpr size_t STAGE_BYTES = sizeof(C) * (size_t)PITCH * P;
    constexpr size_t GROUP_BYTES = (size_t)STAGES * STAGE_BYTES + EX::bytes;
    static_assert(!(TSTORE && GROUPS > 1), "bulk-group waits are per thread: one group only");
    extern __shared__ __align__(1024) unsigned char b2f_rot_smem[];
    __shared__ uint64_t full_all[GROUPS * STAGES];

    const int g = GROUPS > 1 ? (int)threadIdx.x / GT : 0;
    const int tid = GROUPS > 1 ? (int)threadIdx.x - g * GT : (int)threadIdx.x;
    unsigned char* gsm = b2f_rot_smem + (size_t)g * GROUP_BYTES;
    uint64_t* full = full_all + g * STAGES;
    typename EX::X* xbuf = reinterpret_cast<typename EX::X*>(gsm + (size_t)STAGES * STAGE_BYTES);
    auto sync = [&]() {
        if constexpr (GROUPS > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(GT) : "memory");
        else __syncthreads();
    };

    const int p = TF::pencil_of(tid);
    const int q = TF::slot_of(tid);
    const C* __restrict__ tw = reinterpret_cast<const C*>(prm.tw);
    const long long first = (long long)blockIdx.x * GROUPS + g, step = (long long)gridDim.x * GROUPS;
    if constexpr ((OPT & 1) != 0) {
        C* tws = reinterpret_cast<C*>(b2f_rot_smem + (size_t)GROUPS * GROUP_BYTES);
        for (int k = threadIdx.x; k < RAD::tw_total(); k += GROUPS * GT) tws[k] = tw[k];
        tw = tws;
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tile t = (batch, o, i-tile): i-tiles of one o are consecutive, so the CTAs (and groups)
    // that run side by side write neighbouring runs of the same output rows
    auto issue = [&](long long t, int s) {
        const long long b = t / prm.tiles_per_b;
        const long long rem = t - b * prm.tiles_per_b;
        const long long o = rem / prm.tiles_per_o;
        const long long i0 = (rem - o * prm.tiles_per_o) * P;
        const long long left = prm.I - i0;
        const int np = left < P ? (int)left : P;
        const C* src = reinterpret_cast<const C*>(prm.in) + b * prm.in_bstride + o * prm.in_ostride + i0 * prm.in_istride;
        unsigned char* dst = gsm + (size_t)s * STAGE_BYTES;
        // generic-proxy reads of this stage are ordered before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[s], ROW_BYTES * (uint32_t)np);
        for (int j = 0; j < np; ++j)
            bulk_load(dst + (size_t)j * PITCH * sizeof(C), src + (long long)j * prm.in_istride, ROW_BYTES, &full[s]);
    };
    // OPT bit 2: the stage is filled with cp.async (LDGSTS, 16 bytes per request, consecutive
    // threads -> consecutive addresses of a pencil) by all threads of the group instead of bulk
    // copies; one commit group per tile slot, empty or not
    constexpr bool CPA = (OPT & 4) != 0;
    auto issue_cpa = [&](long long t, int s) {
        if (t < prm.ntiles) {
            const long long b = t / prm.tiles_per_b;
            const long long rem = t - b * prm.tiles_per_b;
            const long long o = rem / prm.tiles_per_o;
            const long long i0 = (rem - o * prm.tiles_per_o) * P;
            const long long left = prm.I - i0;
            const int np = left < P ? (int)left : P;
            const char* src = reinterpret_cast<const char*>(reinterpret_cast<const C*>(prm.in) + b * prm.in_bstride +
                                                            o * prm.in_ostride + i0 * prm.in_istride);
            unsigned char* dst = gsm + (size_t)s * STAGE_BYTES;
            constexpr int CHUNKS = (int)(ROW_BYTES / 16);
            for (int j = 0; j < np; ++j) {
                const char* sp = src + (long long)j * prm.in_istride * (long long)sizeof(C);
                unsigned char* dp = dst + (size_t)j * PITCH * sizeof(C);
#pragma unroll 4
                for (int c = tid; c < CHUNKS; c += GT) cp_async_elem<16>(dp + c * 16, sp + c * 16, true);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if constexpr (CPA) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) issue_cpa(first + s * step, s);
    } else if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (first + s * step < prm.ntiles) issue(first + s * step, s);
    }

    int s = 0;
    uint32_t parity = 0;
    for (long long t = first; t < prm.ntiles; t += step) {
        const long long b = t / prm.tiles_per_b;
        const long long rem = t - b * prm.tiles_per_b;
        const long long o = rem / prm.tiles_per_o;
        const long long i = (rem - o * prm.tiles_per_o) * P + p;
        const bool valid = i < prm.I;
        C* gout = reinterpret_cast<C*>(prm.out) + b * prm.out_bstride + o * prm.out_ostride + i;

        C v[E];
        if constexpr (CPA) {
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
            sync();   // everybody's copies of this tile have landed
        } else {
            mbar_wait(&full[s], parity);
        }
        load_rows<TF, PITCH>(v, p, q, reinterpret_cast<const C*>(gsm + (size_t)s * STAGE_BYTES), SWAP);
        if (TSTORE && tid == 0) bulk_wait_read();   // the previous tile's stores have read the exchange buffer
        sync();   // every thread has read stage s (and is done with the exchange buffer of the previous tile)
        const long long tn = t + (long long)STAGES * step;
        if constexpr (CPA) issue_cpa(tn, s);
        else if (tid == 0 && tn < prm.ntiles) issue(tn, s);
        TF::template twiddle_dft<0>(v, q, tw);
        RotMid<TF, EX, 1>::run(v, p, q, xbuf, tw, SPLIT, sync);
        if constexpr (TSTORE) {
            using OC = OutChunks<TF, EX>;
            C* xo = reinterpret_cast<C*>(xbuf);
            const int i0 = (int)((rem - o * prm.tiles_per_o) * P);
#pragma unroll
            for (int c = 0; c < OC::count; ++c) {
                if (c > 0 && tid == 0) bulk_wait_read();
                sync();   // exchange reads (c = 0) / the TMA reads of the previous chunk are done
                stage_out_rows<TF, OC::count>(v, p, q, xo, c, SWAP, (T)prm.scale);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                sync();
                if (tid == 0) {
#pragma unroll
                    for (int k = 0; k < OC::rows; k += OC::box_rows)
                        tma_store_4d(map_out, xo + (size_t)k * P, 2 * i0, c * OC::rows + k, (int)o, (int)b);
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
__global__ void __launch_bounds__((N / E) * P * (1 << ((MO >> 8) & 3)), (MO & 15))
fft_rot_kernel(const __grid_constant__ CUtensorMap map_out, const RotParams prm) {
    if (prm.swap) fft_rot_body<T, N, E, RAD, P, PS, STAGES, SPLIT, true, (MO >> 4)>(&map_out, prm);
    else fft_rot_body<T, N, E, RAD, P, PS, STAGES, SPLIT, false, (MO >> 4)>(&map_out, prm);
}

#endif  // __CUDACC__

}  // namespace b2f
