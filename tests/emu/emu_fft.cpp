// emu_fft.cpp -- CPU stepping of the pow2 Stockham kernels (TEST INFRASTRUCTURE).
//
// Runs the very same per-thread phase functions (mpi4py_fft_b200/csrc/fft_core.cuh)
// that fft_pow2_kernel runs on the GPU, one "thread" after another, with the
// barriers where the kernel has __syncthreads().  It lets the index maps of
// every row of fft_configs.h be checked against numpy on a box with no GPU.
// It is never loaded by the product.
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../mpi4py_fft_b200/csrc/fft_core.cuh"
#include "../../mpi4py_fft_b200/csrc/fft_pow2.cuh"
#include "../../mpi4py_fft_b200/csrc/fft_configs.h"

using namespace b2f;

template <class TF, int S>
struct EmuMid {
    using C = typename TF::C;
    static void run(std::vector<C>& regs, C* smem, const C* tw) {
        if constexpr (S < TF::NPASS - 1) {
            constexpr int E = TF::TP ? (int)(sizeof(C) * 0 + 0) : 0;  // unused
            (void)E;
            const int nthr = TF::THREADS;
            const int EE = (int)(regs.size() / nthr);
            for (int tid = 0; tid < nthr; ++tid) {
                C* v = &regs[(size_t)tid * EE];
                TF::template load_shared<S>(v, TF::pencil_of(tid), TF::slot_of(tid), smem);
                TF::template twiddle_dft<S>(v, TF::slot_of(tid), tw);
            }
            // __syncthreads()
            for (int tid = 0; tid < nthr; ++tid) {
                C* v = &regs[(size_t)tid * EE];
                TF::template store_shared<S>(v, TF::pencil_of(tid), TF::slot_of(tid), smem);
            }
            // __syncthreads()
            EmuMid<TF, S + 1>::run(regs, smem, tw);
        }
    }
};

template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS>
static int emu_one(const FftParams& prm_in, long long outer) {
    using TF = TileFFT<T, N, E, RAD, P, STRIDED, PS>;
    using C = cplx<T>;
    FftParams prm = prm_in;
    long long grid;
    if (STRIDED) {
        prm.tiles_per_outer = (prm.inner + P - 1) / P;
        grid = outer * prm.tiles_per_outer;
    } else {
        grid = (prm.npencils + P - 1) / P;
    }
    const C* tw = reinterpret_cast<const C*>(prm.tw);
    const bool swap = prm.swap != 0;
    std::vector<C> smem((size_t)TF::SI::tile_elems);
    std::vector<C> regs((size_t)TF::THREADS * E);
    struct Loc { const C* gin; C* gout; long long in_ns, out_ns; bool valid; };
    std::vector<Loc> loc(TF::THREADS);
    for (long long bid = 0; bid < grid; ++bid) {
        // poison shared memory so that a read of a never-written slot shows up
        for (auto& s : smem) { s.x = (T)1e30; s.y = (T)-1e30; }
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            const int p = TF::pencil_of(tid);
            Loc& L = loc[tid];
            if (STRIDED) {
                const long long o = bid / prm.tiles_per_outer;
                const long long i = (bid - o * prm.tiles_per_outer) * P + p;
                L.valid = i < prm.inner;
                L.gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + i;
                L.gout = reinterpret_cast<C*>(prm.out) + o * prm.out_ostride + i;
                L.in_ns = prm.in_nstride;
                L.out_ns = prm.out_nstride;
            } else {
                const long long gp = bid * P + p;
                L.valid = gp < prm.npencils;
                L.gin = reinterpret_cast<const C*>(prm.in) + gp * prm.in_ostride;
                L.gout = reinterpret_cast<C*>(prm.out) + gp * prm.out_ostride;
                L.in_ns = 1;
                L.out_ns = 1;
            }
        }
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            C* v = &regs[(size_t)tid * E];
            const int p = TF::pencil_of(tid), q = TF::slot_of(tid);
            TF::load_global(v, q, loc[tid].gin, loc[tid].in_ns, loc[tid].valid, swap);
            TF::template twiddle_dft<0>(v, q, tw);
            if (TF::NPASS > 1) TF::template store_shared<0>(v, p, q, smem.data());
        }
        if constexpr (TF::NPASS > 1) {
            // __syncthreads()
            EmuMid<TF, 1>::run(regs, smem.data(), tw);
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                C* v = &regs[(size_t)tid * E];
                const int p = TF::pencil_of(tid), q = TF::slot_of(tid);
                TF::template load_shared<TF::NPASS - 1>(v, p, q, smem.data());
                TF::template twiddle_dft<TF::NPASS - 1>(v, q, tw);
            }
        }
        // all loads of the CTA precede its stores (in-place safety is checked by
        // running in == out from the python side)
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            C* v = &regs[(size_t)tid * E];
            TF::store_global(v, TF::slot_of(tid), loc[tid].gout, loc[tid].out_ns, loc[tid].valid, swap,
                             (T)prm.scale);
        }
    }
    return 0;
}

#define EMU_ROW(N, VAR, E, PC, PSC, PST, PSS, MINB, ...)                                  \
    if (n == N && var == VAR) {                                                           \
        using RAD = Radices<__VA_ARGS__>;                                                 \
        return strided ? emu_one<T, N, E, RAD, PST, true, PSS>(prm, outer)                \
                       : emu_one<T, N, E, RAD, PC, false, PSC>(prm, outer);               \
    }

template <class T>
static int emu_dispatch(int n, int var, bool strided, const FftParams& prm, long long outer) {
    B2F_POW2_TABLE_ALL(EMU_ROW)
    return -1;
}

// (outer, n, inner) C-contiguous block, transform along the middle axis.
// in/out: interleaved complex of the given precision (4 or 8); tw: n entries.
extern "C" int emu_fft_pow2(int precision, int n, int var, long long outer, long long inner,
                            const void* in, void* out, const void* tw, double scale, int swap) {
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.in = in;
    prm.out = out;
    prm.tw = tw;
    prm.scale = scale;
    prm.swap = swap;
    const bool strided = inner > 1;
    if (strided) {
        prm.in_ostride = prm.out_ostride = (long long)n * inner;
        prm.in_nstride = prm.out_nstride = inner;
        prm.inner = inner;
    } else {
        prm.in_ostride = prm.out_ostride = n;
        prm.npencils = outer;
    }
    if (precision == 8) return emu_dispatch<double>(n, var, strided, prm, outer);
    return emu_dispatch<float>(n, var, strided, prm, outer);
}
