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
#include "../../mpi4py_fft_b200/csrc/fft_tma.cuh"
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
    std::vector<C> twv((size_t)RAD::tw_total());
    build_pass_twiddles<T, RAD>(twv.data());   // same host builder as the library (fft_core.cuh)
    const C* tw = twv.data();
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
                L.gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + (prm.in_istride ? i * prm.in_istride : i);
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
            if (prm.trunc.n > 0 && swap) TF::load_global_padded(v, q, loc[tid].gin, loc[tid].in_ns, loc[tid].valid, swap, prm.trunc);
            else TF::load_global(v, q, loc[tid].gin, loc[tid].in_ns, loc[tid].valid, swap);
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
        std::vector<C> nyq((size_t)P);
        if (prm.trunc.n > 0 && !swap) {
            // dealiasing flavour, forward: publish the Nyquist partner, barrier, truncated store
            for (auto& z : nyq) { z.x = (T)1e30; z.y = (T)-1e30; }
            for (int tid = 0; tid < TF::THREADS; ++tid)
                TF::store_truncated_publish(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), nyq.data(), prm.trunc,
                                            (T)prm.scale);
            // __syncthreads()
            for (int tid = 0; tid < TF::THREADS; ++tid)
                TF::store_truncated(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), loc[tid].gout, loc[tid].out_ns,
                                    loc[tid].valid, swap, (T)prm.scale, nyq.data(), prm.trunc);
            continue;
        }
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            C* v = &regs[(size_t)tid * E];
            if (prm.peer.p > 0) {
                // the kernel's PEER flavour: pencil coordinates -> (part, rest), then store_peer
                long long po, pi, part = 0, rest = 0;
                if (STRIDED) {
                    po = bid / prm.tiles_per_outer;
                    pi = (bid - po * prm.tiles_per_outer) * P + TF::pencil_of(tid);
                } else {
                    po = bid * P + TF::pencil_of(tid);
                    pi = 0;
                }
                if (loc[tid].valid) prm.peer.locate(po, pi, &part, &rest);
                TF::store_peer(v, TF::slot_of(tid), prm.peer, part, rest, loc[tid].valid, swap, (T)prm.scale);
            } else {
                TF::store_global(v, TF::slot_of(tid), loc[tid].gout, loc[tid].out_ns, loc[tid].valid, swap,
                                 (T)prm.scale);
            }
        }
    }
    return 0;
}

#define EMU_CONTIG(N, VAR, E, P, PS, MINB, ...) \
    if (n == N && var == VAR) return emu_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS>(prm, outer);
// float strided tiles hold twice the pencils (same bytes per row), as in fft_pow2_inst.cuh
#define EMU_STRIDED(N, VAR, E, P, PS, MINB, ...) \
    if (n == N && var == VAR)                    \
        return emu_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS>(prm, outer);

template <class T>
static int emu_dispatch(int n, int var, bool strided, const FftParams& prm, long long outer) {
    if (strided) {
        B2F_STRIDED_ALL(EMU_STRIDED)
    } else {
        B2F_CONTIG_ALL(EMU_CONTIG)
    }
    return -1;
}

// ---- TMA-staged strided kernel (fft_tma.cuh): the box load is emulated by a
// dense copy with zero fill outside the array, everything after it is the
// kernel's own per-thread code with its barriers.
template <class TF, class EX, int S>
struct EmuTmaMid {
    using C = typename TF::C;
    static void run(std::vector<C>& regs, typename EX::X* xbuf, const C* tw, bool split) {
        if constexpr (S < TF::NPASS) {
            const int nthr = TF::THREADS, E = TF::EPT;
            for (int c = 0; c < (split ? 2 : 1); ++c) {
                for (int tid = 0; tid < nthr; ++tid)
                    EX::template put<S - 1>(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), xbuf, c);
                // __syncthreads()
                for (int tid = 0; tid < nthr; ++tid)
                    EX::template get<S>(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), xbuf, c);
                // __syncthreads()
            }
            for (int tid = 0; tid < nthr; ++tid)
                TF::template twiddle_dft<S>(&regs[(size_t)tid * E], TF::slot_of(tid), tw);
            EmuTmaMid<TF, EX, S + 1>::run(regs, xbuf, tw, split);
        }
    }
};

template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT>
static int emu_tma_one(const FftParams& prm, long long outer) {
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    using C = cplx<T>;
    const long long tpo = (prm.inner + P - 1) / P, ntiles = outer * tpo;
    std::vector<C> twv((size_t)RAD::tw_total());
    build_pass_twiddles<T, RAD>(twv.data());
    std::vector<C> stage((size_t)N * P);
    std::vector<typename EX::X> xbuf((size_t)TF::SI::tile_elems);
    std::vector<C> regs((size_t)TF::THREADS * E);
    const C* gin = reinterpret_cast<const C*>(prm.in);
    C* gout = reinterpret_cast<C*>(prm.out);
    const bool swap = prm.swap != 0;
    for (long long t = 0; t < ntiles; ++t) {
        const long long o = t / tpo, i0 = (t - o * tpo) * P;
        for (int n = 0; n < N; ++n)
            for (int pp = 0; pp < P; ++pp) {
                C z = {(T)0, (T)0};
                stage[(size_t)n * P + pp] = (i0 + pp < prm.inner) ? gin[o * prm.in_ostride + (long long)n * prm.in_nstride + i0 + pp] : z;
            }
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            C* v = &regs[(size_t)tid * E];
            load_stage<TF>(v, TF::pencil_of(tid), TF::slot_of(tid), stage.data(), swap);
        }
        // __syncthreads(); next box load is issued here
        for (auto& x : xbuf) x = typename EX::X{};
        for (int tid = 0; tid < TF::THREADS; ++tid)
            TF::template twiddle_dft<0>(&regs[(size_t)tid * E], TF::slot_of(tid), twv.data());
        EmuTmaMid<TF, EX, 1>::run(regs, xbuf.data(), twv.data(), SPLIT);
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            const long long i = i0 + TF::pencil_of(tid);
            if (prm.peer.p > 0) {
                long long part = 0, rest = 0;
                if (i < prm.inner) prm.peer.locate(o, i, &part, &rest);
                TF::store_peer(&regs[(size_t)tid * E], TF::slot_of(tid), prm.peer, part, rest, i < prm.inner, swap,
                               (T)prm.scale);
            } else {
                TF::store_global(&regs[(size_t)tid * E], TF::slot_of(tid), gout + o * prm.out_ostride + i,
                                 prm.out_nstride, i < prm.inner, swap, (T)prm.scale);
            }
        }
    }
    return 0;
}

#define EMU_TMA(N, VAR, E, P, PS, STAGES, SPLIT, MINB, ...) \
    if (n == N && var == VAR)                                \
        return emu_tma_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), PS, STAGES, SPLIT != 0>(prm, outer);

template <class T>
static int emu_tma_dispatch(int n, int var, const FftParams& prm, long long outer) {
    B2F_TMA_TABLE(EMU_TMA)
    return -1;
}

extern "C" int emu_fft_tma(int precision, int n, int var, long long outer, long long inner, const void* in, void* out,
                           double scale, int swap) {
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.in = in;
    prm.out = out;
    prm.scale = scale;
    prm.swap = swap;
    prm.in_ostride = prm.out_ostride = (long long)n * inner;
    prm.in_nstride = prm.out_nstride = inner;
    prm.inner = inner;
    if (precision == 8) return emu_tma_dispatch<double>(n, var, prm, outer);
    return emu_tma_dispatch<float>(n, var, prm, outer);
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

// ---- peer-memory put (transfer_put.h): the kernel's index map replayed on host
// memory.  peer_dst are host arrays standing in for the peers' windows.
#include <cstdint>
#include "../../mpi4py_fft_b200/csrc/transfer_put.h"
extern "C" int emu_put(int ndims, const long long* shape, int itemsize, int axisS, int axisD, int p, int rank,
                       const void* src, void* const* peer_dst, int* vec_out) {
    PutParams prm;
    if (put_build(&prm, ndims, shape, itemsize, axisS, axisD, p, rank, src, peer_dst)) return -1;
    if (vec_out) *vec_out = prm.vec;
    for (int k = 0; k < prm.npeers; ++k) {
        const PutPeer& pr = prm.peer[k];
        for (long long u = 0; u < pr.units; ++u) {
            long long su, du;
            put_locate(prm, pr, u, &su, &du);
            std::memcpy(pr.dst + du * prm.vec, prm.src + su * prm.vec, (size_t)prm.vec);
        }
    }
    return 0;
}

// ---- fused redistribution: FFT along axisS of this rank's block, last pass
// storing into the owners' arrays (PeerStore).  shape = group-local shape.
extern "C" int emu_fft_scatter(int precision, int ndims, const long long* shape, int axisS, int axisD, int p, int rank,
                               int var, int staged, const void* in, void* const* peer_dst, double scale, int swap) {
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    if (peer_store_build(&prm.peer, ndims, shape, axisS, axisD, p, rank, peer_dst)) return -2;
    long long nD, sD;
    put_blockdist(shape[axisD], p, rank, &nD, &sD);
    long long outer = 1, inner = 1;
    for (int i = 0; i < axisS; ++i) outer *= (i == axisD ? nD : shape[i]);
    for (int i = axisS + 1; i < ndims; ++i) inner *= (i == axisD ? nD : shape[i]);
    const int n = (int)shape[axisS];
    prm.in = in;
    prm.out = nullptr;
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
    if (staged) {
        if (!strided) return -3;
        return precision == 8 ? emu_tma_dispatch<double>(n, var, prm, outer) : emu_tma_dispatch<float>(n, var, prm, outer);
    }
    if (precision == 8) return emu_dispatch<double>(n, var, strided, prm, outer);
    return emu_dispatch<float>(n, var, strided, prm, outer);
}

// ---- real transforms (fft_pow2.cuh fft_real_body) stepped on the CPU: the same
// phase functions, barriers where the kernel has __syncthreads()
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MODE>
static int emu_real_one(const FftParams& prm_in, long long outer) {
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
    std::vector<C> twv((size_t)RAD::tw_total()), rtw((size_t)N), qtw((size_t)N + 1);
    build_pass_twiddles<T, RAD>(twv.data());
    build_real_twiddles<T>(rtw.data(), N);
    build_quarter_twiddles<T>(qtw.data(), N);
    std::vector<C> t4((size_t)2 * N);
    build_dct4_twiddles<T>(t4.data(), N);
    const C* tw = twv.data();
    std::vector<C> smem((size_t)TF::SI::tile_elems + P);
    std::vector<C> regs((size_t)TF::THREADS * E);
    const long long in_ns = STRIDED ? prm.in_nstride : 1, out_ns = STRIDED ? prm.out_nstride : 1;
    for (long long bid = 0; bid < grid; ++bid) {
        for (auto& x : smem) { x.x = (T)1e30; x.y = (T)-1e30; }
        auto coords = [&](int tid, long long& o, long long& i, bool& valid) {
            const int p = TF::pencil_of(tid);
            if (STRIDED) {
                o = bid / prm.tiles_per_outer;
                i = (bid - o * prm.tiles_per_outer) * P + p;
                valid = i < prm.inner;
            } else {
                o = bid * P + p;
                i = 0;
                valid = o < prm.npencils;
            }
        };
        constexpr int R0 = RAD::get(0);
        constexpr int RL = RAD::get(TF::NPASS - 1);
        if (MODE == 6) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
                TF::r2r4_load(&regs[(size_t)tid * E], TF::slot_of(tid), gin, in_ns, valid, prm.flip != 0, t4.data());
            }
        } else if (MODE == 5) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
                TF::r2r1_load(&regs[(size_t)tid * E], TF::slot_of(tid), gin, in_ns, valid, prm.flip != 0);
            }
        } else if (MODE == 3) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
                TF::r2r_load(&regs[(size_t)tid * E], TF::slot_of(tid), gin, in_ns, valid, prm.flip != 0);
            }
        } else if (MODE == 4) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
                TF::r2r_fill(TF::pencil_of(tid), TF::slot_of(tid), smem.data(), qtw.data(), gin, in_ns, valid, prm.flip != 0);
            }
            // __syncthreads()
            for (int tid = 0; tid < TF::THREADS; ++tid)
                TF::c2r_pre(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), smem.data(), rtw.data());
            // __syncthreads()
        } else if (MODE == 1) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                C* v = &regs[(size_t)tid * E];
                const int q = TF::slot_of(tid);
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                if (STRIDED) {
                    const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
                    for (int b = 0; b < E / R0; ++b)
                        for (int r = 0; r < R0; ++r) {
                            const int n = q + b * TF::TP + r * (N / R0);
                            C a = {(T)0, (T)0};
                            if (valid) { a.x = gin[(long long)(2 * n) * in_ns]; a.y = gin[(long long)(2 * n + 1) * in_ns]; }
                            v[b * R0 + r] = a;
                        }
                } else {
                    TF::load_global(v, q, reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride, 1, valid, false);
                }
            }
        } else {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                const int p = TF::pencil_of(tid), q = TF::slot_of(tid);
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                const C* gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + i;
                const int keep = prm.trunc.n;
                for (int e = 0; e < E; ++e) {
                    const int k = q + e * TF::TP;
                    C a = {(T)0, (T)0};
                    if (valid && (keep == 0 || k < keep)) {
                        a = gin[(long long)k * in_ns];
                        if (keep > 0 && keep % 2 == 0 && k == keep - 1) { a.x *= (T)0.5; a.y = (T)0; }
                    }
                    smem[TF::SI::at(p, k)] = a;
                }
                if (q == 0) {
                    C a = {(T)0, (T)0};
                    if (valid && (keep == 0 || N < keep)) {
                        a = gin[(long long)N * in_ns];
                        if (keep > 0 && keep % 2 == 0 && N == keep - 1) { a.x *= (T)0.5; a.y = (T)0; }
                    }
                    smem[TF::SI::tile_elems + p] = a;
                }
            }
            // __syncthreads()
            for (int tid = 0; tid < TF::THREADS; ++tid)
                TF::c2r_pre(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), smem.data(), rtw.data());
            // __syncthreads()
        }
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            C* v = &regs[(size_t)tid * E];
            TF::template twiddle_dft<0>(v, TF::slot_of(tid), tw);
            if (TF::NPASS > 1) TF::template store_shared<0>(v, TF::pencil_of(tid), TF::slot_of(tid), smem.data());
        }
        if constexpr (TF::NPASS > 1) {
            EmuMid<TF, 1>::run(regs, smem.data(), tw);
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                C* v = &regs[(size_t)tid * E];
                TF::template load_shared<TF::NPASS - 1>(v, TF::pencil_of(tid), TF::slot_of(tid), smem.data());
                TF::template twiddle_dft<TF::NPASS - 1>(v, TF::slot_of(tid), tw);
            }
        }
        if (MODE == 6) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
                TF::r2r4_store(&regs[(size_t)tid * E], TF::slot_of(tid), gout, out_ns, valid, (T)prm.scale, prm.flip != 0, t4.data());
            }
        } else if (MODE == 4) {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
                TF::r2r_store(&regs[(size_t)tid * E], TF::slot_of(tid), gout, out_ns, valid, (T)prm.scale, prm.flip != 0);
            }
        } else if (MODE == 1 || MODE == 3 || MODE == 5) {
            // __syncthreads()
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                const C* v = &regs[(size_t)tid * E];
                const int p = TF::pencil_of(tid), q = TF::slot_of(tid);
                for (int b = 0; b < E / RL; ++b)
                    for (int r = 0; r < RL; ++r) smem[TF::SI::at(p, q + b * TF::TP + r * (N / RL))] = v[b * RL + r];
            }
            // __syncthreads()
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                if (MODE == 5) {
                    T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
                    TF::r2r1_post(TF::pencil_of(tid), TF::slot_of(tid), smem.data(), rtw.data(), gout, out_ns, valid, (T)prm.scale,
                                  prm.flip != 0);
                } else if (MODE == 3) {
                    T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
                    TF::r2r_post(TF::pencil_of(tid), TF::slot_of(tid), smem.data(), rtw.data(), qtw.data(), gout, out_ns, valid,
                                 (T)prm.scale, prm.flip != 0);
                } else {
                    C* gout = reinterpret_cast<C*>(prm.out) + o * prm.out_ostride + i;
                    TF::r2c_post(TF::pencil_of(tid), TF::slot_of(tid), smem.data(), rtw.data(), gout, out_ns, valid, (T)prm.scale, prm.trunc.n);
                }
            }
        } else {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                const C* v = &regs[(size_t)tid * E];
                const int q = TF::slot_of(tid);
                long long o, i; bool valid;
                coords(tid, o, i, valid);
                if (STRIDED) {
                    if (!valid) continue;
                    T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
                    for (int b = 0; b < E / RL; ++b)
                        for (int r = 0; r < RL; ++r) {
                            const int n = q + b * TF::TP + r * (N / RL);
                            gout[(long long)(2 * n) * out_ns] = v[b * RL + r].y * (T)prm.scale;
                            gout[(long long)(2 * n + 1) * out_ns] = v[b * RL + r].x * (T)prm.scale;
                        }
                } else {
                    TF::store_global(v, q, reinterpret_cast<C*>(prm.out) + o * prm.out_ostride, 1, valid, true, (T)prm.scale);
                }
            }
        }
    }
    return 0;
}

#define EMU_REAL_CONTIG(N, E, P, PS, MINB, ...)                                                          \
    if (n == N) return mode == 1 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, 1>(prm, outer) \
                     : mode == 3 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, 3>(prm, outer) \
                     : mode == 4 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, 4>(prm, outer) \
                     : mode == 5 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, 5>(prm, outer) \
                     : mode == 6 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, 6>(prm, outer) \
                                 : emu_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, 2>(prm, outer);
#define EMU_REAL_STRIDED(N, E, P, PS, MINB, ...)                                                         \
    if (n == N)                                                                                          \
        return mode == 1 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS, 1>(prm, outer) \
             : mode == 3 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS, 3>(prm, outer) \
             : mode == 4 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS, 4>(prm, outer) \
             : mode == 5 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS, 5>(prm, outer) \
             : mode == 6 ? emu_real_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS, 6>(prm, outer) \
                         : emu_real_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS, 2>(prm, outer);

template <class T>
static int emu_real_dispatch(int n, int mode, bool strided, const FftParams& prm, long long outer) {
    if (strided) {
        B2F_REAL_STRIDED(EMU_REAL_STRIDED)
    } else {
        B2F_REAL_CONTIG(EMU_REAL_CONTIG)
    }
    return -1;
}

// (outer, nreal, inner) real <-> (outer, nreal/2+1, inner) complex; mode 1 = r2c, 2 = c2r
extern "C" int emu_fft_real(int precision, int nreal, int mode, long long outer, long long inner, const void* in,
                            void* out, double scale) {
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.in = in;
    prm.out = out;
    prm.scale = scale;
    const int nc = nreal / 2;
    const long long n_in = mode == 1 ? nreal : nc + 1, n_out = mode == 1 ? nc + 1 : nreal;
    const bool strided = inner > 1;
    if (strided) {
        prm.in_ostride = n_in * inner;
        prm.out_ostride = n_out * inner;
        prm.in_nstride = prm.out_nstride = inner;
        prm.inner = inner;
    } else {
        prm.in_ostride = mode == 1 ? nc : n_in;
        prm.out_ostride = mode == 1 ? n_out : nc;
        prm.npencils = outer;
    }
    if (precision == 8) return emu_real_dispatch<double>(nc, mode, strided, prm, outer);
    return emu_real_dispatch<float>(nc, mode, strided, prm, outer);
}

// ---- chirp-z kernels (chirpz.cuh) stepped on the CPU -------------------------
#include "../../mpi4py_fft_b200/csrc/chirpz.cuh"
#include "../../mpi4py_fft_b200/csrc/chirpz_host.h"

template <class TF>
static void emu_chirp_fft(std::vector<typename TF::C>& regs, typename TF::C* smem, const typename TF::C* tw) {
    using C = typename TF::C;
    const int E = TF::EPT;
    for (int tid = 0; tid < TF::THREADS; ++tid) {
        C* v = &regs[(size_t)tid * E];
        TF::template twiddle_dft<0>(v, TF::slot_of(tid), tw);
        if (TF::NPASS > 1) TF::template store_shared<0>(v, TF::pencil_of(tid), TF::slot_of(tid), smem);
    }
    if constexpr (TF::NPASS > 1) {
        EmuMid<TF, 1>::run(regs, smem, tw);
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            C* v = &regs[(size_t)tid * E];
            TF::template load_shared<TF::NPASS - 1>(v, TF::pencil_of(tid), TF::slot_of(tid), smem);
            TF::template twiddle_dft<TF::NPASS - 1>(v, TF::slot_of(tid), tw);
        }
    }
}

template <class T, int M, int E, class RAD, int P, bool STRIDED, int PS>
static int emu_chirp_one(const ChirpParams& prm_in, long long outer) {
    using TF = TileFFT<T, M, E, RAD, P, STRIDED, PS>;
    using C = cplx<T>;
    ChirpParams prm = prm_in;
    long long grid;
    if (STRIDED) {
        prm.tiles_per_outer = (prm.inner + P - 1) / P;
        grid = outer * prm.tiles_per_outer;
    } else {
        grid = (prm.npencils + P - 1) / P;
    }
    std::vector<C> twv((size_t)RAD::tw_total());
    build_pass_twiddles<T, RAD>(twv.data());
    std::vector<C> smem((size_t)TF::SI::tile_elems);
    std::vector<C> regs((size_t)TF::THREADS * E);
    const int in_size = prm.in_mode == 1 ? (int)sizeof(T) : (int)sizeof(C);
    const int out_size = prm.out_real ? (int)sizeof(T) : (int)sizeof(C);
    for (long long bid = 0; bid < grid; ++bid) {
        for (auto& x : smem) { x.x = (T)1e30; x.y = (T)-1e30; }
        auto coords = [&](int tid, long long& o, long long& i, bool& valid) {
            const int p = TF::pencil_of(tid);
            if (STRIDED) {
                o = bid / prm.tiles_per_outer;
                i = (bid - o * prm.tiles_per_outer) * P + p;
                valid = i < prm.inner;
            } else {
                o = bid * P + p;
                i = 0;
                valid = o < prm.npencils;
            }
        };
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            long long o, i; bool valid;
            coords(tid, o, i, valid);
            const char* gin = reinterpret_cast<const char*>(prm.in) + (o * prm.in_ostride + i) * in_size;
            chirp_phase_load<TF>(&regs[(size_t)tid * E], TF::slot_of(tid), gin, prm.in_nstride, valid, prm);
        }
        emu_chirp_fft<TF>(regs, smem.data(), twv.data());
        // __syncthreads()
        for (int tid = 0; tid < TF::THREADS; ++tid)
            chirp_phase_filter<TF>(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), smem.data(), prm);
        // __syncthreads()
        for (int tid = 0; tid < TF::THREADS; ++tid)
            chirp_phase_reload<TF>(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), smem.data());
        // __syncthreads()
        emu_chirp_fft<TF>(regs, smem.data(), twv.data());
        for (int tid = 0; tid < TF::THREADS; ++tid) {
            long long o, i; bool valid;
            coords(tid, o, i, valid);
            char* gout = reinterpret_cast<char*>(prm.out) + (o * prm.out_ostride + i) * out_size;
            chirp_phase_store<TF>(&regs[(size_t)tid * E], TF::slot_of(tid), gout, prm.out_nstride, valid, prm);
        }
    }
    return 0;
}

#define EMU_CHIRP_CONTIG(N, E, P, PS, MINB, ...) \
    if (m == N) return emu_chirp_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS>(prm, outer);
#define EMU_CHIRP_STRIDED(N, E, P, PS, MINB, ...) \
    if (m == N) return emu_chirp_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), true, PS>(prm, outer);

template <class T>
static int emu_chirp_dispatch(int m, bool strided, const ChirpParams& prm, long long outer) {
    if (strided) {
        B2F_REAL_STRIDED_POW2(EMU_CHIRP_STRIDED)
    } else {
        B2F_REAL_CONTIG_POW2(EMU_CHIRP_CONTIG)
    }
    return -1;
}

// (outer, n_stored_in, inner) -> (outer, n_stored_out, inner) along the middle axis;
// kind = FFTW integer, n = logical length
extern "C" int emu_chirpz(int precision, int kind, long long n, long long outer, long long inner, const void* in,
                          void* out, double scale) {
    ChirpSpec sp;
    if (chirp_build(kind, n, &sp)) return -2;
    ChirpParams prm;
    std::memset(&prm, 0, sizeof(prm));
    std::vector<double> pre8, filt8, post8;
    std::vector<float> pre4, filt4, post4;
    auto conv = [](const std::vector<long double>& v, std::vector<double>& d, std::vector<float>& f) {
        d.resize(v.size());
        f.resize(v.size());
        for (size_t i = 0; i < v.size(); ++i) { d[i] = (double)v[i]; f[i] = (float)v[i]; }
    };
    conv(sp.pre, pre8, pre4);
    conv(sp.filt, filt8, filt4);
    conv(sp.post, post8, post4);
    prm.in = in;
    prm.out = out;
    prm.pre = precision == 8 ? (const void*)pre8.data() : (const void*)pre4.data();
    prm.filt = precision == 8 ? (const void*)filt8.data() : (const void*)filt4.data();
    prm.post = precision == 8 ? (const void*)post8.data() : (const void*)post4.data();
    prm.n_in = (int)sp.n_in;
    prm.n_out = (int)sp.n_out;
    prm.in_mode = sp.in_mode;
    prm.out_real = sp.out_real;
    prm.scale = scale;
    const long long stored_in = sp.in_mode == 2 ? n / 2 + 1 : sp.n_in;
    prm.in_ostride = stored_in * inner;
    prm.out_ostride = sp.n_out * inner;
    prm.in_nstride = prm.out_nstride = inner;
    prm.inner = inner;
    prm.npencils = outer;
    const bool strided = inner > 1;
    if (precision == 8) return emu_chirp_dispatch<double>(sp.M, strided, prm, outer);
    return emu_chirp_dispatch<float>(sp.M, strided, prm, outer);
}

// ---- partial launches of a fused stage (capi.cu run_plan + ChunkSpec restated for
// the test): the same pointer shifts / extents / PeerStore offsets the library
// applies, then the kernels' own per-thread code.  mode 1 = inner range (with
// view_outer > 0: rows view_ostride apart, transformed axis first), 2 = outer range.
extern "C" int emu_fft_scatter_chunk(int precision, int ndims, const long long* shape, int axisS, int axisD, int p,
                                     int rank, const void* in, void* const* peer_dst, double scale, int swap,
                                     int mode, long long begin, long long count, long long view_outer,
                                     long long view_ostride) {
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    if (peer_store_build(&prm.peer, ndims, shape, axisS, axisD, p, rank, peer_dst)) return -2;
    long long nD, sD;
    put_blockdist(shape[axisD], p, rank, &nD, &sD);
    long long outer = 1, inner = 1;
    for (int i = 0; i < axisS; ++i) outer *= (i == axisD ? nD : shape[i]);
    for (int i = axisS + 1; i < ndims; ++i) inner *= (i == axisD ? nD : shape[i]);
    const int n = (int)shape[axisS];
    const bool strided = inner > 1;
    prm.scale = scale;
    prm.swap = swap;
    if (strided) {
        prm.in_ostride = prm.out_ostride = (long long)n * inner;
        prm.in_nstride = prm.out_nstride = inner;
        prm.inner = inner;
    } else {
        prm.in_ostride = prm.out_ostride = n;
        prm.npencils = outer;
    }
    const long long esz = 2LL * precision;
    long long off;
    if (mode == 2) {
        if (begin + count > outer || view_outer) return -4;
        off = begin * prm.in_ostride;
        outer = count;
        prm.npencils = outer;
        prm.peer.ooff = begin;
    } else {
        if (!strided || begin + count > inner) return -4;
        off = begin;
        prm.inner = count;
        prm.peer.ioff = begin;
        if (view_outer > 0) {
            if (outer != 1) return -4;
            outer = view_outer;
            prm.in_ostride = prm.out_ostride = view_ostride;
            prm.peer.vstride = view_ostride;
        }
    }
    prm.in = (const char*)in + off * esz;
    prm.out = nullptr;
    if (outer == 0 || (strided && prm.inner == 0)) return 0;
    if (precision == 8) return emu_dispatch<double>(n, 0, strided, prm, outer);
    return emu_dispatch<float>(n, 0, strided, prm, outer);
}

// ---- rotating kernels (fft_rot.cuh) stepped on the CPU: the bulk copies are
// emulated by memcpy of whole pencils into the padded stage, everything after is
// the kernel's own per-thread code; the schedule (which strides each of the three
// steps uses) is the library's own build_rotation (rot_plan.h).
#include "../../mpi4py_fft_b200/csrc/fft_rot.cuh"
#include "../../mpi4py_fft_b200/csrc/rot_plan.h"

struct EmuRotStep {
    const void* in;
    void* out;
    long long batches, I, O, in_i, in_o, in_b, out_o, out_n, out_b;
    double scale;
    int swap;
};

template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int OPT>
static int emu_rot_one(const EmuRotStep& st) {
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    using C = cplx<T>;
    constexpr int PITCH = N + RotPad<T, P>::value;
    const long long tpo = (st.I + P - 1) / P, tpb = st.O * tpo, ntiles = st.batches * tpb;
    std::vector<C> twv((size_t)RAD::tw_total());
    build_pass_twiddles<T, RAD>(twv.data());
    std::vector<C> stage((size_t)PITCH * P);
    std::vector<typename EX::X> xbuf((size_t)TF::SI::tile_elems);
    std::vector<C> regs((size_t)TF::THREADS * E);
    const C* gin = reinterpret_cast<const C*>(st.in);
    C* gout = reinterpret_cast<C*>(st.out);
    for (long long t = 0; t < ntiles; ++t) {
        const long long b = t / tpb, rem = t - b * tpb, o = rem / tpo, i0 = (rem - o * tpo) * P;
        for (auto& x : stage) { x.x = (T)1e30; x.y = (T)-1e30; }
        for (int j = 0; j < P && i0 + j < st.I; ++j)
            std::memcpy(&stage[(size_t)j * PITCH], gin + b * st.in_b + o * st.in_o + (i0 + j) * st.in_i, sizeof(C) * N);
        for (int tid = 0; tid < TF::THREADS; ++tid)
            load_rows<TF, PITCH>(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), stage.data(), st.swap != 0);
        for (auto& x : xbuf) x = typename EX::X{};
        for (int tid = 0; tid < TF::THREADS; ++tid)
            TF::template twiddle_dft<0>(&regs[(size_t)tid * E], TF::slot_of(tid), twv.data());
        EmuTmaMid<TF, EX, 1>::run(regs, xbuf.data(), twv.data(), SPLIT);
        if constexpr ((OPT & 8) != 0) {
            // TMA-store flavour: chunks of rows staged in the exchange buffer, then box stores
            // (emulated by a dense copy clipped at the tensor's extent I)
            using OC = OutChunks<TF, EX>;
            C* xo = reinterpret_cast<C*>(xbuf.data());
            static_assert(sizeof(typename EX::X) * TF::SI::tile_elems >= sizeof(C) * (size_t)OC::rows * P, "staging fits");
            for (int c = 0; c < OC::count; ++c) {
                // __syncthreads()
                for (int k = 0; k < OC::rows * P; ++k) { xo[k].x = (T)1e30; xo[k].y = (T)-1e30; }
                for (int tid = 0; tid < TF::THREADS; ++tid)
                    stage_out_rows<TF, OC::count>(&regs[(size_t)tid * E], TF::pencil_of(tid), TF::slot_of(tid), xo, c,
                                                  st.swap != 0, (T)st.scale);
                // __syncthreads(); boxes of box_rows rows
                for (int k = 0; k < OC::rows; ++k)
                    for (int pp = 0; pp < P && i0 + pp < st.I; ++pp)
                        gout[b * st.out_b + o * st.out_o + (long long)(c * OC::rows + k) * st.out_n + i0 + pp] = xo[(size_t)k * P + pp];
            }
        } else {
            for (int tid = 0; tid < TF::THREADS; ++tid) {
                const long long i = i0 + TF::pencil_of(tid);
                TF::store_global(&regs[(size_t)tid * E], TF::slot_of(tid), gout + b * st.out_b + o * st.out_o + i, st.out_n,
                                 i < st.I, st.swap != 0, (T)st.scale);
            }
        }
    }
    return 0;
}

#define EMU_ROT(N, VAR, E, P, PS, STAGES, SPLIT, MINB, ...) \
    if (n == N && var == VAR)                                \
        return emu_rot_one<T, N, E, Radices<__VA_ARGS__>, P * (int)(sizeof(double) / sizeof(T)), PS, STAGES, SPLIT != 0, ((MINB) >> 4)>(st);

template <class T>
static int emu_rot_dispatch(int n, int var, const EmuRotStep& st) {
    B2F_ROT_TABLE(EMU_ROT)
    return -1;
}

// the whole 3-step schedule of a block (ndims >= 3, axes = its last three in any order);
// var < 0: the default variant of each length.  -1: schedule does not apply / variant not built.
extern "C" int emu_rot_plan(int precision, int ndims, const long long* sizes, const int* axes, int var, const void* in,
                            void* out, void* scratch, double scale, int swap) {
    std::vector<RotPlanStep> steps;
    long long elems = 0;
    if (!build_rotation(ndims, sizes, axes, 3, &steps, &elems)) return -1;
    void* bufs[3] = {const_cast<void*>(in), out, scratch};
    for (size_t si = 0; si < steps.size(); ++si) {
        const RotPlanStep& r = steps[si];
        EmuRotStep st{bufs[r.src], bufs[r.dst], r.batches, r.I, r.O, r.in_i, r.in_o, r.in_b, r.out_o, r.out_n, r.out_b,
                      si + 1 == steps.size() ? scale : 1.0, swap};
        int rc;
        if (var >= 1000) {
            // register-path engine: the strided kernels with whole pencils in (in_istride), rows out
            FftParams prm;
            std::memset(&prm, 0, sizeof(prm));
            prm.scale = st.scale;
            prm.swap = swap;
            prm.in_ostride = r.in_o;
            prm.out_ostride = r.out_o;
            prm.in_nstride = 1;
            prm.out_nstride = r.out_n;
            prm.in_istride = r.in_i;
            prm.inner = r.I;
            rc = 0;
            for (long long b = 0; b < r.batches && rc == 0; ++b) {
                prm.in = (const char*)st.in + b * r.in_b * 2 * precision;
                prm.out = (char*)st.out + b * r.out_b * 2 * precision;
                rc = precision == 8 ? emu_dispatch<double>(r.n, var - 1000, true, prm, r.O)
                                    : emu_dispatch<float>(r.n, var - 1000, true, prm, r.O);
            }
        } else {
            const int v = var >= 0 ? var : rot_default(r.n);
            rc = precision == 8 ? emu_rot_dispatch<double>(r.n, v, st) : emu_rot_dispatch<float>(r.n, v, st);
        }
        if (rc) return rc;
    }
    return 0;
}

// one rotating step on its own: in [B][I][O][n] -> out [B][O][n][I]
extern "C" int emu_rot_step(int precision, int n, int var, long long batches, long long I, long long O, const void* in,
                            void* out, double scale, int swap) {
    EmuRotStep st{in, out, batches, I, O, O * n, n, I * O * n, (long long)n * I, I, I * O * n, scale, swap};
    return precision == 8 ? emu_rot_dispatch<double>(n, var, st) : emu_rot_dispatch<float>(n, var, st);
}


// ---- dealiasing folded into a stage (TruncMap, fft_core.cuh): (outer, n_pad, inner) physical side,
// (outer, keep, inner) spectrum side.  kind: -1 forward c2c (truncating store), +1 backward c2c
// (padding load), -2 r2c, +2 c2r; n = padded logical length.  Strides as capi.cu run_plan sets them.
extern "C" int emu_fft_trunc(int precision, int kind, int n, int keep, long long outer, long long inner, const void* in,
                             void* out, double scale) {
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.in = in;
    prm.out = out;
    prm.scale = scale;
    const bool strided = inner > 1;
    if (kind == -1 || kind == 1) {
        prm.swap = kind == 1;
        prm.trunc.n = keep;
        prm.trunc.np = n;
        if (strided) {
            prm.in_ostride = prm.out_ostride = (long long)n * inner;
            prm.in_nstride = prm.out_nstride = inner;
            prm.inner = inner;
            (prm.swap ? prm.in_ostride : prm.out_ostride) = (long long)keep * inner;
        } else {
            prm.in_ostride = prm.out_ostride = n;
            prm.npencils = outer;
            (prm.swap ? prm.in_ostride : prm.out_ostride) = keep;
        }
        if (precision == 8) return emu_dispatch<double>(n, 0, strided, prm, outer);
        return emu_dispatch<float>(n, 0, strided, prm, outer);
    }
    const int mode = kind == -2 ? 1 : 2;
    const int nc = n / 2;
    const long long n_in = mode == 1 ? n : nc + 1, n_out = mode == 1 ? nc + 1 : n;
    prm.trunc.n = keep;
    prm.trunc.np = nc + 1;
    if (strided) {
        prm.in_ostride = n_in * inner;
        prm.out_ostride = n_out * inner;
        prm.in_nstride = prm.out_nstride = inner;
        prm.inner = inner;
        (mode == 1 ? prm.out_ostride : prm.in_ostride) = (long long)keep * inner;
    } else {
        prm.in_ostride = mode == 1 ? nc : n_in;
        prm.out_ostride = mode == 1 ? n_out : nc;
        prm.npencils = outer;
        (mode == 1 ? prm.out_ostride : prm.in_ostride) = keep;
    }
    if (precision == 8) return emu_real_dispatch<double>(nc, mode, strided, prm, outer);
    return emu_real_dispatch<float>(nc, mode, strided, prm, outer);
}


// ---- r2r kinds II / III through the real-transform kernels (fft_core.cuh r2r_*): (outer, n, inner) real in and out;
// kind = FFTW integer 5 (REDFT10), 4 (REDFT01), 9 (RODFT10), 8 (RODFT01).  Strides as capi.cu run_plan sets them.
extern "C" int emu_fft_r2r(int precision, int kind, int n, long long outer, long long inner, const void* in, void* out,
                           double scale) {
    if (kind == 3 || kind == 7) {
        // kinds I: N = n - 1 (REDFT00) / n + 1 (RODFT00) complex points
        FftParams prm;
        std::memset(&prm, 0, sizeof(prm));
        prm.in = in;
        prm.out = out;
        prm.scale = scale;
        prm.flip = kind == 7;
        const bool strided = inner > 1;
        if (strided) {
            prm.in_ostride = prm.out_ostride = (long long)n * inner;
            prm.in_nstride = prm.out_nstride = inner;
            prm.inner = inner;
        } else {
            prm.in_ostride = prm.out_ostride = n;
            prm.npencils = outer;
        }
        const int nc = kind == 3 ? n - 1 : n + 1;
        if (precision == 8) return emu_real_dispatch<double>(nc, 5, strided, prm, outer);
        return emu_real_dispatch<float>(nc, 5, strided, prm, outer);
    }
    if (n % 2 || (kind != 5 && kind != 4 && kind != 9 && kind != 8 && kind != 6 && kind != 10)) return -2;
    FftParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.in = in;
    prm.out = out;
    prm.scale = scale;
    prm.flip = (kind == 9 || kind == 8 || kind == 10) ? 1 : 0;
    const int mode = (kind == 6 || kind == 10) ? 6 : (kind == 5 || kind == 9) ? 3 : 4;
    const bool strided = inner > 1;
    if (strided) {
        prm.in_ostride = prm.out_ostride = (long long)n * inner;
        prm.in_nstride = prm.out_nstride = inner;
        prm.inner = inner;
    } else {
        prm.in_ostride = prm.out_ostride = n;
        prm.npencils = outer;
    }
    if (precision == 8) return emu_real_dispatch<double>(n / 2, mode, strided, prm, outer);
    return emu_real_dispatch<float>(n / 2, mode, strided, prm, outer);
}

// ---- four-step split of a c2c length beyond one tile (csrc/lengths.h fourstep_split, capi.cu STEP_FOURSTEP):
// (outer, n, inner) block, n = n1 * n2.  Step 1: n2-point transforms along j2 of the (outer, n2, n1 * inner) view,
// in -> scratch; step 2: twiddle; step 3: n1-point transforms stored k1-major -- the rotating kernel when the axis
// is contiguous, the strided kernel with transposed output strides otherwise.  Same parameters as capi.cu.
#include "../../mpi4py_fft_b200/csrc/lengths.h"
template <class T>
static int emu_fourstep_t(int precision, long long n, long long outer, long long inner, const void* in, void* out,
                          void* scratch, double scale, int swap, long long* n1_out, long long* n2_out) {
    FourStep fs;
    if (is_stockham(n) || !fourstep_split(n, &fs)) return -1;
    *n1_out = fs.n1;
    *n2_out = fs.n2;
    const long long n1 = fs.n1, n2 = fs.n2;
    // 1
    FftParams p1;
    std::memset(&p1, 0, sizeof(p1));
    p1.in = in;
    p1.out = scratch;
    p1.scale = 1.0;
    p1.swap = swap;
    p1.in_ostride = p1.out_ostride = n2 * n1 * inner;
    p1.in_nstride = p1.out_nstride = n1 * inner;
    p1.inner = n1 * inner;
    int rc = emu_dispatch<T>((int)n2, 0, true, p1, outer);
    if (rc) return rc;
    // 2
    cplx<T>* sc = reinterpret_cast<cplx<T>*>(scratch);
    const long long total = outer * n2 * n1 * inner;
    for (long long idx = 0; idx < total; ++idx) {
        long long t = idx / inner;
        const long long j1 = t % n1;
        t /= n1;
        const long long k2 = t % n2;
        sc[idx] = cmul(sc[idx], fourstep_twiddle<T>((j1 * k2) % n, n, swap != 0));
    }
    // 3
    if (inner == 1) {
        EmuRotStep st{scratch, out, outer, n2, 1, n1, n1, n, n, n2, n, scale, swap};
        return emu_rot_dispatch<T>((int)n1, rot_default((int)n1), st);
    }
    FftParams p3;
    std::memset(&p3, 0, sizeof(p3));
    p3.scale = scale;
    p3.swap = swap;
    p3.in_ostride = n1 * inner;
    p3.in_nstride = inner;
    p3.out_ostride = inner;
    p3.out_nstride = n2 * inner;
    p3.inner = inner;
    for (long long o = 0; o < outer; ++o) {
        p3.in = (const char*)scratch + (size_t)o * n * inner * 2 * precision;
        p3.out = (char*)out + (size_t)o * n * inner * 2 * precision;
        rc = emu_dispatch<T>((int)n1, 0, true, p3, n2);
        if (rc) return rc;
    }
    return 0;
}
extern "C" int emu_fourstep(int precision, long long n, long long outer, long long inner, const void* in, void* out,
                            void* scratch, double scale, int swap, long long* n1_out, long long* n2_out) {
    if (precision == 8) return emu_fourstep_t<double>(precision, n, outer, inner, in, out, scratch, scale, swap, n1_out, n2_out);
    return emu_fourstep_t<float>(precision, n, outer, inner, in, out, scratch, scale, swap, n1_out, n2_out);
}
