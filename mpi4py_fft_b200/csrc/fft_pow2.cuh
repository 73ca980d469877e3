// fft_pow2.cuh -- batched power-of-two c2c Stockham kernels (sm_100a).
//
// One CTA transforms a tile of P pencils x N points; every thread keeps E points
// in registers, the first radix pass reads HBM directly (coalesced 16-byte
// loads: consecutive threads -> consecutive addresses), the passes in between
// exchange through padded shared memory, and the last pass writes HBM directly
// with the normalisation fused in.  A pencil is read once and written once --
// the algorithmic minimum for one axis -- and no local transpose is ever
// materialised for strided axes (tile = N x P contiguous elements).
//
// The backward transform is the forward kernel with real and imaginary parts
// swapped on load and on store:  B(x) = swap(F(swap(x))).
#pragma once
#include "fft_core.cuh"

namespace b2f {

struct FftParams {
    const void* in;
    void* out;
    const void* tw;          // N forward twiddles exp(-2 pi i k / N)
    long long in_ostride;    // elements between consecutive outer indices (CONTIG: row pitch)
    long long out_ostride;
    long long in_nstride;    // elements between consecutive points of a pencil (STRIDED)
    long long out_nstride;
    long long in_istride;    // STRIDED: elements between consecutive inner indices on the INPUT side (0 = 1: the
                             // plain layout; > 1 with in_nstride = 1: whole pencils in, rotated rows out, rot_plan.h)
    long long inner;         // STRIDED: extent of the contiguous inner index
    long long npencils;      // CONTIG: number of rows
    long long tiles_per_outer;
    double scale;
    int swap;
    PeerStore peer;          // peer.p > 0: the last pass stores into the owners' arrays (fused redistribution)
    const void* rtw;         // real transforms: exp(-2 pi i k / 2N), k < N
    TruncMap trunc;          // trunc.n > 0: the spectrum side holds only trunc.n modes (padded transform, fft_core.cuh)
    const void* qtw;         // r2r kinds II / III: exp(-i pi k / 4N), k <= N
    int flip;                // r2r: sine kinds (RODFT10 / RODFT01)
};

#if defined(__CUDACC__)

template <class TF, int S>
struct MidPasses {
    using C = typename TF::C;
    static __device__ __forceinline__ void run(C* v, int p, int q, C* smem, const C* __restrict__ tw) {
        if constexpr (S < TF::NPASS - 1) {
            TF::template load_shared<S>(v, p, q, smem);
            TF::template twiddle_dft<S>(v, q, tw);
            __syncthreads();
            TF::template store_shared<S>(v, p, q, smem);
            __syncthreads();
            MidPasses<TF, S + 1>::run(v, p, q, smem, tw);
        }
    }
};

// SWAP is a compile-time constant inside the body so that the re/im exchange
// of the backward transform costs no register moves around the 16-byte
// loads/stores (the kernel branches once on prm.swap).
// TRUNC: dealiasing folded in (TruncMap): the forward transform (SWAP = false) stores only the
// kept modes, the backward transform (SWAP = true) reads the kept modes and zeros for the rest.
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, bool SWAP, bool PEER, bool TRUNC = false>
__device__ __forceinline__ void fft_pow2_body(const FftParams& prm) {
    using TF = TileFFT<T, N, E, RAD, P, STRIDED, PS>;
    using C = cplx<T>;
    extern __shared__ __align__(16) unsigned char b2f_smem_raw[];
    C* smem = reinterpret_cast<C*>(b2f_smem_raw);

    const int tid = threadIdx.x;
    const int p = TF::pencil_of(tid);
    const int q = TF::slot_of(tid);

    const C* gin;
    C* gout;
    long long in_ns, out_ns;
    long long po, pi;   // pencil coordinates (outer, inner) for the fused peer store
    bool valid;
    if (STRIDED) {
        const long long bid = blockIdx.x;
        const long long o = bid / prm.tiles_per_outer;
        const long long i = (bid - o * prm.tiles_per_outer) * P + p;
        valid = i < prm.inner;
        gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + (prm.in_istride ? i * prm.in_istride : i);
        gout = reinterpret_cast<C*>(prm.out) + o * prm.out_ostride + i;
        in_ns = prm.in_nstride;
        out_ns = prm.out_nstride;
        po = o;
        pi = i;
    } else {
        const long long gp = (long long)blockIdx.x * P + p;
        valid = gp < prm.npencils;
        gin = reinterpret_cast<const C*>(prm.in) + gp * prm.in_ostride;
        gout = reinterpret_cast<C*>(prm.out) + gp * prm.out_ostride;
        in_ns = 1;
        out_ns = 1;
        po = gp;
        pi = 0;
    }
    const C* __restrict__ tw = reinterpret_cast<const C*>(prm.tw);

    C v[E];
    if constexpr (TRUNC && SWAP) TF::load_global_padded(v, q, gin, in_ns, valid, SWAP, prm.trunc);
    else TF::load_global(v, q, gin, in_ns, valid, SWAP);
    TF::template twiddle_dft<0>(v, q, tw);
    if constexpr (TF::NPASS > 1) {
        TF::template store_shared<0>(v, p, q, smem);
        __syncthreads();
        MidPasses<TF, 1>::run(v, p, q, smem, tw);
        TF::template load_shared<TF::NPASS - 1>(v, p, q, smem);
        TF::template twiddle_dft<TF::NPASS - 1>(v, q, tw);
    }
    if constexpr (TRUNC && !SWAP) {
        // the upper copy of an even-N Nyquist mode travels through P extra slots behind the tile
        C* nyq = smem + (TF::NPASS > 1 ? TF::SI::tile_elems : 0);
        TF::store_truncated_publish(v, p, q, nyq, prm.trunc, (T)prm.scale);
        __syncthreads();
        TF::store_truncated(v, p, q, gout, out_ns, valid, SWAP, (T)prm.scale, nyq, prm.trunc);
    } else if constexpr (PEER) {
        long long part = 0, rest = 0;
        if (valid) prm.peer.locate(po, pi, &part, &rest);
        TF::store_peer(v, q, prm.peer, part, rest, valid, SWAP, (T)prm.scale);
    } else {
        TF::store_global(v, q, gout, out_ns, valid, SWAP, (T)prm.scale);
    }
}

template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
__global__ void __launch_bounds__((N / E) * P, MINB) fft_pow2_kernel(const FftParams prm) {
    if (prm.swap) fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, true, false>(prm);
    else fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, false, false>(prm);
}

// dealiasing folded in: forward truncates on store, backward pads on load (optionally with the
// fused redistribution store: the backward transform of a padded stage can scatter)
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
__global__ void __launch_bounds__((N / E) * P, MINB) fft_pow2_trunc_kernel(const FftParams prm) {
    if (prm.swap) fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, true, false, true>(prm);
    else fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, false, false, true>(prm);
}
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
__global__ void __launch_bounds__((N / E) * P, MINB) fft_pow2_trunc_peer_kernel(const FftParams prm) {
    fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, true, true, true>(prm);
}

// the same transform with the last pass storing into the owners' arrays (PeerStore)
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
__global__ void __launch_bounds__((N / E) * P, MINB) fft_pow2_peer_kernel(const FftParams prm) {
    if (prm.swap) fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, true, true>(prm);
    else fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, false, true>(prm);
}


// ---------------------------------------------------------------------------
// Real transforms of even length 2N on the same tile machinery (MODE 1 = r2c,
// 2 = c2r): the pencil is packed into N complex points, transformed by the
// N-point schedule, and split / merged through shared memory (TileFFT::r2c_post,
// c2r_pre).  Strides in FftParams are in elements of the respective array:
//   r2c: in real (outer, 2N, inner), out complex (outer, N+1, inner)
//   c2r: in complex (outer, N+1, inner), out real (outer, 2N, inner)
// ---------------------------------------------------------------------------
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MODE>
__device__ __forceinline__ void fft_real_body(const FftParams& prm) {
    using TF = TileFFT<T, N, E, RAD, P, STRIDED, PS>;
    using C = cplx<T>;
    extern __shared__ __align__(16) unsigned char b2f_smem_raw[];
    C* smem = reinterpret_cast<C*>(b2f_smem_raw);
    const int tid = threadIdx.x;
    const int p = TF::pencil_of(tid);
    const int q = TF::slot_of(tid);
    long long o, i;
    bool valid;
    if (STRIDED) {
        const long long bid = blockIdx.x;
        o = bid / prm.tiles_per_outer;
        i = (bid - o * prm.tiles_per_outer) * P + p;
        valid = i < prm.inner;
    } else {
        o = (long long)blockIdx.x * P + p;
        i = 0;
        valid = o < prm.npencils;
    }
    const C* __restrict__ tw = reinterpret_cast<const C*>(prm.tw);
    const C* __restrict__ rtw = reinterpret_cast<const C*>(prm.rtw);
    const long long in_ns = STRIDED ? prm.in_nstride : 1, out_ns = STRIDED ? prm.out_nstride : 1;
    const C* __restrict__ qtw = reinterpret_cast<const C*>(prm.qtw);
    C v[E];
    if constexpr (MODE == 6) {
        // DCT-IV / DST-IV: pre-twiddled pairs (x[2m], x[n-1-2m]), no real-transform split afterwards
        const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
        TF::r2r4_load(v, q, gin, in_ns, valid, prm.flip != 0, qtw);
    } else if constexpr (MODE == 5) {
        // DCT-I / DST-I: even / odd extension of the n = N + 1 (N - 1) input points, read through an index map
        const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
        TF::r2r1_load(v, q, gin, in_ns, valid, prm.flip != 0);
    } else if constexpr (MODE == 3) {
        // DCT-II / DST-II: real (outer, 2N, inner) in, permuted pairs
        const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
        TF::r2r_load(v, q, gin, in_ns, valid, prm.flip != 0);
    } else if constexpr (MODE == 4) {
        // DCT-III / DST-III: the half spectrum is built from the real input in shared memory
        const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
        TF::r2r_fill(p, q, smem, qtw, gin, in_ns, valid, prm.flip != 0);
        __syncthreads();
        TF::c2r_pre(v, p, q, smem, rtw);
        __syncthreads();
    } else if constexpr (MODE == 1) {
        constexpr int R0 = RAD::get(0);
        if (STRIDED) {
            const T* gin = reinterpret_cast<const T*>(prm.in) + o * prm.in_ostride + i;
#pragma unroll
            for (int b = 0; b < E / R0; ++b)
#pragma unroll
                for (int r = 0; r < R0; ++r) {
                    const int n = q + b * TF::TP + r * (N / R0);
                    C a = {(T)0, (T)0};
                    if (valid) {
                        a.x = gin[(long long)(2 * n) * in_ns];
                        a.y = gin[(long long)(2 * n + 1) * in_ns];
                    }
                    v[b * R0 + r] = a;
                }
        } else {
            const C* gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride;
            TF::load_global(v, q, gin, 1, valid, false);
        }
    } else {
        // c2r: the half spectrum goes to shared memory first (X[N] in the extra slot)
        // padded transform (keep > 0): the input holds only the first `keep` modes, the last of
        // them real and halved when keep is even (reference libfft.py:289-298); the rest is zero
        const C* gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + i;
        const int keep = prm.trunc.n;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = q + e * TF::TP;
            // unconditional load (a dropped mode re-reads mode 0), value selected afterwards: all E
            // loads of the thread stay in flight together
            const bool have = keep == 0 || k < keep;
            C a = {(T)0, (T)0};
            if (valid) a = gin[(long long)(have ? k : 0) * in_ns];
            if (!have) a = {(T)0, (T)0};
            if (keep > 0 && keep % 2 == 0 && k == keep - 1) {
                a.x *= (T)0.5;
                a.y = (T)0;
            }
            smem[TF::SI::at(p, k)] = a;
        }
        if (q == 0) {
            C a = {(T)0, (T)0};
            if (valid && (keep == 0 || N < keep)) {
                a = gin[(long long)N * in_ns];
                if (keep > 0 && keep % 2 == 0 && N == keep - 1) {
                    a.x *= (T)0.5;
                    a.y = (T)0;
                }
            }
            smem[TF::SI::tile_elems + p] = a;
        }
        __syncthreads();
        TF::c2r_pre(v, p, q, smem, rtw);
        __syncthreads();   // every thread has its inputs: the tile may be overwritten
    }
    TF::template twiddle_dft<0>(v, q, tw);
    if constexpr (TF::NPASS > 1) {
        TF::template store_shared<0>(v, p, q, smem);
        __syncthreads();
        MidPasses<TF, 1>::run(v, p, q, smem, tw);
        TF::template load_shared<TF::NPASS - 1>(v, p, q, smem);
        TF::template twiddle_dft<TF::NPASS - 1>(v, q, tw);
    }
    if constexpr (MODE == 6) {
        // in place: every thread of the tile must have read its inputs before anybody stores -- with one
        // pass nothing else separates the two
        if constexpr (TF::NPASS == 1) __syncthreads();
        T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
        TF::r2r4_store(v, q, gout, out_ns, valid, (T)prm.scale, prm.flip != 0, qtw);
    } else if constexpr (MODE == 4) {
        T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
        TF::r2r_store(v, q, gout, out_ns, valid, (T)prm.scale, prm.flip != 0);
    } else if constexpr (MODE == 1 || MODE == 3 || MODE == 5) {
        // Z in natural order -> shared memory -> split into the half spectrum
        if constexpr (TF::NPASS > 1) __syncthreads();   // the last pass has read the tile
        constexpr int RL = RAD::get(TF::NPASS - 1);
#pragma unroll
        for (int b = 0; b < E / RL; ++b)
#pragma unroll
            for (int r = 0; r < RL; ++r) smem[TF::SI::at(p, q + b * TF::TP + r * (N / RL))] = v[b * RL + r];
        __syncthreads();
        if constexpr (MODE == 3) {
            T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
            TF::r2r_post(p, q, smem, rtw, qtw, gout, out_ns, valid, (T)prm.scale, prm.flip != 0);
        } else if constexpr (MODE == 5) {
            T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
            TF::r2r1_post(p, q, smem, rtw, gout, out_ns, valid, (T)prm.scale, prm.flip != 0);
        } else {
            C* gout = reinterpret_cast<C*>(prm.out) + o * prm.out_ostride + i;
            TF::r2c_post(p, q, smem, rtw, gout, out_ns, valid, (T)prm.scale, prm.trunc.n);
        }
    } else {
        if (STRIDED) {
            if (valid) {
                T* gout = reinterpret_cast<T*>(prm.out) + o * prm.out_ostride + i;
                constexpr int RL = RAD::get(TF::NPASS - 1);
                const T sc = (T)prm.scale;
#pragma unroll
                for (int b = 0; b < E / RL; ++b)
#pragma unroll
                    for (int r = 0; r < RL; ++r) {
                        const int n = q + b * TF::TP + r * (N / RL);
                        // swapped back: x[2n] = im part of the forward result, x[2n+1] = re part
                        gout[(long long)(2 * n) * out_ns] = v[b * RL + r].y * sc;
                        gout[(long long)(2 * n + 1) * out_ns] = v[b * RL + r].x * sc;
                    }
            }
        } else {
            C* gout = reinterpret_cast<C*>(prm.out) + o * prm.out_ostride;
            TF::store_global(v, q, gout, 1, valid, true, (T)prm.scale);
        }
    }
}

template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB, int MODE>
__global__ void __launch_bounds__((N / E) * P, MINB) fft_real_kernel(const FftParams prm) {
    fft_real_body<T, N, E, RAD, P, STRIDED, PS, MODE>(prm);
}

#endif  // __CUDACC__

}  // namespace b2f
