// chirpz.cuh -- any-length, any-kind batched 1-D transform as a chirp-z (Bluestein)
// convolution on the power-of-two Stockham tile machinery (sm_100a).
//
// One CTA owns P pencils.  Per pencil (tables from chirpz_host.h):
//   load x[j] (complex | real | hermitian half spectrum) * pre[j], zero padded to M
//   M-point forward FFT                      (TileFFT passes, shared-memory exchanges)
//   * filter spectrum (includes 1/M), re/im swap, natural order -> shared memory
//   M-point FFT again (= backward, by the swap trick)
//   y[k] = swap(result)[k] * post[k] * scale, complex or real part, k < n_out
// The pencil is read once and written once; O(M log M) work replaces the O(n^2)
// dense map for every length the power-of-two kernels do not serve, and for all
// eight r2r kinds.  Replaces the FFTW plans behind
// /root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:52-75 for those cases.
#pragma once
#include "fft_core.cuh"

namespace b2f {

struct ChirpParams {
    const void* in;
    void* out;
    const void* tw;        // per-pass twiddles of the M-point schedule
    const void* pre;       // n_in complex
    const void* filt;      // M complex
    const void* post;      // n_out complex
    long long in_ostride, out_ostride;   // elements of the respective array between outer indices
    long long in_nstride, out_nstride;   // elements between consecutive points of a pencil
    long long inner, npencils, tiles_per_outer;
    int n_in, n_out;       // logical points read / written per pencil
    int in_mode;           // 0 complex, 1 real, 2 hermitian (n_in/2+1 stored)
    int out_real;
    double scale;
};

// one input point of a pencil, by load mode
template <class T>
B2F_HD cplx<T> chirp_load(const void* base, long long ns, int j, int n_in, int in_mode) {
    cplx<T> a;
    if (in_mode == 0) {
        a = reinterpret_cast<const cplx<T>*>(base)[(long long)j * ns];
    } else if (in_mode == 1) {
        a.x = reinterpret_cast<const T*>(base)[(long long)j * ns];
        a.y = (T)0;
    } else {
        const int half = n_in / 2;
        const int k = j <= half ? j : n_in - j;
        a = reinterpret_cast<const cplx<T>*>(base)[(long long)k * ns];
        if (j > half) a.y = -a.y;
        if (k == 0 || 2 * k == n_in) a.y = (T)0;   // FFTW ignores these imaginary parts
    }
    return a;
}

// phase 1: load, pre-chirp, zero pad (pass-0 register pattern)
template <class TF>
B2F_HD void chirp_phase_load(typename TF::C* v, int q, const void* base, long long ns, bool valid,
                             const ChirpParams& prm) {
    using C = typename TF::C;
    using T = typename TF::Real;
    constexpr int R = TF::RADS::get(0);
    const C* __restrict__ pre = reinterpret_cast<const C*>(prm.pre);
#pragma unroll
    for (int b = 0; b < TF::EPT / R; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = q + b * TF::TP + r * (TF::LEN / R);
            C a = {(T)0, (T)0};
            if (valid && n < prm.n_in) a = cmul(chirp_load<T>(base, ns, n, prm.n_in, prm.in_mode), pre[n]);
            v[b * R + r] = a;
        }
}

// phase 2: spectrum of the padded pencil (last-pass register pattern) * filter,
// swapped, to shared memory in natural order
template <class TF>
B2F_HD void chirp_phase_filter(const typename TF::C* v, int p, int q, typename TF::C* smem, const ChirpParams& prm) {
    using C = typename TF::C;
    constexpr int R = TF::RADS::get(TF::NPASS - 1);
    const C* __restrict__ fh = reinterpret_cast<const C*>(prm.filt);
#pragma unroll
    for (int b = 0; b < TF::EPT / R; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = q + b * TF::TP + r * (TF::LEN / R);
            const C a = cmul(v[b * R + r], fh[n]);
            smem[TF::SI::at(p, n)] = {a.y, a.x};
        }
}

// phase 3: pass-0 register pattern of the second transform
template <class TF>
B2F_HD void chirp_phase_reload(typename TF::C* v, int p, int q, const typename TF::C* smem) {
    constexpr int R = TF::RADS::get(0);
#pragma unroll
    for (int b = 0; b < TF::EPT / R; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) v[b * R + r] = smem[TF::SI::at(p, q + b * TF::TP + r * (TF::LEN / R))];
}

// phase 4: post-chirp, scale, store the first n_out points
template <class TF>
B2F_HD void chirp_phase_store(const typename TF::C* v, int q, void* base, long long ns, bool valid,
                              const ChirpParams& prm) {
    using C = typename TF::C;
    using T = typename TF::Real;
    constexpr int R = TF::RADS::get(TF::NPASS - 1);
    const C* __restrict__ post = reinterpret_cast<const C*>(prm.post);
    const T sc = (T)prm.scale;
    if (!valid) return;
#pragma unroll
    for (int b = 0; b < TF::EPT / R; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = q + b * TF::TP + r * (TF::LEN / R);
            if (n < prm.n_out) {
                const C y = {v[b * R + r].y, v[b * R + r].x};   // swap back: the backward transform
                const C z = cmul(y, post[n]);
                if (prm.out_real) reinterpret_cast<T*>(base)[(long long)n * ns] = z.x * sc;
                else reinterpret_cast<C*>(base)[(long long)n * ns] = {z.x * sc, z.y * sc};
            }
        }
}

#if defined(__CUDACC__)

template <class TF, int S>
struct ChirpMid {
    using C = typename TF::C;
    static __device__ __forceinline__ void run(C* v, int p, int q, C* smem, const C* __restrict__ tw) {
        if constexpr (S < TF::NPASS - 1) {
            TF::template load_shared<S>(v, p, q, smem);
            TF::template twiddle_dft<S>(v, q, tw);
            __syncthreads();
            TF::template store_shared<S>(v, p, q, smem);
            __syncthreads();
            ChirpMid<TF, S + 1>::run(v, p, q, smem, tw);
        }
    }
};

template <class TF>
__device__ __forceinline__ void chirp_fft(typename TF::C* v, int p, int q, typename TF::C* smem,
                                          const typename TF::C* __restrict__ tw) {
    TF::template twiddle_dft<0>(v, q, tw);
    if constexpr (TF::NPASS > 1) {
        TF::template store_shared<0>(v, p, q, smem);
        __syncthreads();
        ChirpMid<TF, 1>::run(v, p, q, smem, tw);
        TF::template load_shared<TF::NPASS - 1>(v, p, q, smem);
        TF::template twiddle_dft<TF::NPASS - 1>(v, q, tw);
    }
}

template <class T, int M, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
__global__ void __launch_bounds__((M / E) * P, MINB) chirpz_kernel(const ChirpParams prm) {
    using TF = TileFFT<T, M, E, RAD, P, STRIDED, PS>;
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
    const int in_size = prm.in_mode == 1 ? (int)sizeof(T) : (int)sizeof(C);
    const int out_size = prm.out_real ? (int)sizeof(T) : (int)sizeof(C);
    const char* gin = reinterpret_cast<const char*>(prm.in) + (o * prm.in_ostride + i) * in_size;
    char* gout = reinterpret_cast<char*>(prm.out) + (o * prm.out_ostride + i) * out_size;
    const C* __restrict__ tw = reinterpret_cast<const C*>(prm.tw);

    C v[E];
    chirp_phase_load<TF>(v, q, gin, prm.in_nstride, valid, prm);
    chirp_fft<TF>(v, p, q, smem, tw);
    if constexpr (TF::NPASS > 1) __syncthreads();   // the last pass has read the tile
    chirp_phase_filter<TF>(v, p, q, smem, prm);
    __syncthreads();
    chirp_phase_reload<TF>(v, p, q, smem);
    __syncthreads();
    chirp_fft<TF>(v, p, q, smem, tw);
    chirp_phase_store<TF>(v, q, gout, prm.out_nstride, valid, prm);
}

#endif  // __CUDACC__

}  // namespace b2f
