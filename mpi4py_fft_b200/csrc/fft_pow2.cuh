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
    long long inner;         // STRIDED: extent of the contiguous inner index
    long long npencils;      // CONTIG: number of rows
    long long tiles_per_outer;
    double scale;
    int swap;
    PeerStore peer;          // peer.p > 0: the last pass stores into the owners' arrays (fused redistribution)
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
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, bool SWAP, bool PEER>
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
        gin = reinterpret_cast<const C*>(prm.in) + o * prm.in_ostride + i;
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
    TF::load_global(v, q, gin, in_ns, valid, SWAP);
    TF::template twiddle_dft<0>(v, q, tw);
    if constexpr (TF::NPASS > 1) {
        TF::template store_shared<0>(v, p, q, smem);
        __syncthreads();
        MidPasses<TF, 1>::run(v, p, q, smem, tw);
        TF::template load_shared<TF::NPASS - 1>(v, p, q, smem);
        TF::template twiddle_dft<TF::NPASS - 1>(v, q, tw);
    }
    if constexpr (PEER) {
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

// the same transform with the last pass storing into the owners' arrays (PeerStore)
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
__global__ void __launch_bounds__((N / E) * P, MINB) fft_pow2_peer_kernel(const FftParams prm) {
    if (prm.swap) fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, true, true>(prm);
    else fft_pow2_body<T, N, E, RAD, P, STRIDED, PS, false, true>(prm);
}

#endif  // __CUDACC__

}  // namespace b2f
