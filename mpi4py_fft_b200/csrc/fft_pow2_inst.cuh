// fft_pow2_inst.cuh -- turns rows of fft_configs.h into launchers.
// Each fft_pow2_<group>_<prec>.cu includes this and expands the CONTIG and
// STRIDED tables of its size group, so that the (large, fully unrolled) kernels
// compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <utility>
#include <vector>
#include "fft_pow2.cuh"
#include "fft_configs.h"
#include "internal.h"

namespace b2f {

// per-pass twiddle tables of one radix schedule, uploaded once per device
template <class T, class RAD>
static const void* pass_twiddles() {
    static std::mutex mu;
    static const void* cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    const void*& slot = cache[dev & 63];
    if (!slot) {
        std::vector<cplx<T>> h((size_t)RAD::tw_total());
        build_pass_twiddles<T, RAD>(h.data());
        void* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(cplx<T>)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        slot = d;
    }
    return slot;
}

// translation units that define B2F_GROUP_TRUNC also build the dealiasing flavour of their
// kernels (TruncMap): the 3 * 2^k lengths, which is what a 3/2-rule padded solver transforms
#ifndef B2F_GROUP_TRUNC
#define B2F_GROUP_TRUNC 0
#endif

template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
static cudaError_t launch_one(const FftParams& prm_in, long long outer, cudaStream_t st) {
    using TF = TileFFT<T, N, E, RAD, P, STRIDED, PS>;
    const bool peer = prm_in.peer.p > 0;
    const bool trunc = prm_in.trunc.n > 0;
    auto kern = peer ? fft_pow2_peer_kernel<T, N, E, RAD, P, STRIDED, PS, MINB>
                     : fft_pow2_kernel<T, N, E, RAD, P, STRIDED, PS, MINB>;
    size_t smem = TF::NPASS > 1 ? sizeof(cplx<T>) * (size_t)TF::SI::tile_elems : 0;
    if (trunc) {
#if B2F_GROUP_TRUNC
        if (peer && !prm_in.swap) return cudaErrorInvalidValue;   // a truncating store cannot scatter
        kern = peer ? fft_pow2_trunc_peer_kernel<T, N, E, RAD, P, STRIDED, PS, MINB>
                    : fft_pow2_trunc_kernel<T, N, E, RAD, P, STRIDED, PS, MINB>;
        smem += sizeof(cplx<T>) * (size_t)P;
#else
        return cudaErrorInvalidValue;
#endif
    }
    static bool attr_done[4] = {false, false, false, false};   // per instantiation and flavour
    if (!attr_done[peer + 2 * trunc]) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        attr_done[peer + 2 * trunc] = true;
    }
    FftParams prm = prm_in;
    prm.tw = pass_twiddles<T, RAD>();
    if (!prm.tw) return cudaErrorMemoryAllocation;
    long long grid;
    if (STRIDED) {
        prm.tiles_per_outer = (prm.inner + P - 1) / P;
        grid = outer * prm.tiles_per_outer;
    } else {
        grid = (prm.npencils + P - 1) / P;
    }
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    kern<<<(unsigned)grid, TF::THREADS, smem, st>>>(prm);
    count_launch();
    return cudaGetLastError();
}

// split / merge twiddles of the real transforms, one table per (T, N) and device
template <class T>
static const void* real_twiddles(int N) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, const void*> cache;   // (device, N): lengths of different families share a log2
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    const void*& slot = cache[std::make_pair(dev, N)];
    if (!slot) {
        std::vector<cplx<T>> h((size_t)N);
        build_real_twiddles<T>(h.data(), N);
        void* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(cplx<T>)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        slot = d;
    }
    return slot;
}

// quarter-wave twiddles of the r2r kinds II / III, one table per (T, N) and device
template <class T>
static const void* quarter_twiddles(int N) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, const void*> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    const void*& slot = cache[std::make_pair(dev, N)];
    if (!slot) {
        std::vector<cplx<T>> h((size_t)N + 1);
        build_quarter_twiddles<T>(h.data(), N);
        void* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(cplx<T>)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        slot = d;
    }
    return slot;
}

// input- and output-side twiddles of the r2r kinds IV (build_dct4_twiddles), one table per (T, N) and device
template <class T>
static const void* dct4_twiddles(int N) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, const void*> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    const void*& slot = cache[std::make_pair(dev, N)];
    if (!slot) {
        std::vector<cplx<T>> h((size_t)2 * N);
        build_dct4_twiddles<T>(h.data(), N);
        void* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(cplx<T>)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        slot = d;
    }
    return slot;
}

template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB, int MODE>
static cudaError_t launch_real_one(const FftParams& prm_in, long long outer, cudaStream_t st) {
    using TF = TileFFT<T, N, E, RAD, P, STRIDED, PS>;
    auto kern = fft_real_kernel<T, N, E, RAD, P, STRIDED, PS, MINB, MODE>;
    constexpr size_t smem = sizeof(cplx<T>) * ((size_t)TF::SI::tile_elems + P);
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        attr_done = true;
    }
    FftParams prm = prm_in;
    prm.tw = pass_twiddles<T, RAD>();
    prm.rtw = real_twiddles<T>(N);
    if (!prm.tw || !prm.rtw) return cudaErrorMemoryAllocation;
    if (MODE == 3 || MODE == 4) {
        prm.qtw = quarter_twiddles<T>(N);
        if (!prm.qtw) return cudaErrorMemoryAllocation;
    }
    if (MODE == 6) {
        prm.qtw = dct4_twiddles<T>(N);
        if (!prm.qtw) return cudaErrorMemoryAllocation;
    }
    long long grid;
    if (STRIDED) {
        prm.tiles_per_outer = (prm.inner + P - 1) / P;
        grid = outer * prm.tiles_per_outer;
    } else {
        grid = (prm.npencils + P - 1) / P;
    }
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    kern<<<(unsigned)grid, TF::THREADS, smem, st>>>(prm);
    count_launch();
    return cudaGetLastError();
}

// mode 1 = r2c, 2 = c2r, 3 = r2r kinds II (REDFT10 / RODFT10), 4 = kinds III (REDFT01 / RODFT01), 5 = kinds I (REDFT00 / RODFT00), 6 = kinds IV (REDFT11 / RODFT11)
#define B2F_INST_REAL_CONTIG(N, E, P, PS, MINB, ...)                                                       \
    if (n == N) {                                                                                          \
        if (mode == 1) return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB, 1>(prm, outer, st); \
        if (mode == 3) return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB, 3>(prm, outer, st); \
        if (mode == 4) return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB, 4>(prm, outer, st); \
        if (mode == 5) return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB, 5>(prm, outer, st); \
        if (mode == 6) return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB, 6>(prm, outer, st); \
        return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB, 2>(prm, outer, st);      \
    }
#define B2F_INST_REAL_STRIDED(N, E, P, PS, MINB, ...)                                                      \
    if (n == N) {                                                                                          \
        if (mode == 1)                                                                                     \
            return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB, 1>(prm, outer, st); \
        if (mode == 3)                                                                                     \
            return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB, 3>(prm, outer, st); \
        if (mode == 4)                                                                                     \
            return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB, 4>(prm, outer, st); \
        if (mode == 5)                                                                                     \
            return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB, 5>(prm, outer, st); \
        if (mode == 6)                                                                                     \
            return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB, 6>(prm, outer, st); \
        return launch_real_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB, 2>(prm, outer, st);     \
    }

// strided tiles are sized in bytes: a float tile takes twice the pencils of a
// double tile, so that a row of the tile is the same contiguous run in HBM
template <class T> struct StridedScale { static constexpr int value = (int)(sizeof(double) / sizeof(T)); };

#define B2F_INST_CONTIG(N, VAR, E, P, PS, MINB, ...)                                        \
    if (n == N && var == VAR)                                                               \
        return launch_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB>(prm, outer, st);
#define B2F_INST_STRIDED(N, VAR, E, P, PS, MINB, ...)                                       \
    if (n == N && var == VAR)                                                               \
        return launch_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB>(prm, outer, st);

#define B2F_DEFINE_GROUP(FN, T_, CONTIG_TABLE, STRIDED_TABLE)                                               \
    cudaError_t FN(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st) { \
        using T = T_;                                                                                       \
        if (strided) {                                                                                      \
            STRIDED_TABLE(B2F_INST_STRIDED)                                                                 \
        } else {                                                                                            \
            CONTIG_TABLE(B2F_INST_CONTIG)                                                                   \
        }                                                                                                   \
        return cudaErrorInvalidValue;                                                                       \
    }

}  // namespace b2f
