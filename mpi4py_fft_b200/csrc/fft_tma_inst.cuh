// fft_tma_inst.cuh -- launchers for the TMA-staged strided kernels (fft_tma.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <vector>
#include "fft_tma.cuh"
#include "fft_pow2_inst.cuh"

namespace b2f {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder();   // capi.cu

// the layout constraints of a TMA descriptor (16-byte base and strides)
template <class T>
static bool tma_can_serve(const TmaStep& st) {
    const size_t esz = 2 * sizeof(T);
    if (st.pitch || st.ostride) return false;          // partial launches go through the cp.async flavour
    if (((uintptr_t)st.in & 15) || ((uintptr_t)st.out & 15)) return false;
    if ((st.inner * esz) % 16) return false;
    if (2 * st.inner >= (1LL << 32) || st.outer >= (1LL << 32)) return false;
    if ((double)st.n * st.inner * esz >= (double)(1ULL << 40)) return false;
    return true;
}

static inline void fill_params(TmaParams& prm, const TmaStep& st, int P) {
    const long long pitch = st.pitch ? st.pitch : st.inner;
    const long long ostride = st.ostride ? st.ostride : st.n * pitch;
    prm.in = st.in;
    prm.in_ostride = ostride;
    prm.in_nstride = pitch;
    prm.out = st.out;
    prm.out_ostride = ostride;
    prm.out_nstride = pitch;
    prm.inner = st.inner;
    prm.tiles_per_outer = (st.inner + P - 1) / P;
    prm.ntiles = st.outer * prm.tiles_per_outer;
    prm.scale = st.scale;
    prm.swap = st.swap;
    if (st.peer) prm.peer = *st.peer;
    else prm.peer.p = 0;
}

// cp.async-staged flavour: no descriptor, any layout whose elements are naturally aligned
template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MINB>
static cudaError_t launch_cpa_one(const TmaStep& st, cudaStream_t stream) {
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    const bool peer = st.peer != nullptr && st.peer->p > 0;
    auto kern = peer ? fft_cpa_peer_kernel<T, N, E, RAD, P, PS, STAGES, SPLIT, MINB>
                     : fft_cpa_kernel<T, N, E, RAD, P, PS, STAGES, SPLIT, MINB>;
    constexpr size_t tile_bytes = sizeof(cplx<T>) * (size_t)N * P;
    constexpr size_t smem = STAGES * tile_bytes + EX::bytes + (((MINB >> 4) & 1) ? sizeof(cplx<T>) * (size_t)RAD::tw_total() : 0);
    static_assert(smem <= 227 * 1024, "tile does not fit shared memory");
    static int ctas_cache[2] = {0, 0};   // per instantiation and flavour
    int& ctas_per_sm = ctas_cache[peer];
    if (!ctas_per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, TF::THREADS, smem);
        if (e != cudaSuccess) return e;
        if (nb < 1) return cudaErrorInvalidConfiguration;
        ctas_per_sm = nb;
    }
    if (((uintptr_t)st.in | (uintptr_t)st.out) & (sizeof(cplx<T>) - 1)) return cudaErrorInvalidValue;
    TmaParams prm;
    fill_params(prm, st, P);
    prm.tw = pass_twiddles<T, RAD>();
    if (!prm.tw) return cudaErrorMemoryAllocation;
    if (prm.ntiles <= 0) return cudaSuccess;
    long long grid = (long long)sm_count() * ctas_per_sm;
    if (st.grid_cap > 0 && grid > (long long)st.grid_cap * ctas_per_sm) grid = (long long)st.grid_cap * ctas_per_sm;
    if (grid > prm.ntiles) grid = prm.ntiles;
    if constexpr (((MINB >> 4) & 8) != 0) {
        // TMA-store flavour: whole-block launches without a fused peer store only
        if (peer || st.pitch || st.ostride) return cudaErrorInvalidValue;
        using OC = OutChunks<TF, EX>;
        auto kts = fft_cpa_ts_kernel<T, N, E, RAD, P, PS, STAGES, SPLIT, MINB>;
        static bool ts_attr = false;
        if (!ts_attr) {
            cudaError_t e = cudaFuncSetAttribute(kts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            ts_attr = true;
        }
        TensorMapEncodeFn enc = tensor_map_encoder();
        if (!enc) return cudaErrorNotSupported;
        const size_t esz = sizeof(cplx<T>);
        if (((uintptr_t)st.out & 15) || (st.inner * esz) % 16 || 2 * st.inner >= (1LL << 32)) return cudaErrorInvalidValue;
        CUtensorMap map;
        const cuuint64_t dims[4] = {(cuuint64_t)(2 * st.inner), (cuuint64_t)N, (cuuint64_t)st.outer, 1};
        const cuuint64_t strides[3] = {(cuuint64_t)(st.inner * esz), (cuuint64_t)(st.inner * esz) * N,
                                       (cuuint64_t)(st.inner * esz) * N * (cuuint64_t)st.outer};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * P), (cuuint32_t)OC::box_rows, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                         st.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
        kts<<<(unsigned)grid, TF::THREADS, smem, stream>>>(map, prm);
        count_launch();
        return cudaGetLastError();
    }
    kern<<<(unsigned)grid, TF::THREADS, smem, stream>>>(prm);
    count_launch();
    return cudaGetLastError();
}

template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MINB>
static cudaError_t launch_tma_one(const TmaStep& st, cudaStream_t stream) {
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    const bool peer = st.peer != nullptr && st.peer->p > 0;
    auto kern = peer ? fft_tma_peer_kernel<T, N, E, RAD, P, PS, STAGES, SPLIT, MINB>
                     : fft_tma_kernel<T, N, E, RAD, P, PS, STAGES, SPLIT, MINB>;
    constexpr size_t tile_bytes = sizeof(cplx<T>) * (size_t)N * P;
    constexpr size_t smem = STAGES * tile_bytes + EX::bytes + (((MINB >> 4) & 1) ? sizeof(cplx<T>) * (size_t)RAD::tw_total() : 0);
    static_assert(smem <= 227 * 1024, "tile does not fit shared memory");
    static int ctas_cache[2] = {0, 0};   // per instantiation and flavour
    int& ctas_per_sm = ctas_cache[peer];
    if (!ctas_per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, TF::THREADS, smem);
        if (e != cudaSuccess) return e;
        if (nb < 1) return cudaErrorInvalidConfiguration;
        ctas_per_sm = nb;
    }
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return cudaErrorNotSupported;
    constexpr int BR = N > 256 ? 256 : N;
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)(2 * st.inner), (cuuint64_t)N, (cuuint64_t)st.outer};
    const cuuint64_t strides[2] = {(cuuint64_t)(st.inner * sizeof(cplx<T>)), (cuuint64_t)(st.inner * sizeof(cplx<T>)) * N};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * P), (cuuint32_t)BR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void*>(st.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

    TmaParams prm;
    fill_params(prm, st, P);
    prm.tw = pass_twiddles<T, RAD>();
    if (!prm.tw) return cudaErrorMemoryAllocation;
    if (prm.ntiles <= 0) return cudaSuccess;
    long long grid = (long long)sm_count() * ctas_per_sm;
    if (st.grid_cap > 0 && grid > (long long)st.grid_cap * ctas_per_sm) grid = (long long)st.grid_cap * ctas_per_sm;
    if (grid > prm.ntiles) grid = prm.ntiles;
    kern<<<(unsigned)grid, TF::THREADS, smem, stream>>>(map, prm);
    count_launch();
    return cudaGetLastError();
}

#define B2F_INST_TMA(N, VAR, E, P, PS, STAGES, SPLIT, MINB, ...)                                              \
    if (n == N && var == VAR)                                                                                 \
        return launch_tma_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, PS, STAGES, SPLIT != 0, \
                              MINB>(st, stream);

#define B2F_INST_CPA(N, VAR, E, P, PS, STAGES, SPLIT, MINB, ...)                                              \
    if (n == N && var == 100 + VAR)                                                                           \
        return launch_cpa_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, PS, STAGES, SPLIT != 0, \
                              MINB>(st, stream);

}  // namespace b2f
