// fft_rot_inst.cuh -- launchers for the rotating kernels (fft_rot.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <cuda.h>
#include "fft_rot.cuh"
#include "fft_pow2_inst.cuh"
#include "fft_tma_inst.cuh"

namespace b2f {

template <class T, int N, int E, class RAD, int P, int PS, int STAGES, bool SPLIT, int MINB>
static cudaError_t launch_rot_one(const RotStep& st, cudaStream_t stream) {
    using TF = TileFFT<T, N, E, RAD, P, true, PS>;
    using EX = Exchange<TF, SPLIT>;
    auto kern = fft_rot_kernel<T, N, E, RAD, P, PS, STAGES, SPLIT, MINB>;
    constexpr size_t stage_bytes = sizeof(cplx<T>) * (size_t)(N + RotPad<T, P>::value) * P;
    constexpr int GROUPS = 1 << ((MINB >> 8) & 3);     // OPT bits 4-5 (fft_rot.cuh)
    constexpr int THREADS = TF::THREADS * GROUPS;
    constexpr size_t smem = GROUPS * (STAGES * stage_bytes + EX::bytes) + (((MINB >> 4) & 1) ? sizeof(cplx<T>) * (size_t)RAD::tw_total() : 0);
    static_assert(smem <= 227 * 1024, "tile does not fit shared memory");
    static int ctas_per_sm = 0;   // per instantiation
    if (!ctas_per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, THREADS, smem);
        if (e != cudaSuccess) return e;
        if (nb < 1) return cudaErrorInvalidConfiguration;
        ctas_per_sm = nb;
        if (option("verbose", 0))
            fprintf(stderr, "b2f: fft_rot_kernel n=%d E=%d P=%d stages=%d split=%d groups=%d threads=%d smem=%zu ctas/sm=%d\n", N, E, P,
                    STAGES, (int)SPLIT, GROUPS, THREADS, smem, nb);
    }
    // bulk copies need 16-byte aligned pencils on both sides
    if (((uintptr_t)st.in & 15) || ((uintptr_t)st.out & (sizeof(cplx<T>) - 1))) return cudaErrorInvalidValue;
    if ((st.in_istride * sizeof(cplx<T>)) % 16 || (st.in_ostride * sizeof(cplx<T>)) % 16 ||
        (st.in_bstride * sizeof(cplx<T>)) % 16)
        return cudaErrorInvalidValue;
    RotParams prm;
    prm.in = st.in;
    prm.out = st.out;
    prm.tw = pass_twiddles<T, RAD>();
    if (!prm.tw) return cudaErrorMemoryAllocation;
    prm.in_istride = st.in_istride;
    prm.in_ostride = st.in_ostride;
    prm.in_bstride = st.in_bstride;
    prm.out_ostride = st.out_ostride;
    prm.out_nstride = st.out_nstride;
    prm.out_bstride = st.out_bstride;
    prm.I = st.I;
    prm.O = st.O;
    prm.tiles_per_o = (st.I + P - 1) / P;
    prm.tiles_per_b = st.O * prm.tiles_per_o;
    prm.ntiles = st.batches * prm.tiles_per_b;
    prm.scale = st.scale;
    prm.swap = st.swap;
    if (prm.ntiles <= 0) return cudaSuccess;
    long long grid = (long long)sm_count() * ctas_per_sm;
    if (st.grid_cap > 0 && grid > (long long)st.grid_cap * ctas_per_sm) grid = (long long)st.grid_cap * ctas_per_sm;
    if (grid > (prm.ntiles + GROUPS - 1) / GROUPS) grid = (prm.ntiles + GROUPS - 1) / GROUPS;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if constexpr (((MINB >> 4) & 8) != 0) {
        // the output as a rank-4 tensor of T: [batch][o][k][2 * I], box = one row run of P elements x box_rows rows
        using OC = OutChunks<TF, EX>;
        TensorMapEncodeFn enc = tensor_map_encoder();
        if (!enc) return cudaErrorNotSupported;
        const size_t esz = sizeof(cplx<T>);
        if (((uintptr_t)st.out & 15) || (st.out_nstride * esz) % 16 || (st.out_ostride * esz) % 16 || (st.out_bstride * esz) % 16 ||
            2 * st.I >= (1LL << 32))
            return cudaErrorInvalidValue;
        const cuuint64_t dims[4] = {(cuuint64_t)(2 * st.I), (cuuint64_t)N, (cuuint64_t)st.O, (cuuint64_t)st.batches};
        const cuuint64_t strides[3] = {(cuuint64_t)(st.out_nstride * esz), (cuuint64_t)(st.out_ostride * esz),
                                       (cuuint64_t)(st.out_bstride * esz)};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * P), (cuuint32_t)OC::box_rows, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                         st.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(map, prm);
    count_launch();
    return cudaGetLastError();
}

#define B2F_INST_ROT(N, VAR, E, P, PS, STAGES, SPLIT, MINB, ...)                                              \
    if (n == N && var == VAR)                                                                                 \
        return launch_rot_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, PS, STAGES, SPLIT != 0, \
                              MINB>(st, stream);

}  // namespace b2f
