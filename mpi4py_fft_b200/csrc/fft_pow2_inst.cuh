// fft_pow2_inst.cuh -- turns rows of fft_configs.h into launchers.
// Each fft_pow2_<group>.cu includes this with B2F_INST_TABLE / B2F_INST_NAME set
// so that the (large, fully unrolled) kernels compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include "fft_pow2.cuh"
#include "fft_configs.h"
#include "internal.h"

namespace b2f {

template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
static cudaError_t launch_one(const FftParams& prm_in, long long outer, cudaStream_t st) {
    using TF = TileFFT<T, N, E, RAD, P, STRIDED, PS>;
    auto kern = fft_pow2_kernel<T, N, E, RAD, P, STRIDED, PS, MINB>;
    constexpr size_t smem = TF::NPASS > 1 ? sizeof(cplx<T>) * (size_t)TF::SI::tile_elems : 0;
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        attr_done = true;
    }
    FftParams prm = prm_in;
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

#define B2F_INST_ROW(N, VAR, E, PC, PSC, PST, PSS, MINB, ...)                                        \
    if (n == N && var == VAR) {                                                                      \
        using RAD = Radices<__VA_ARGS__>;                                                            \
        return strided ? launch_one<T, N, E, RAD, PST, true, PSS, MINB>(prm, outer, st)              \
                       : launch_one<T, N, E, RAD, PC, false, PSC, MINB>(prm, outer, st);             \
    }

}  // namespace b2f
