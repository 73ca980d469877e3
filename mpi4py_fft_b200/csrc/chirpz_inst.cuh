// chirpz_inst.cuh -- launchers of the chirp-z kernels, one instance per
// convolution length M of the B2F_REAL_* tables (fft_configs.h; same row format).
#pragma once
#include "chirpz.cuh"
#include "fft_pow2_inst.cuh"

namespace b2f {

template <class T, int M, int E, class RAD, int P, bool STRIDED, int PS, int MINB>
static cudaError_t launch_chirp_one(const ChirpParams& prm_in, long long outer, cudaStream_t st) {
    using TF = TileFFT<T, M, E, RAD, P, STRIDED, PS>;
    auto kern = chirpz_kernel<T, M, E, RAD, P, STRIDED, PS, MINB>;
    constexpr size_t smem = sizeof(cplx<T>) * (size_t)TF::SI::tile_elems;
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        attr_done = true;
    }
    ChirpParams prm = prm_in;
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

#define B2F_INST_CHIRP_CONTIG(N, E, P, PS, MINB, ...) \
    if (m == N) return launch_chirp_one<T, N, E, Radices<__VA_ARGS__>, P, false, PS, MINB>(prm, outer, st);
#define B2F_INST_CHIRP_STRIDED(N, E, P, PS, MINB, ...) \
    if (m == N)                                        \
        return launch_chirp_one<T, N, E, Radices<__VA_ARGS__>, P * StridedScale<T>::value, true, PS, MINB>(prm, outer, st);

}  // namespace b2f
