// instances of fft_real_kernel (r2c / c2r / r2r modes), float, N = 7 * 2^k
#include "fft_pow2_inst.cuh"
namespace b2f {
cudaError_t launch_real_mixed7_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    using T = float;
    if (strided) {
        B2F_REAL_STRIDED_MIXED7(B2F_INST_REAL_STRIDED)
    } else {
        B2F_REAL_CONTIG_MIXED7(B2F_INST_REAL_CONTIG)
    }
    return cudaErrorInvalidValue;
}
}  // namespace b2f
