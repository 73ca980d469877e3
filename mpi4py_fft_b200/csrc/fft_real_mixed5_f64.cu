// instances of fft_real_kernel (r2c / c2r / r2r modes), double, N = 5 * 2^k
#include "fft_pow2_inst.cuh"
namespace b2f {
cudaError_t launch_real_mixed5_f64(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    using T = double;
    if (strided) {
        B2F_REAL_STRIDED_MIXED5(B2F_INST_REAL_STRIDED)
    } else {
        B2F_REAL_CONTIG_MIXED5(B2F_INST_REAL_CONTIG)
    }
    return cudaErrorInvalidValue;
}
}  // namespace b2f
