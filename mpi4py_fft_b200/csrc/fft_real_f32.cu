// instances of fft_real_kernel (r2c / c2r through the N-point complex schedule), float
#include "fft_pow2_inst.cuh"
namespace b2f {
// n = complex length N (real length 2N); mode 1 = r2c, 2 = c2r
cudaError_t launch_real_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    using T = float;
    if (strided) {
        B2F_REAL_STRIDED(B2F_INST_REAL_STRIDED)
    } else {
        B2F_REAL_CONTIG(B2F_INST_REAL_CONTIG)
    }
    return cudaErrorInvalidValue;
}
}  // namespace b2f
