// instances of fft_real_kernel (r2c / c2r through the N-point complex schedule), float, N = 2^k
#include "fft_pow2_inst.cuh"
namespace b2f {
cudaError_t launch_real_mixed_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_real_mixed5_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_real_mixed7_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
// n = complex length N (real length 2N); mode 1 = r2c, 2 = c2r
cudaError_t launch_real_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    using T = float;
    if (n % 3 == 0) return launch_real_mixed_f32(n, mode, strided, prm, outer, st);
    if (n % 5 == 0) return launch_real_mixed5_f32(n, mode, strided, prm, outer, st);
    if (n % 7 == 0) return launch_real_mixed7_f32(n, mode, strided, prm, outer, st);
    if (strided) {
        B2F_REAL_STRIDED_POW2(B2F_INST_REAL_STRIDED)
    } else {
        B2F_REAL_CONTIG_POW2(B2F_INST_REAL_CONTIG)
    }
    return cudaErrorInvalidValue;
}
}  // namespace b2f
