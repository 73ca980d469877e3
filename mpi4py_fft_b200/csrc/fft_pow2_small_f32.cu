// generated shape: instances of fft_pow2_kernel for the "small" size group, float
#include "fft_pow2_inst.cuh"
namespace b2f {
cudaError_t launch_pow2_small_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    using T = float;
    B2F_POW2_TABLE_SMALL(B2F_INST_ROW)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
