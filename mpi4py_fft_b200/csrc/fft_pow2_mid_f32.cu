// generated shape: instances of fft_pow2_kernel for the "mid" size group, float
#include "fft_pow2_inst.cuh"
namespace b2f {
cudaError_t launch_pow2_mid_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    using T = float;
    B2F_POW2_TABLE_MID(B2F_INST_ROW)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
