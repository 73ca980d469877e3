// instances of the TMA-loaded staged strided c2c kernels (fft_tma.cuh), float, size group B of fft_configs.h
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_tma_b_f32(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = float;
    B2F_TMA_TABLE_B(B2F_INST_TMA)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
