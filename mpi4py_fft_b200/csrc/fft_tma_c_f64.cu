// instances of the TMA-loaded staged strided c2c kernels (fft_tma.cuh), double, size group C of fft_configs.h
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_tma_c_f64(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = double;
    B2F_TMA_TABLE_C(B2F_INST_TMA)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
