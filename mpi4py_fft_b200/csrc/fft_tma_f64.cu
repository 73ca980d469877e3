// instances of fft_tma_kernel (TMA-staged strided c2c), double
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_tma_f64(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = double;
    if (!tma_can_serve<T>(st)) return cudaErrorInvalidValue;
    B2F_TMA_TABLE(B2F_INST_TMA)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
