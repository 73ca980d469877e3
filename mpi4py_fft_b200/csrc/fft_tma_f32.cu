// instances of the staged strided c2c kernels (fft_tma.cuh), float:
// variants 0..99 are TMA-loaded, 100.. the cp.async-loaded table
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_tma_f32(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = float;
    if (var >= 100) {
        B2F_CPA_TABLE(B2F_INST_CPA)
        return cudaErrorInvalidValue;
    }
    if (!tma_can_serve<T>(st)) return cudaErrorInvalidValue;
    B2F_TMA_TABLE(B2F_INST_TMA)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
