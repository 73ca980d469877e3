// instances of the cp.async-loaded staged strided c2c kernels (fft_tma.cuh), float, size group C of fft_configs.h
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_cpa_c_f32(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = float;
    B2F_CPA_TABLE_C(B2F_INST_CPA)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
