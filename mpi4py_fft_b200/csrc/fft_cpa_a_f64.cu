// instances of the cp.async-loaded staged strided c2c kernels (fft_tma.cuh), double, size group A of fft_configs.h
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_cpa_a_f64(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = double;
    B2F_CPA_TABLE_A(B2F_INST_CPA)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
