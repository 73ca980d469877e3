// dispatch of the staged strided c2c kernels (fft_tma.cuh), float: variants 0..99 are
// TMA-loaded, 100.. the cp.async-loaded table; the instances live in fft_tma_{a,b,c}_f32.cu
// and fft_cpa_{a,b,c}_f32.cu (one translation unit per size group, so that they compile in parallel)
#include "fft_tma_inst.cuh"
namespace b2f {
cudaError_t launch_tma_a_f32(int n, int var, const TmaStep& st, cudaStream_t stream);
cudaError_t launch_tma_b_f32(int n, int var, const TmaStep& st, cudaStream_t stream);
cudaError_t launch_tma_c_f32(int n, int var, const TmaStep& st, cudaStream_t stream);
cudaError_t launch_cpa_a_f32(int n, int var, const TmaStep& st, cudaStream_t stream);
cudaError_t launch_cpa_b_f32(int n, int var, const TmaStep& st, cudaStream_t stream);
cudaError_t launch_cpa_c_f32(int n, int var, const TmaStep& st, cudaStream_t stream);

static int size_group(int n) { return (n <= 256 || n == 384) ? 0 : (n == 512 || n == 768) ? 1 : 2; }

cudaError_t launch_tma_f32(int n, int var, const TmaStep& st, cudaStream_t stream) {
    using T = float;
    const int g = size_group(n);
    if (var >= 100)
        return g == 0 ? launch_cpa_a_f32(n, var, st, stream) : g == 1 ? launch_cpa_b_f32(n, var, st, stream)
                                                                      : launch_cpa_c_f32(n, var, st, stream);
    if (!tma_can_serve<T>(st)) return cudaErrorInvalidValue;
    return g == 0 ? launch_tma_a_f32(n, var, st, stream) : g == 1 ? launch_tma_b_f32(n, var, st, stream)
                                                                  : launch_tma_c_f32(n, var, st, stream);
}
}  // namespace b2f
