// instances of the rotating c2c kernels (fft_rot.cuh), double, size group B of fft_configs.h
#include "fft_rot_inst.cuh"
namespace b2f {
cudaError_t launch_rot_b_f64(int n, int var, const RotStep& st, cudaStream_t stream) {
    using T = double;
    B2F_ROT_TABLE_B(B2F_INST_ROT)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
