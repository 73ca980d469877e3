// instances of the rotating c2c kernels (fft_rot.cuh), float, size group A of fft_configs.h
#include "fft_rot_inst.cuh"
namespace b2f {
cudaError_t launch_rot_a_f32(int n, int var, const RotStep& st, cudaStream_t stream) {
    using T = float;
    B2F_ROT_TABLE_A(B2F_INST_ROT)
    return cudaErrorInvalidValue;
}
}  // namespace b2f
