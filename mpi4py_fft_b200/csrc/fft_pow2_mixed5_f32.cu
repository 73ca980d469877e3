// instances of fft_pow2_kernel for the lengths 5 * 2^k, float
#include "fft_pow2_inst.cuh"
namespace b2f {
B2F_DEFINE_GROUP(launch_pow2_mixed5_f32, float, B2F_CONTIG_MIXED5, B2F_STRIDED_MIXED5)
}  // namespace b2f
