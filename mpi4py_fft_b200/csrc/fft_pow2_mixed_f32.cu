// instances of fft_pow2_kernel for the lengths 3 * 2^k, float
#define B2F_GROUP_TRUNC 1
#include "fft_pow2_inst.cuh"
namespace b2f {
B2F_DEFINE_GROUP(launch_pow2_mixed_f32, float, B2F_CONTIG_MIXED, B2F_STRIDED_MIXED)
}  // namespace b2f
