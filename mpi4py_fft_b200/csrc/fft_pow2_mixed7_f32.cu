// instances of fft_pow2_kernel for the lengths 7 * 2^k, float
#include "fft_pow2_inst.cuh"
namespace b2f {
B2F_DEFINE_GROUP(launch_pow2_mixed7_f32, float, B2F_CONTIG_MIXED7, B2F_STRIDED_MIXED7)
}  // namespace b2f
