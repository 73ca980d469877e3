// instances of fft_pow2_kernel for the lengths 5 * 2^k, double
#include "fft_pow2_inst.cuh"
namespace b2f {
B2F_DEFINE_GROUP(launch_pow2_mixed5_f64, double, B2F_CONTIG_MIXED5, B2F_STRIDED_MIXED5)
}  // namespace b2f
