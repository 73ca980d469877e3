// instances of fft_pow2_kernel for the "large" size group, double
#define B2F_GROUP_TRUNC 1
#include "fft_pow2_inst.cuh"
namespace b2f {
B2F_DEFINE_GROUP(launch_pow2_large_f64, double, B2F_CONTIG_LARGE, B2F_STRIDED_LARGE)
}  // namespace b2f
