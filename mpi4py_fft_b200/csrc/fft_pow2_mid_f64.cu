// instances of fft_pow2_kernel for the "mid" size group, double
#define B2F_GROUP_TRUNC 1
#include "fft_pow2_inst.cuh"
namespace b2f {
B2F_DEFINE_GROUP(launch_pow2_mid_f64, double, B2F_CONTIG_MID, B2F_STRIDED_MID)
}  // namespace b2f
