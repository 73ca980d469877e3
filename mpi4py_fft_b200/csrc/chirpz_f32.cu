// instances of chirpz_kernel (any length / any kind through two M-point FFTs), float
#include "chirpz_inst.cuh"
namespace b2f {
cudaError_t launch_chirp_f32(int m, bool strided, const ChirpParams& prm, long long outer, cudaStream_t st) {
    using T = float;
    if (strided) {
        B2F_REAL_STRIDED_POW2(B2F_INST_CHIRP_STRIDED)
    } else {
        B2F_REAL_CONTIG_POW2(B2F_INST_CHIRP_CONTIG)
    }
    return cudaErrorInvalidValue;
}
}  // namespace b2f
