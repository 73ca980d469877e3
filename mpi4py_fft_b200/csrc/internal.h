// internal.h -- shared declarations between the translation units of libb200fft.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "fft_pow2.cuh"
#include "chirpz.cuh"

namespace b2f {

void count_launch();
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);   // records message, returns B2F_ECUDA

int64_t option(const char* key, int64_t dflt);

// power-of-two c2c launchers, one translation unit per size group.
// returns cudaErrorInvalidValue if (n, var) is not in the group.
cudaError_t launch_pow2_small_f64(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mid_f64  (int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_large_f64(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_small_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mid_f32  (int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_large_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mixed_f64(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mixed_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mixed5_f64(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mixed5_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mixed7_f64(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_pow2_mixed7_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
inline cudaError_t launch_pow2_mixed57_f64(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    return n % 5 == 0 ? launch_pow2_mixed5_f64(n, var, strided, prm, outer, st) : launch_pow2_mixed7_f64(n, var, strided, prm, outer, st);
}
inline cudaError_t launch_pow2_mixed57_f32(int n, int var, bool strided, const FftParams& prm, long long outer, cudaStream_t st) {
    return n % 5 == 0 ? launch_pow2_mixed5_f32(n, var, strided, prm, outer, st) : launch_pow2_mixed7_f32(n, var, strided, prm, outer, st);
}

// real transforms of even length 2n through the n-point schedule (fft_real_*.cu); mode 1 = r2c, 2 = c2r
cudaError_t launch_real_f64(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_real_f32(int n, int mode, bool strided, const FftParams& prm, long long outer, cudaStream_t st);

// chirp-z kernels (chirpz_*.cu): m = convolution length
cudaError_t launch_chirp_f64(int m, bool strided, const ChirpParams& prm, long long outer, cudaStream_t st);
cudaError_t launch_chirp_f32(int m, bool strided, const ChirpParams& prm, long long outer, cudaStream_t st);

// TMA-staged strided c2c kernels (fft_tma_*.cu): one (outer, n, inner) step.
// cudaErrorInvalidValue = (n, var) not built or the layout cannot be described
// to TMA (16-byte base and row pitch); the caller then uses the register path.
struct TmaStep {
    const void* in;
    void* out;
    long long outer, n, inner;
    double scale;
    int swap;
    const PeerStore* peer;   // non-null: the last pass stores into the owners' arrays
    long long pitch;         // elements between consecutive points of a pencil (0: = inner)
    long long ostride;       // elements between consecutive outer indices (0: = n * pitch)
    int grid_cap;            // persistent grid limit (0: every SM) -- lets two stages share the GPU
};
cudaError_t launch_tma_f64(int n, int var, const TmaStep& st, cudaStream_t stream);
cudaError_t launch_tma_f32(int n, int var, const TmaStep& st, cudaStream_t stream);
int sm_count();   // SMs of the current device (cached)
// twiddle pass of a four-step split (fourstep.cu): data viewed (outer, n2, n1, inner), times W_(n1 n2)^(j1 k2)
cudaError_t launch_fourstep_twiddle(int precision, void* data, long long outer, long long n2, long long n1, long long inner,
                                    int backward, cudaStream_t st);

// rotating c2c kernels (fft_rot_*.cu): transform the contiguous axis of in[b][i][o][n]
// (pencil (i, o) of batch b starts at b*in_bstride + i*in_istride + o*in_ostride) and
// store out[b][o][k][i] = out + b*out_bstride + o*out_ostride + k*out_nstride + i.
// cudaErrorInvalidValue = (n, var) not built or pencils not 16-byte aligned.
struct RotStep {
    const void* in;
    void* out;
    long long batches, I, O;
    long long in_istride, in_ostride, in_bstride;
    long long out_ostride, out_nstride, out_bstride;
    double scale;
    int swap;
    int grid_cap;
};
cudaError_t launch_rot_a_f64(int n, int var, const RotStep& st, cudaStream_t stream);
cudaError_t launch_rot_b_f64(int n, int var, const RotStep& st, cudaStream_t stream);
cudaError_t launch_rot_a_f32(int n, int var, const RotStep& st, cudaStream_t stream);
cudaError_t launch_rot_b_f32(int n, int var, const RotStep& st, cudaStream_t stream);
inline cudaError_t launch_rot(int precision, int n, int var, const RotStep& st, cudaStream_t stream) {
    if (precision == 8) return n <= 512 ? launch_rot_a_f64(n, var, st, stream) : launch_rot_b_f64(n, var, st, stream);
    return n <= 512 ? launch_rot_a_f32(n, var, st, stream) : launch_rot_b_f32(n, var, st, stream);
}


}  // namespace b2f
struct b2f_plan_s;
namespace b2f {
// A launch over part of a one-axis stage, so that the stage feeding a redistribution
// and the stage consuming it can be pipelined chunk by chunk:
//   mode 1: inner indices [begin, begin+count) of every pencil row; with view_outer > 0
//           the block is re-viewed as view_outer rows view_ostride elements apart
//           (a range of the LAST array axis when other axes follow the transformed one)
//   mode 2: outer indices [begin, begin+count)
struct ChunkSpec {
    int mode;
    long long begin, count;
    long long view_outer, view_ostride;
    int grid_cap;
};
// plan execution with an optional fused peer store in the last step (capi.cu)
int run_plan(b2f_plan_s* pl, const void* d_in, void* d_out, double scale, cudaStream_t st, const PeerStore* peer_last,
             int (*before_last)(void*, cudaStream_t), void* ctx, const ChunkSpec* chunk = nullptr);
// last step of a plan: false unless it is a Stockham step (the only kind that can scatter)
bool plan_scatter_info(b2f_plan_s* pl, int* axis, long long* n, int* precision, const long long** out_shape, int* ndims);

// generic (any n, any kind) dense-matrix path, dft_generic.cu
struct GenericParams;
int build_matrix(int kind, long long n, std::vector<double>& M, long long& rows, long long& cols);

// matrix cache (capi.cu); device pointers live until process exit
const double* generic_matrix(int kind, long long n, int* rows, int* cols);

}  // namespace b2f
