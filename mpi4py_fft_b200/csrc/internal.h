// internal.h -- shared declarations between the translation units of libb200fft.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "fft_pow2.cuh"

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

// generic (any n, any kind) dense-matrix path, dft_generic.cu
struct GenericParams;
int build_matrix(int kind, long long n, std::vector<double>& M, long long& rows, long long& cols);

// twiddle / matrix caches (capi.cu); device pointers live until process exit
const void* twiddle_table(int n, int precision);          // n forward twiddles
const double* generic_matrix(int kind, long long n, int* rows, int* cols);

}  // namespace b2f
