#pragma once
#include <cuda_runtime.h>
namespace b2f {
struct GenericParams {
    const void* in;
    void* out;
    const double* M;
    long long npencils;   // outer * inner
    long long inner;
    long long in_n, out_n;     // elements (of the array dtype) along the axis, in and out
    int in_c, out_c;           // reals per element: 1 (real) or 2 (complex)
    int rows, cols;            // = out_n*out_c, in_n*in_c
    int pb;                    // pencils per CTA (set by the launcher)
    double scale;
};
cudaError_t launch_generic(int precision, const GenericParams& prm, cudaStream_t st);
}
