// fourstep.cu -- the twiddle pass between the two transforms of a four-step split (lengths.h):
// element (o, k2, j1, i) of the (outer, n2, n1, inner) view is multiplied by W_n^(j1 k2).
#include <cuda_runtime.h>
#include "internal.h"

namespace b2f {

template <class T>
__global__ void __launch_bounds__(256) fourstep_twiddle_kernel(cplx<T>* __restrict__ data, long long total, long long n2,
                                                              long long n1, long long inner, long long n, int backward) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
        long long t = idx / inner;
        const long long j1 = t % n1;
        t /= n1;
        const long long k2 = t % n2;
        const cplx<T> w = fourstep_twiddle<T>((j1 * k2) % n, n, backward != 0);
        data[idx] = cmul(data[idx], w);
    }
}

cudaError_t launch_fourstep_twiddle(int precision, void* data, long long outer, long long n2, long long n1, long long inner,
                                    int backward, cudaStream_t st) {
    const long long total = outer * n2 * n1 * inner;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (precision == 8)
        fourstep_twiddle_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((cplx<double>*)data, total, n2, n1, inner, n1 * n2, backward);
    else
        fourstep_twiddle_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((cplx<float>*)data, total, n2, n1, inner, n1 * n2, backward);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2f
