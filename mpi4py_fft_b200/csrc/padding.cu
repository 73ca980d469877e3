// padding.cu -- truncation / zero padding of a spectrum along one axis: the
// dealiasing step of a padded transform (3/2-rule and friends).
//
// Replaces FFTBase._truncation_forward / _padding_backward of the reference
// (/root/reference/mpi4py_fft/libfft.py:263-311), including its symmetric
// treatment of the Nyquist mode:
//   complex spectrum, N kept of Np:  keep k <= N/2 and the last N/2 modes; for
//       even N the two halves meet at N/2 (forward: summed; backward: both
//       copies halved);
//   half (r2c) spectrum, N' kept:    keep k < N'; when N' is even the last kept
//       mode is made real and doubled (forward) / halved (backward).
// One thread per 16-byte (complex128) or 8-byte (complex64) output element,
// consecutive threads walk the contiguous inner index.
#include <cuda_runtime.h>
#include "b200fft.h"
#include "internal.h"

namespace b2f {

// mode 0: truncate (n_src = padded, n_dst = kept)   mode 1: pad (n_src = kept, n_dst = padded)
template <class T, int MODE, bool HALF>
__global__ void __launch_bounds__(256) pad_trunc_kernel(const cplx<T>* __restrict__ src, cplx<T>* __restrict__ dst,
                                                       long long outer, long long n_src, long long n_dst,
                                                       long long inner, T scale) {
    const long long total = outer * n_dst * inner;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
        const long long i = idx % inner;
        const long long t = idx / inner;
        const long long k = t % n_dst;
        const long long o = t / n_dst;
        const cplx<T>* s = src + o * n_src * inner + i;
        cplx<T> v = {(T)0, (T)0};
        if (MODE == 0) {
            const long long N = n_dst, Np = n_src;
            if (HALF) {
                v = s[k * inner];
                if (N % 2 == 0 && k == N - 1) {
                    v.x *= (T)2;
                    v.y = (T)0;
                }
            } else if (k <= N / 2) {
                v = s[k * inner];
                if (N % 2 == 0 && k == N / 2) {
                    const cplx<T> w = s[(Np - N / 2) * inner];
                    v.x += w.x;
                    v.y += w.y;
                }
            } else {
                v = s[(Np - N + k) * inner];
            }
        } else {
            const long long N = n_src, Np = n_dst;
            if (HALF) {
                if (k < N) {
                    v = s[k * inner];
                    if (N % 2 == 0 && k == N - 1) {
                        v.x *= (T)0.5;
                        v.y = (T)0;
                    }
                }
            } else {
                if (k <= N / 2) v = s[k * inner];
                if (k >= Np - N / 2) v = s[(N - (Np - k)) * inner];   // second assignment wins, as in the reference
                if (N % 2 == 0 && (k == N / 2 || k == Np - N / 2)) {
                    v.x *= (T)0.5;
                    v.y *= (T)0.5;
                }
            }
        }
        v.x *= scale;
        v.y *= scale;
        dst[idx] = v;
    }
}

template <class T>
static cudaError_t launch_pad(int mode, int half, const void* src, void* dst, long long outer, long long n_src,
                              long long n_dst, long long inner, double scale, cudaStream_t st) {
    const long long total = outer * n_dst * inner;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const cplx<T>* s = reinterpret_cast<const cplx<T>*>(src);
    cplx<T>* d = reinterpret_cast<cplx<T>*>(dst);
    if (mode == 0 && half) pad_trunc_kernel<T, 0, true><<<(unsigned)blocks, 256, 0, st>>>(s, d, outer, n_src, n_dst, inner, (T)scale);
    else if (mode == 0) pad_trunc_kernel<T, 0, false><<<(unsigned)blocks, 256, 0, st>>>(s, d, outer, n_src, n_dst, inner, (T)scale);
    else if (half) pad_trunc_kernel<T, 1, true><<<(unsigned)blocks, 256, 0, st>>>(s, d, outer, n_src, n_dst, inner, (T)scale);
    else pad_trunc_kernel<T, 1, false><<<(unsigned)blocks, 256, 0, st>>>(s, d, outer, n_src, n_dst, inner, (T)scale);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_pad_truncate(int mode, int half_spectrum, int precision, const void* d_src, void* d_dst,
                                int64_t outer, int64_t n_src, int64_t n_dst, int64_t inner, double scale,
                                void* stream) {
    if (!d_src || !d_dst || (mode != 0 && mode != 1) || (precision != 4 && precision != 8) || outer < 0 ||
        n_src < 1 || n_dst < 1 || inner < 1 || d_src == d_dst) {
        set_error("b2f_pad_truncate: bad arguments");
        return B2F_EINVAL;
    }
    if ((mode == 0 && n_dst > n_src) || (mode == 1 && n_src > n_dst)) {
        set_error("b2f_pad_truncate: the kept extent must not exceed the padded one");
        return B2F_EINVAL;
    }
    cudaError_t e = precision == 8
        ? launch_pad<double>(mode, half_spectrum, d_src, d_dst, outer, n_src, n_dst, inner, scale, (cudaStream_t)stream)
        : launch_pad<float>(mode, half_spectrum, d_src, d_dst, outer, n_src, n_dst, inner, scale, (cudaStream_t)stream);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "pad/truncate kernel");
}
