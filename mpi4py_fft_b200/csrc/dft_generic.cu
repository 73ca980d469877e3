// dft_generic.cu -- batched 1-D transform of ANY length and ANY kind as a dense
// real-linear map, out = M * in, applied along one axis of an (outer, n, inner)
// block.  This is the coverage path: it serves every transform kind FFTW's guru
// interface offers the reference (c2c, r2c, c2r, the eight r2r DCT/DST kinds,
// /root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:49-76) for lengths that the
// Stockham kernels do not handle (non powers of two: the reference's own tests
// use 5..13).  O(n^2) per pencil, so it is meant for short axes; the plan layer
// refuses n above B2F_GENERIC_MAX_N.
//
// Complex data is treated as interleaved reals: a pencil of n complex points is
// a vector of 2n reals, and the matrix is (out reals) x (in reals), computed on
// the host in long double from the FFTW definitions and stored as double.
#include <cuda_runtime.h>
#include <math.h>
#include <vector>
#include "internal.h"
#include "dft_generic.h"
#include "b200fft.h"

namespace b2f {

// ---- host: the matrices ---------------------------------------------------
static long double angle(long long num, long long den) {
    // 2*pi*num/den with the integer part of num/den removed first
    num %= den;
    return 2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
}

// logical length n (for r2c/c2r: the real length); returns rows x cols doubles
int build_matrix(int kind, long long n, std::vector<double>& M, long long& rows, long long& cols) {
    const long double PI = 3.14159265358979323846264338327950288L;
    if (kind == B2F_FORWARD || kind == B2F_BACKWARD) {
        rows = cols = 2 * n;
        M.assign((size_t)(rows * cols), 0.0);
        const long double sg = (kind == B2F_FORWARD) ? -1.0L : 1.0L;
        for (long long k = 0; k < n; ++k)
            for (long long j = 0; j < n; ++j) {
                const long double a = angle(k * j, n);
                const double c = (double)cosl(a), s = (double)(sg * sinl(a));
                // (xr + i xi)(c + i s)
                M[(size_t)((2 * k) * cols + 2 * j)] = c;
                M[(size_t)((2 * k) * cols + 2 * j + 1)] = -s;
                M[(size_t)((2 * k + 1) * cols + 2 * j)] = s;
                M[(size_t)((2 * k + 1) * cols + 2 * j + 1)] = c;
            }
        return 0;
    }
    if (kind == B2F_R2C) {
        const long long h = n / 2 + 1;
        rows = 2 * h;
        cols = n;
        M.assign((size_t)(rows * cols), 0.0);
        for (long long k = 0; k < h; ++k)
            for (long long j = 0; j < n; ++j) {
                const long double a = angle(k * j, n);
                M[(size_t)((2 * k) * cols + j)] = (double)cosl(a);
                M[(size_t)((2 * k + 1) * cols + j)] = (double)(-sinl(a));
            }
        return 0;
    }
    if (kind == B2F_C2R) {
        const long long h = n / 2 + 1;
        rows = n;
        cols = 2 * h;
        M.assign((size_t)(rows * cols), 0.0);
        for (long long j = 0; j < n; ++j)
            for (long long k = 0; k < h; ++k) {
                const bool self_conj = (k == 0) || (2 * k == n);
                const long double w = self_conj ? 1.0L : 2.0L;
                const long double a = angle(k * j, n);
                M[(size_t)(j * cols + 2 * k)] = (double)(w * cosl(a));
                M[(size_t)(j * cols + 2 * k + 1)] = self_conj ? 0.0 : (double)(-w * sinl(a));
            }
        return 0;
    }
    if (kind >= B2F_REDFT00 && kind <= B2F_RODFT11) {
        rows = cols = n;
        M.assign((size_t)(rows * cols), 0.0);
        // FFTW manual, "1d Real-even DFTs (DCTs)" / "1d Real-odd DFTs (DSTs)".
        // arguments are kept as exact integer ratios: cos(pi * num / den)
        auto cospi = [&](long long num, long long den) {
            num %= (2 * den);
            return cosl(PI * (long double)num / (long double)den);
        };
        auto sinpi = [&](long long num, long long den) {
            num %= (2 * den);
            return sinl(PI * (long double)num / (long double)den);
        };
        for (long long k = 0; k < n; ++k)
            for (long long j = 0; j < n; ++j) {
                long double v = 0;
                switch (kind) {
                    case B2F_REDFT00:  // DCT-I, n >= 2
                        if (n < 2) return -1;
                        if (j == 0) v = 1;
                        else if (j == n - 1) v = (k % 2) ? -1 : 1;
                        else v = 2 * cospi(j * k, n - 1);
                        break;
                    case B2F_REDFT10:  // DCT-II: 2 cos(pi (j+1/2) k / n)
                        v = 2 * cospi((2 * j + 1) * k, 2 * n);
                        break;
                    case B2F_REDFT01:  // DCT-III: x0 + 2 sum cos(pi j (k+1/2)/n)
                        v = (j == 0) ? 1 : 2 * cospi(j * (2 * k + 1), 2 * n);
                        break;
                    case B2F_REDFT11:  // DCT-IV
                        v = 2 * cospi((2 * j + 1) * (2 * k + 1), 4 * n);
                        break;
                    case B2F_RODFT00:  // DST-I
                        v = 2 * sinpi((j + 1) * (k + 1), n + 1);
                        break;
                    case B2F_RODFT10:  // DST-II: 2 sin(pi (j+1/2)(k+1)/n)
                        v = 2 * sinpi((2 * j + 1) * (k + 1), 2 * n);
                        break;
                    case B2F_RODFT01:  // DST-III
                        v = (j == n - 1) ? ((k % 2) ? -1 : 1) : 2 * sinpi((j + 1) * (2 * k + 1), 2 * n);
                        break;
                    case B2F_RODFT11:  // DST-IV
                        v = 2 * sinpi((2 * j + 1) * (2 * k + 1), 4 * n);
                        break;
                }
                M[(size_t)(k * cols + j)] = (double)v;
            }
        return 0;
    }
    return -1;
}

// ---- device ---------------------------------------------------------------
// one CTA = pb consecutive pencils (consecutive inner positions when inner > 1);
// the input pencils are staged in shared memory first, so in == out is safe.
template <class T>
__global__ void dft_matrix_kernel(const GenericParams prm) {
    extern __shared__ double b2f_gsm[];   // [cols][pb]
    const T* in = reinterpret_cast<const T*>(prm.in);
    T* out = reinterpret_cast<T*>(prm.out);
    const long long g0 = (long long)blockIdx.x * prm.pb;
    const int pb = prm.pb;
    for (int idx = threadIdx.x; idx < prm.cols * pb; idx += blockDim.x) {
        const int pl = idx % pb, c = idx / pb;
        const long long g = g0 + pl;
        double v = 0.0;
        if (g < prm.npencils) {
            const long long o = g / prm.inner, i = g - o * prm.inner;
            const long long n = c / prm.in_c;
            const int part = c - (int)n * prm.in_c;
            v = (double)in[((o * prm.in_n + n) * prm.inner + i) * prm.in_c + part];
        }
        b2f_gsm[c * pb + pl] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < prm.rows * pb; idx += blockDim.x) {
        const int pl = idx % pb, r = idx / pb;
        const long long g = g0 + pl;
        if (g >= prm.npencils) continue;
        const double* __restrict__ mrow = prm.M + (size_t)r * prm.cols;
        double acc = 0.0;
        for (int c = 0; c < prm.cols; ++c) acc = fma(mrow[c], b2f_gsm[c * pb + pl], acc);
        const long long o = g / prm.inner, i = g - o * prm.inner;
        const long long k = r / prm.out_c;
        const int part = r - (int)k * prm.out_c;
        out[((o * prm.out_n + k) * prm.inner + i) * prm.out_c + part] = (T)(acc * prm.scale);
    }
}

cudaError_t launch_generic(int precision, const GenericParams& prm_in, cudaStream_t st) {
    GenericParams prm = prm_in;
    // pencils per CTA: as many as fit 64 KiB of staging, at most 32
    long long pb = (64 * 1024) / ((long long)prm.cols * 8);
    if (pb > 32) pb = 32;
    if (pb < 1) pb = 1;
    if (prm.inner > 1 && pb > prm.inner) pb = prm.inner;   // keep a CTA inside one outer index (coalescing)
    prm.pb = (int)pb;
    const size_t smem = (size_t)prm.cols * pb * 8;
    const long long grid = (prm.npencils + pb - 1) / pb;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    auto kern = (precision == 8) ? dft_matrix_kernel<double> : dft_matrix_kernel<float>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)grid, 256, smem, st>>>(prm);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2f
