// chirpz_host.h -- host-side tables of the chirp-z (Bluestein) kernels (chirpz.cuh).
//
// Every 1-D transform FFTW's guru interface offers the reference
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:49-76: c2c, r2c, c2r and the
// eight r2r kinds, utilities.pyx:7-26) is a sum
//        y[k] = Re|id { post[k] * sum_j  pre[j] x[j] * f[k - j] }
// once the kernel  exp(-+ i pi (j+a)(k+b)/L)  is split with
//        (j+a)(k+b) = ( (j+a)^2 + (k+b)^2 - (k-j+b-a)^2 ) / 2 :
//   pre[j]  = g_j exp(-+ i pi (2j+2a)^2 / 8L)      (g_j: the endpoint weights of DCT-I/III, DST-III)
//   post[k] =     exp(-+ i pi (2k+2b)^2 / 8L)      (times i for the sine kinds: Re(i z) = -Im z)
//   f[m]    =     exp(+- i pi (2m+2b-2a)^2 / 8L),  m = -(n_in-1) .. n_out-1
// The convolution runs as two power-of-two FFTs of length M >= n_in + n_out - 1.
// Tables are evaluated in long double with the angle reduced as an exact integer
// ratio; the filter spectrum (divided by M) is computed here with a long double
// radix-2 FFT.  Host only: shared by the library and by tests/emu.
#pragma once
#include <math.h>
#include <vector>

namespace b2f {

struct ChirpSpec {
    long long n_in = 0, n_out = 0;   // points read / written per pencil (hermitian: n_in = logical length)
    int in_mode = 0;                 // 0 complex, 1 real, 2 hermitian half spectrum (n_in/2+1 stored)
    int out_real = 0;                // 1: the real part is stored
    int M = 0;                       // convolution length (power of two)
    std::vector<long double> pre, post, filt;   // interleaved re, im; filt = FFT_M(f) / M
};

inline void chirp_unit(long long num, long long den, int sign, long double* re, long double* im) {
    // exp(sign * i * pi * num / den), num reduced modulo 2*den first
    const long double PI = 3.14159265358979323846264338327950288L;
    long long r = num % (2 * den);
    if (r < 0) r += 2 * den;
    const long double a = PI * (long double)r / (long double)den;
    *re = cosl(a);
    *im = (long double)sign * sinl(a);
}

inline void chirp_fft_ld(std::vector<long double>& a, int M) {   // in place, forward, interleaved
    for (int i = 1, j = 0; i < M; ++i) {
        int bit = M >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            std::swap(a[2 * i], a[2 * j]);
            std::swap(a[2 * i + 1], a[2 * j + 1]);
        }
    }
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int len = 2; len <= M; len <<= 1) {
        for (int i = 0; i < M; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const long double ang = -2 * PI * (long double)k / (long double)len;
                const long double wr = cosl(ang), wi = sinl(ang);
                long double* u = &a[2 * (i + k)];
                long double* v = &a[2 * (i + k + len / 2)];
                const long double tr = v[0] * wr - v[1] * wi, ti = v[0] * wi + v[1] * wr;
                v[0] = u[0] - tr;
                v[1] = u[1] - ti;
                u[0] += tr;
                u[1] += ti;
            }
    }
}

// kind: FFTW integers (-1, +1 c2c; -2 r2c; +2 c2r; 3..10 r2r); n = logical length
// (the real length for r2c / c2r).  Returns 0, or -1 for an unknown kind / n.
inline int chirp_build(int kind, long long n, ChirpSpec* sp) {
    if (n < 1) return -1;
    long long den;          // 8L
    int a2 = 0, b2 = 0;     // 2a, 2b
    int sign = -1;          // sign of the transform's exponent
    bool sine = false;
    sp->in_mode = 0;
    sp->out_real = 0;
    sp->n_in = sp->n_out = n;
    switch (kind) {
        case -1: den = 4 * n; break;
        case 1: den = 4 * n; sign = 1; break;
        case -2: den = 4 * n; sp->in_mode = 1; sp->n_out = n / 2 + 1; break;
        case 2: den = 4 * n; sign = 1; sp->in_mode = 2; sp->out_real = 1; break;
        case 3: if (n < 2) return -1; den = 8 * (n - 1); break;                     // REDFT00
        case 4: den = 8 * n; b2 = 1; break;                                        // REDFT01
        case 5: den = 8 * n; a2 = 1; break;                                        // REDFT10
        case 6: den = 8 * n; a2 = 1; b2 = 1; break;                                // REDFT11
        case 7: den = 8 * (n + 1); a2 = 2; b2 = 2; sine = true; break;             // RODFT00
        case 8: den = 8 * n; a2 = 2; b2 = 1; sine = true; break;                   // RODFT01
        case 9: den = 8 * n; a2 = 1; b2 = 2; sine = true; break;                   // RODFT10
        case 10: den = 8 * n; a2 = 1; b2 = 1; sine = true; break;                  // RODFT11
        default: return -1;
    }
    if (kind >= 3) {
        sp->in_mode = 1;
        sp->out_real = 1;
    }
    const long long ni = sp->n_in, no = sp->n_out;
    int M = 1;
    while (M < ni + no - 1) M <<= 1;
    sp->M = M;
    sp->pre.assign((size_t)(2 * ni), 0.0L);
    sp->post.assign((size_t)(2 * no), 0.0L);
    for (long long j = 0; j < ni; ++j) {
        long double g = 1.0L;
        if (kind >= 3) {
            g = 2.0L;
            if (kind == 3 && (j == 0 || j == n - 1)) g = 1.0L;
            if (kind == 4 && j == 0) g = 1.0L;
            if (kind == 8 && j == n - 1) g = 1.0L;
        }
        long double re, im;
        chirp_unit((2 * j + a2) * (2 * j + a2), den, sign, &re, &im);
        sp->pre[2 * j] = g * re;
        sp->pre[2 * j + 1] = g * im;
    }
    for (long long k = 0; k < no; ++k) {
        long double re, im;
        chirp_unit((2 * k + b2) * (2 * k + b2), den, sign, &re, &im);
        if (sine) {   // times i
            const long double t = re;
            re = -im;
            im = t;
        }
        sp->post[2 * k] = re;
        sp->post[2 * k + 1] = im;
    }
    std::vector<long double> f((size_t)(2 * M), 0.0L);
    for (long long m = -(ni - 1); m <= no - 1; ++m) {
        long double re, im;
        const long long t = 2 * m + b2 - a2;
        chirp_unit(t * t, den, -sign, &re, &im);
        const long long idx = m < 0 ? m + M : m;
        f[2 * idx] = re;
        f[2 * idx + 1] = im;
    }
    chirp_fft_ld(f, M);
    for (auto& x : f) x /= (long double)M;
    sp->filt.swap(f);
    return 0;
}

}  // namespace b2f
