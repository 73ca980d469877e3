/*
 * dft_ref.c -- plain C restatement of the 1-D transforms the reference obtains
 * from FFTW (TEST INFRASTRUCTURE: oracle, never linked into the product).
 *
 * The arithmetic of the hot path lives in a third-party library that is absent
 * from /root/reference (FFTW 3, located at build time by name only,
 * /root/reference/setup.py:64-81; no version pinned).  What FFTW computes is
 * published: the unnormalised DFT with exponent sign -1 (forward) / +1
 * (backward), the r2c/c2r half-spectrum forms and the eight r2r kinds of the
 * FFTW manual ("What FFTW Really Computes").  The reference selects among them
 * in fftw_planxfftn (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:49-76)
 * with the kind integers of fftw/utilities.pyx:7-26.  This file evaluates those
 * definitions directly, O(n^2), in long double -- an independent check of the
 * numpy/scipy (pocketfft) oracle and of the CUDA kernels.
 *
 * kind: -1 c2c forward, +1 c2c backward, -2 r2c, +2 c2r (n = real length),
 *       3..10 = REDFT00, REDFT01, REDFT10, REDFT11, RODFT00, RODFT01, RODFT10, RODFT11.
 * in/out are double arrays (complex = interleaved re,im).  Returns 0 on success.
 */
#include <math.h>

static const long double PI_L = 3.14159265358979323846264338327950288L;

int dft_ref(int kind, long n, const double *in, double *out)
{
    long j, k;
    if (n < 1) return -1;
    if (kind == -1 || kind == 1) {
        for (k = 0; k < n; k++) {
            long double re = 0, im = 0;
            for (j = 0; j < n; j++) {
                long double a = 2 * PI_L * (long double)((k * j) % n) / (long double)n;
                long double c = cosl(a), s = kind * sinl(a);
                re += in[2 * j] * c - in[2 * j + 1] * s;
                im += in[2 * j] * s + in[2 * j + 1] * c;
            }
            out[2 * k] = (double)re;
            out[2 * k + 1] = (double)im;
        }
        return 0;
    }
    if (kind == -2) {                       /* r2c: n reals -> n/2+1 complex */
        for (k = 0; k <= n / 2; k++) {
            long double re = 0, im = 0;
            for (j = 0; j < n; j++) {
                long double a = 2 * PI_L * (long double)((k * j) % n) / (long double)n;
                re += in[j] * cosl(a);
                im -= in[j] * sinl(a);
            }
            out[2 * k] = (double)re;
            out[2 * k + 1] = (double)im;
        }
        return 0;
    }
    if (kind == 2) {                        /* c2r: n/2+1 complex -> n reals */
        for (j = 0; j < n; j++) {
            long double acc = in[0];
            for (k = 1; k <= n / 2; k++) {
                long double a = 2 * PI_L * (long double)((k * j) % n) / (long double)n;
                if (2 * k == n) acc += in[2 * k] * cosl(a);
                else acc += 2 * (in[2 * k] * cosl(a) - in[2 * k + 1] * sinl(a));
            }
            out[j] = (double)acc;
        }
        return 0;
    }
    if (kind < 3 || kind > 10) return -1;
    for (k = 0; k < n; k++) {
        long double acc = 0;
        for (j = 0; j < n; j++) {
            long double x = in[j];
            switch (kind) {
            case 3:  /* REDFT00: X0 + (-1)^k X_{n-1} + 2 sum_{1}^{n-2} X_j cos(pi j k/(n-1)) */
                if (n < 2) return -1;
                if (j == 0) acc += x;
                else if (j == n - 1) acc += (k % 2 ? -x : x);
                else acc += 2 * x * cosl(PI_L * (long double)j * k / (long double)(n - 1));
                break;
            case 5:  /* REDFT10: 2 sum X_j cos(pi (j+1/2) k / n) */
                acc += 2 * x * cosl(PI_L * ((long double)j + 0.5L) * k / (long double)n);
                break;
            case 4:  /* REDFT01: X0 + 2 sum_{1}^{n-1} X_j cos(pi j (k+1/2)/n) */
                if (j == 0) acc += x;
                else acc += 2 * x * cosl(PI_L * (long double)j * ((long double)k + 0.5L) / (long double)n);
                break;
            case 6:  /* REDFT11 */
                acc += 2 * x * cosl(PI_L * ((long double)j + 0.5L) * ((long double)k + 0.5L) / (long double)n);
                break;
            case 7:  /* RODFT00: 2 sum X_j sin(pi (j+1)(k+1)/(n+1)) */
                acc += 2 * x * sinl(PI_L * (long double)(j + 1) * (k + 1) / (long double)(n + 1));
                break;
            case 9:  /* RODFT10: 2 sum X_j sin(pi (j+1/2)(k+1)/n) */
                acc += 2 * x * sinl(PI_L * ((long double)j + 0.5L) * (k + 1) / (long double)n);
                break;
            case 8:  /* RODFT01: (-1)^k X_{n-1} + 2 sum_{0}^{n-2} X_j sin(pi (j+1)(k+1/2)/n) */
                if (j == n - 1) acc += (k % 2 ? -x : x);
                else acc += 2 * x * sinl(PI_L * (long double)(j + 1) * ((long double)k + 0.5L) / (long double)n);
                break;
            case 10: /* RODFT11 */
                acc += 2 * x * sinl(PI_L * ((long double)j + 0.5L) * ((long double)k + 0.5L) / (long double)n);
                break;
            }
        }
        out[k] = (double)acc;
    }
    return 0;
}
