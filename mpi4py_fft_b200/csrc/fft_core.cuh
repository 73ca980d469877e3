// fft_core.cuh -- per-thread building blocks of the sm_100a Stockham kernels.
//
// Everything here is __host__ __device__ and free of CUDA built-ins so that
// tests/emu/emu_fft.cpp can step the exact per-thread code of a kernel on the
// CPU (thread by thread, barrier phase by barrier phase) before a GPU is ever
// involved.  Replaces the FFTW codelets behind fftw_plan_guru_dft
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:52-56).
#pragma once
#ifndef B2F_TW_POW2
#define B2F_TW_POW2 1
#endif
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B2F_HD __host__ __device__ __forceinline__
#define B2F_HDC __host__ __device__ constexpr
#else
#define B2F_HD inline
#define B2F_HDC constexpr
#endif

namespace b2f {

template <class T>
struct alignas(2 * sizeof(T)) cplx {
    T x, y;
};

template <class T> B2F_HD cplx<T> operator+(cplx<T> a, cplx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <class T> B2F_HD cplx<T> operator-(cplx<T> a, cplx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <class T> B2F_HD cplx<T> cmul(cplx<T> a, cplx<T> b) {
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// multiply by -i  (the forward W_4)
template <class T> B2F_HD cplx<T> mul_mi(cplx<T> a) { return {a.y, -a.x}; }

// ---------------------------------------------------------------------------
// multiply by W_32^s = exp(-2*pi*i*s/32), s a loop constant in [0,16)
// ---------------------------------------------------------------------------
template <class T>
B2F_HD cplx<T> mul_w32(cplx<T> a, int s) {
    // cos/sin(2*pi*s/32), s = 0..15
    constexpr double C[16] = {
        1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
        0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785,
        0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
        -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913};
    constexpr double S[16] = {
        0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
        0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913,
        1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
        0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
    if (s == 0) return a;
    if (s == 8) return mul_mi(a);
    if (s == 4) {   // (1 - i)/sqrt2
        const T h = (T)0.70710678118654752440;
        return {(a.x + a.y) * h, (a.y - a.x) * h};
    }
    if (s == 12) {  // (-1 - i)/sqrt2
        const T h = (T)0.70710678118654752440;
        return {(a.y - a.x) * h, -(a.x + a.y) * h};
    }
    const T c = (T)C[s], sn = (T)S[s];
    // (x + iy)(c - i sn)
    return {a.x * c + a.y * sn, a.y * c - a.x * sn};
}

// four-step twiddle W_n^m (m already reduced modulo n), forward sign -1 / backward +1; the angle is
// formed in double for both precisions
template <class T>
B2F_HD cplx<T> fourstep_twiddle(long long m, long long n, bool backward) {
    const double a = 2.0 * (double)m / (double)n;   // in units of pi
    double sn, cs;
#if defined(__CUDA_ARCH__)
    sincospi(a, &sn, &cs);
#else
    sn = sin(3.14159265358979323846 * a);
    cs = cos(3.14159265358979323846 * a);
#endif
    return {(T)cs, (T)(backward ? sn : -sn)};
}

// ---------------------------------------------------------------------------
// R-point forward DFT on registers, natural order in and out (R = 2..32).
// Decimation in time by two; all indices are compile-time constants after
// unrolling so v/e/o live in registers.
// ---------------------------------------------------------------------------
// multiply by W_24^s = exp(-2*pi*i*s/24), s a loop constant in [0,12): the
// butterfly twiddles of the radices 6, 12 and 24 (lengths 3 * 2^k)
template <class T>
B2F_HD cplx<T> mul_w24(cplx<T> a, int s) {
    constexpr double C[12] = {
        1.0, 0.96592582628906828675, 0.86602540378443864676, 0.70710678118654752440,
        0.5, 0.25881904510252076235, 0.0, -0.25881904510252076235,
        -0.5, -0.70710678118654752440, -0.86602540378443864676, -0.96592582628906828675};
    constexpr double S[12] = {
        0.0, 0.25881904510252076235, 0.5, 0.70710678118654752440,
        0.86602540378443864676, 0.96592582628906828675, 1.0, 0.96592582628906828675,
        0.86602540378443864676, 0.70710678118654752440, 0.5, 0.25881904510252076235};
    if (s == 0) return a;
    if (s == 6) return mul_mi(a);
    const T c = (T)C[s], sn = (T)S[s];
    return {a.x * c + a.y * sn, a.y * c - a.x * sn};
}

// multiply by W_40^s = exp(-2*pi*i*s/40), s a loop constant in [0,20): the butterfly
// twiddles of the radices 10 and 20 (lengths 5 * 2^k)
template <class T>
B2F_HD cplx<T> mul_w40(cplx<T> a, int s) {
    constexpr double C[20] = {1.0, 0.98768834059513777035, 0.95105651629515353118, 0.89100652418836789881, 0.80901699437494745126, 0.70710678118654757274, 0.58778525229247313710, 0.45399049973954680448, 0.30901699437494745126, 0.15643446504023092447, 0.00000000000000006123, -0.15643446504023059140, -0.30901699437494734024, -0.45399049973954669346, -0.58778525229247302608, -0.70710678118654746172, -0.80901699437494734024, -0.89100652418836778779, -0.95105651629515353118, -0.98768834059513765933};
    constexpr double S[20] = {0.0, 0.15643446504023086896, 0.30901699437494739575, 0.45399049973954674897, 0.58778525229247313710, 0.70710678118654746172, 0.80901699437494745126, 0.89100652418836778779, 0.95105651629515353118, 0.98768834059513777035, 1.0, 0.98768834059513777035, 0.95105651629515364220, 0.89100652418836789881, 0.80901699437494745126, 0.70710678118654757274, 0.58778525229247324813, 0.45399049973954685999, 0.30901699437494750677, 0.15643446504023097998};
    if (s == 0) return a;
    if (s == 10) return mul_mi(a);
    const T c = (T)C[s], sn = (T)S[s];
    return {a.x * c + a.y * sn, a.y * c - a.x * sn};
}
// multiply by W_56^s = exp(-2*pi*i*s/56), s in [0,28): radices 14 and 28 (lengths 7 * 2^k)
template <class T>
B2F_HD cplx<T> mul_w56(cplx<T> a, int s) {
    constexpr double C[28] = {1.0, 0.99371220989324260398, 0.97492791218182361934, 0.94388333030836757409, 0.90096886790241914600, 0.84672419922828412453, 0.78183148246802980363, 0.70710678118654757274, 0.62348980185873359439, 0.53203207651533657163, 0.43388373911755817591, 0.33027906195516731902, 0.22252093395631444839, 0.11196447610330768907, 0.00000000000000006123, -0.11196447610330757805, -0.22252093395631433737, -0.33027906195516720800, -0.43388373911755806489, -0.53203207651533646061, -0.62348980185873348336, -0.70710678118654746172, -0.78183148246802947057, -0.84672419922828412453, -0.90096886790241903498, -0.94388333030836757409, -0.97492791218182373036, -0.99371220989324260398};
    constexpr double S[28] = {0.0, 0.11196447610330785560, 0.22252093395631439288, 0.33027906195516709698, 0.43388373911755812040, 0.53203207651533657163, 0.62348980185873348336, 0.70710678118654746172, 0.78183148246802980363, 0.84672419922828412453, 0.90096886790241914600, 0.94388333030836746307, 0.97492791218182361934, 0.99371220989324260398, 1.0, 0.99371220989324260398, 0.97492791218182361934, 0.94388333030836746307, 0.90096886790241914600, 0.84672419922828423555, 0.78183148246802991466, 0.70710678118654757274, 0.62348980185873392745, 0.53203207651533668265, 0.43388373911755823142, 0.33027906195516720800, 0.22252093395631408757, 0.11196447610330798050};
    if (s == 0) return a;
    if (s == 14) return mul_mi(a);
    const T c = (T)C[s], sn = (T)S[s];
    return {a.x * c + a.y * sn, a.y * c - a.x * sn};
}

template <int R, class T>
struct DftReg {
    static B2F_HD void run(cplx<T>* v) {
        cplx<T> e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        DftReg<R / 2, T>::run(e);
        DftReg<R / 2, T>::run(o);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            cplx<T> t;
            if constexpr ((R & (R - 1)) == 0) t = mul_w32(o[k], k * (32 / R));
            else if constexpr (R % 5 == 0) t = mul_w40(o[k], k * (40 / R));
            else if constexpr (R % 7 == 0) t = mul_w56(o[k], k * (56 / R));
            else t = mul_w24(o[k], k * (24 / R));
            v[k] = e[k] + t;
            v[k + R / 2] = e[k] - t;
        }
    }
};
template <class T>
struct DftReg<3, T> {
    static B2F_HD void run(cplx<T>* v) {
        const T h = (T)0.86602540378443864676;   // sin(2*pi/3)
        const cplx<T> s = v[1] + v[2], d = v[1] - v[2];
        const cplx<T> m = {v[0].x - (T)0.5 * s.x, v[0].y - (T)0.5 * s.y};
        const cplx<T> r = {h * d.y, -h * d.x};   // -i * h * d
        v[0] = v[0] + s;
        v[1] = m + r;
        v[2] = m - r;
    }
};
// 5- and 7-point DFTs: the inputs pair up as v[k] +- v[R-k]; the sums see the cosines, the
// differences the sines (forward sign: y[j] = m[j] - i s[j], y[R-j] = m[j] + i s[j])
template <class T>
struct DftReg<5, T> {
    static B2F_HD void run(cplx<T>* v) {
        const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;   // cos(2 pi / 5), cos(4 pi / 5)
        const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;    // sin(2 pi / 5), sin(4 pi / 5)
        const cplx<T> a1 = v[1] + v[4], a2 = v[2] + v[3], b1 = v[1] - v[4], b2 = v[2] - v[3];
        const cplx<T> m1 = {v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y};
        const cplx<T> m2 = {v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y};
        const cplx<T> t1 = {s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y};
        const cplx<T> t2 = {s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y};
        v[0] = v[0] + a1 + a2;
        v[1] = {m1.x + t1.y, m1.y - t1.x};   // m1 - i t1
        v[4] = {m1.x - t1.y, m1.y + t1.x};
        v[2] = {m2.x + t2.y, m2.y - t2.x};
        v[3] = {m2.x - t2.y, m2.y + t2.x};
    }
};
template <class T>
struct DftReg<7, T> {
    static B2F_HD void run(cplx<T>* v) {
        const T c1 = (T)0.62348980185873353053, c2 = (T)-0.22252093395631440429, c3 = (T)-0.90096886790241912624;
        const T s1 = (T)0.78183148246802980871, s2 = (T)0.97492791218182360702, s3 = (T)0.43388373911755812048;
        const cplx<T> a1 = v[1] + v[6], a2 = v[2] + v[5], a3 = v[3] + v[4];
        const cplx<T> b1 = v[1] - v[6], b2 = v[2] - v[5], b3 = v[3] - v[4];
        // cos(2 pi j k / 7): rows j = 1, 2, 3 over k = 1, 2, 3 are (c1 c2 c3), (c2 c3 c1), (c3 c1 c2);
        // sin(2 pi j k / 7):                                      (s1 s2 s3), (s2 -s3 -s1), (s3 -s1 s2)
        const cplx<T> m1 = {v[0].x + c1 * a1.x + c2 * a2.x + c3 * a3.x, v[0].y + c1 * a1.y + c2 * a2.y + c3 * a3.y};
        const cplx<T> m2 = {v[0].x + c2 * a1.x + c3 * a2.x + c1 * a3.x, v[0].y + c2 * a1.y + c3 * a2.y + c1 * a3.y};
        const cplx<T> m3 = {v[0].x + c3 * a1.x + c1 * a2.x + c2 * a3.x, v[0].y + c3 * a1.y + c1 * a2.y + c2 * a3.y};
        const cplx<T> t1 = {s1 * b1.x + s2 * b2.x + s3 * b3.x, s1 * b1.y + s2 * b2.y + s3 * b3.y};
        const cplx<T> t2 = {s2 * b1.x - s3 * b2.x - s1 * b3.x, s2 * b1.y - s3 * b2.y - s1 * b3.y};
        const cplx<T> t3 = {s3 * b1.x - s1 * b2.x + s2 * b3.x, s3 * b1.y - s1 * b2.y + s2 * b3.y};
        v[0] = v[0] + a1 + a2 + a3;
        v[1] = {m1.x + t1.y, m1.y - t1.x};
        v[6] = {m1.x - t1.y, m1.y + t1.x};
        v[2] = {m2.x + t2.y, m2.y - t2.x};
        v[5] = {m2.x - t2.y, m2.y + t2.x};
        v[3] = {m3.x + t3.y, m3.y - t3.x};
        v[4] = {m3.x - t3.y, m3.y + t3.x};
    }
};
template <class T>
struct DftReg<1, T> {
    static B2F_HD void run(cplx<T>*) {}
};
template <class T>
struct DftReg<2, T> {
    static B2F_HD void run(cplx<T>* v) {
        cplx<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};
template <class T>
struct DftReg<4, T> {
    static B2F_HD void run(cplx<T>* v) {
        cplx<T> t0 = v[0] + v[2], t1 = v[0] - v[2];
        cplx<T> t2 = v[1] + v[3], t3 = mul_mi(v[1] - v[3]);
        v[0] = t0 + t2;
        v[1] = t1 + t3;
        v[2] = t0 - t2;
        v[3] = t1 - t3;
    }
};

// ---------------------------------------------------------------------------
// compile-time radix schedule of an N-point transform held E points per thread
// ---------------------------------------------------------------------------
template <int... Rs>
struct Radices {
    static constexpr int count = sizeof...(Rs);
    static B2F_HDC int get(int i) {
        constexpr int r[sizeof...(Rs)] = {Rs...};
        return r[i];
    }
    static B2F_HDC int product() {
        int p = 1;
        constexpr int r[sizeof...(Rs)] = {Rs...};
        for (int i = 0; i < (int)sizeof...(Rs); ++i) p *= r[i];
        return p;
    }
    // product of the radices before pass s  (the Stockham "Ns")
    static B2F_HDC int before(int s) {
        int p = 1;
        constexpr int r[sizeof...(Rs)] = {Rs...};
        for (int i = 0; i < s; ++i) p *= r[i];
        return p;
    }
    // Twiddle storage: one table per pass s >= 1, laid out [r-1][k] with
    // k = 0..Ns-1 fastest, entry = W_{Ns*R}^{r*k}.  Consecutive butterflies of a
    // pass (consecutive lanes) read consecutive k, so a warp's twiddle load is
    // one coalesced run instead of a gather over a single length-N table.
    static B2F_HDC int tw_offset(int s) {
        int off = 0;
        constexpr int r[sizeof...(Rs)] = {Rs...};
        int ns = r[0];
        for (int i = 1; i < s; ++i) {
            off += (r[i] - 1) * ns;
            ns *= r[i];
        }
        return off;
    }
    static B2F_HDC int tw_total() { return tw_offset((int)sizeof...(Rs)) > 0 ? tw_offset((int)sizeof...(Rs)) : 1; }
};

// host: fill the per-pass twiddle tables of a radix schedule (long double angles)
template <class T, class RAD>
inline void build_pass_twiddles(cplx<T>* out) {
    const long double TWO_PI = 6.28318530717958647692528676655900577L;
    out[0].x = (T)1;
    out[0].y = (T)0;
    int ns = RAD::get(0);
    for (int s = 1; s < RAD::count; ++s) {
        const int R = RAD::get(s);
        cplx<T>* t = out + RAD::tw_offset(s);
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < ns; ++k) {
                const long long num = ((long long)r * k) % ((long long)ns * R);
                const long double a = TWO_PI * (long double)num / (long double)((long long)ns * R);
                t[(r - 1) * ns + k].x = (T)cosl(a);
                t[(r - 1) * ns + k].y = (T)(-sinl(a));
            }
        ns *= R;
    }
}

// host: w[k] = exp(-2 pi i k / (2N)), k = 0..N-1, the split/merge twiddles of the
// real transforms of length 2N (long double angles)
template <class T>
inline void build_real_twiddles(cplx<T>* out, int N) {
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < N; ++k) {
        const long double a = PI * (long double)k / (long double)N;
        out[k].x = (T)cosl(a);
        out[k].y = (T)(-sinl(a));
    }
}

// host: w4[k] = exp(-i pi k / (2 n)), n = 2N, k = 0..N: the quarter-wave twiddles of the
// DCT / DST II and III through a real transform of the same length
template <class T>
inline void build_quarter_twiddles(cplx<T>* out, int N) {
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int k = 0; k <= N; ++k) {
        const long double a = PI * (long double)k / (long double)(4 * (long long)N);
        out[k].x = (T)cosl(a);
        out[k].y = (T)(-sinl(a));
    }
}

// host: the two twiddle sets of the DCT / DST IV of length n = 2N through an N-point complex transform:
// out[m] = exp(-i pi (4m + 1) / (4n)), m < N (input side), out[N + k] = exp(-i pi k / n), k < N (output side)
template <class T>
inline void build_dct4_twiddles(cplx<T>* out, int N) {
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int m = 0; m < N; ++m) {
        const long double a = PI * (long double)(4 * m + 1) / (long double)(8 * (long long)N);
        out[m].x = (T)cosl(a);
        out[m].y = (T)(-sinl(a));
        const long double b = PI * (long double)m / (long double)(2 * (long long)N);
        out[N + m].x = (T)cosl(b);
        out[N + m].y = (T)(-sinl(b));
    }
}

// ---------------------------------------------------------------------------
// Fused redistribution: the last pass of a stage can store straight into the
// arrays of the ranks that own each part of the transformed axis (peer memory
// over NVLink), instead of writing the local block and moving it afterwards.
// It replaces "stage output -> MPI_Alltoallw" of the reference
// (/root/reference/mpi4py_fft/mpifft.py:70-74 with pencil.py:182-183): the axis
// just transformed (extent N here) is the one the following transfer splits.
//
// Point n of the pencil with outer index o and inner index i goes to owner
// w = owner(n) of the balanced distribution of N over p ranks, at element
//     (part * len(w) + n - start(w)) * stride + rest
// of w's array, where (part, rest) depend on the pencil only:
//   mode 0 (the axis the destination gathers comes BEFORE the transformed axis;
//           local block = (P, nD, M, N, Q), destination = (P, ND, M, len(w), Q)):
//           o = (pp*nD + i1)*M + m ->  part = (pp*ND + sD + i1)*M + m,  rest = i,   stride = Q
//   mode 1 (it comes AFTER; local block = (P, N, M, nD, Q), destination =
//           (P, len(w), M, ND, Q)):
//           i = (m*nD + i2)*Q + qq ->  part = o,  rest = (m*ND + sD + i2)*Q + qq,  stride = M*ND*Q
// p == 1 with one base pointer reproduces the plain local layout.
// ---------------------------------------------------------------------------
#define B2F_MAX_PEERS 16
struct PeerStore {
    void* base[B2F_MAX_PEERS];
    long long stride;
    long long M, nD, ND, sD, Q;
    long long ooff, ioff;   // a launch over part of the block: its pencils start at (outer, inner) = (ooff, ioff)
    long long vstride;      // > 0: the launch sees the block re-viewed as rows vstride apart: inner = o * vstride + i, outer = 0
    int p;            // owners of the transformed axis (0 = fused store not in use)
    int q, r;         // N = p*q + r: the first r owners hold q + 1 points
    int mode;

    B2F_HD void locate(long long o, long long i, long long* part, long long* rest) const {
        if (vstride > 0) {
            i += o * vstride;
            o = 0;
        }
        o += ooff;
        i += ioff;
        if (mode == 0) {
            const long long t = o / M, m = o - t * M;
            const long long pp = t / nD, i1 = t - pp * nD;
            *part = (pp * ND + sD + i1) * M + m;
            *rest = i;
        } else {
            const long long t = i / Q, qq = i - t * Q;
            const long long m = t / nD, i2 = t - m * nD;
            *part = o;
            *rest = (m * ND + sD + i2) * Q + qq;
        }
    }
    B2F_HD int owner(int n, int* len, int* start) const {
        const int big = r * (q + 1);
        int w;
        if (n < big) {
            w = n / (q + 1);
            *len = q + 1;
            *start = w * (q + 1);
        } else {
            w = r + (n - big) / q;
            *len = q;
            *start = big + (w - r) * q;
        }
        return w;
    }
};

// ---------------------------------------------------------------------------
// Dealiasing folded into a stage: a padded transform runs at length Np but its
// SPECTRUM side is stored with only N modes (reference libfft.py:263-311,
// FFTBase._truncation_forward / _padding_backward).  Full (c2c) spectrum:
//   kept modes  k <= N/2            -> m = k
//               k >= Np - N/2       -> m = k - (Np - N)
//   for even N both halves meet at m = N/2: forward stores P[N/2] + P[Np - N/2],
//   backward reads T[N/2] / 2 into both P[N/2] and P[Np - N/2].
// Half (r2c) spectrum: kept k < N; for even N the last kept mode is made real and
// doubled (forward) / halved (backward).  n == 0 switches the map off.
// The forward kernel never writes a dropped mode and the backward kernel never
// reads a zero: the separate truncation / zero-fill pass over the spectrum goes.
// ---------------------------------------------------------------------------
struct TruncMap {
    int n;      // kept modes N (0: no truncation)
    int np;     // modes of the padded spectrum Np (full: transform length; half: Np/2 + 1)
    // full spectrum: index in the truncated array of padded mode k, or -1 (dropped);
    // *partner is set when k is the upper copy of an even-N Nyquist mode
    B2F_HD int full(int k, bool* partner) const {
        *partner = false;
        const int h = n / 2;
        if (k <= h) return k;
        if (k >= np - h) {
            *partner = (n % 2 == 0) && (k == np - h);
            return k - (np - n);
        }
        return -1;
    }
    B2F_HD bool even() const { return n % 2 == 0; }
};

// shared-memory index of point i of pencil p inside a CTA tile.
//   CONTIG : pencils are separate rows, row pitch PITCH, one pad slot every
//            2^PS points (keeps stride-R writes of pass 0 conflict free)
//   STRIDED: pencils interleaved [point][pencil] -- a quarter warp touches P
//            neighbouring pencils at the same point
template <bool STRIDED, int P, int N, int PS>
struct SmemIndex {
    static constexpr int padded_n = N + (N >> PS);
    static constexpr int tile_elems = P * padded_n;
    static B2F_HD int at(int p, int i) {
        const int ip = i + (i >> PS);
        return STRIDED ? ip * P + p : p * padded_n + ip;
    }
};

// ---------------------------------------------------------------------------
// One CTA tile: P pencils of N points, E points per thread, TP = N/E threads per
// pencil.  The four phase functions are separated exactly where the kernel
// puts __syncthreads(); none of them keeps state outside v[] and shared memory.
//
//   pass 0      : global -> registers -> R0-point DFTs -> shared
//   pass s (mid): shared -> twiddle -> DFT (read phase) | -> shared (write phase)
//   last pass   : shared -> twiddle -> DFT -> scale -> global
//
// Stockham indexing (autosort, natural order in and out), butterfly j of a pass
// with radix R and Ns = product of earlier radices:
//   inputs   x[j + r*N/R],               r = 0..R-1
//   twiddles W_{Ns*R}^{r*(j mod Ns)}  =  tw[r * (j mod Ns) * N/(Ns*R)]
//   outputs  y[(j div Ns)*Ns*R + (j mod Ns) + r*Ns]
// ---------------------------------------------------------------------------
template <class T, int N, int E, class RAD, int P, bool STRIDED, int PS>
struct TileFFT {
    using C = cplx<T>;
    using Real = T;
    using RADS = RAD;
    using SI = SmemIndex<STRIDED, P, N, PS>;
    static constexpr int LEN = N;
    static constexpr int EPT = E;              // points per thread
    static constexpr int PEN = P;              // pencils per tile
    static constexpr int TP = N / E;           // threads per pencil
    static constexpr int THREADS = TP * P;
    static constexpr int NPASS = RAD::count;
    static_assert(RAD::product() == N, "radix schedule must multiply to N");
    static_assert(N % E == 0, "E must divide N");

    // thread -> (pencil in tile, slot in pencil)
    static B2F_HD int pencil_of(int tid) { return STRIDED ? tid % P : tid / TP; }
    static B2F_HD int slot_of(int tid) { return STRIDED ? tid / P : tid % TP; }

    // twiddles of one butterfly: W^r, r = 1..R-1, for k fixed.  Only the powers
    // of two are loaded (log2 R coalesced loads instead of R-1); the rest are
    // products of at most three loaded values (<= 2 roundings deep), which
    // keeps the L1 wavefront count of the twiddle traffic below the data's.
    template <int R, int Ns>
    static B2F_HD void apply_twiddles(C* v, const C* __restrict__ t) {
#if B2F_TW_POW2
        if constexpr (R <= 4 || (R & (R - 1)) != 0) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], t[(r - 1) * Ns]);
        } else {
            // r = 8a + b: lo[b] = W^b (b < 8), hi = W^(8a)
            C lo[8];
            lo[1] = t[0];
            lo[2] = t[Ns];
            lo[4] = t[3 * Ns];
            lo[3] = cmul(lo[2], lo[1]);
            lo[5] = cmul(lo[4], lo[1]);
            lo[6] = cmul(lo[4], lo[2]);
            lo[7] = cmul(lo[4], lo[3]);
#pragma unroll
            for (int b = 1; b < 8; ++b) v[b] = cmul(v[b], lo[b]);
            if constexpr (R > 8) {
                C w8 = t[7 * Ns];
                C hi = w8;
#pragma unroll
                for (int a = 1; a < R / 8; ++a) {
                    if (a == 2) hi = t[15 * Ns];          // W^16 is in the table
                    else if (a == 3) hi = cmul(hi, w8);   // W^24 = W^16 * W^8
                    v[8 * a] = cmul(v[8 * a], hi);
#pragma unroll
                    for (int b = 1; b < 8; ++b) v[8 * a + b] = cmul(v[8 * a + b], cmul(hi, lo[b]));
                }
            }
        }
#else
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], t[(r - 1) * Ns]);
#endif
    }

    template <int S>
    static B2F_HD void twiddle_dft(C* v, int q, const C* __restrict__ tw) {
        constexpr int R = RAD::get(S);
        constexpr int Ns = RAD::before(S);
        constexpr int NB = E / R;
        static_assert(E % R == 0, "radix must divide E");
        static_assert(R <= 32, "radix above 32 is not implemented");
        static_assert((R & (R - 1)) == 0 || R == 3 || R == 6 || R == 12 || R == 24 || R == 5 || R == 10 || R == 20 ||
                          R == 7 || R == 14 || R == 28,
                      "radix must be 2^a, 3 * 2^a, 5 * 2^a or 7 * 2^a");
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (S > 0) {
                const int j = q + b * TP;
                const int k = j % Ns;
                constexpr int OFF = RAD::tw_offset(S);
                apply_twiddles<R, Ns>(v + b * R, tw + OFF + k);
            }
            DftReg<R, T>::run(v + b * R);
        }
    }

    // ---- pass 0 load: v[b*R0 + r] = x[j + r*N/R0], j = q + b*TP -------------
    // gin points at point 0 of this thread's pencil; nstride in elements
    static B2F_HD void load_global(C* v, int q, const C* __restrict__ gin, long long nstride,
                                   bool valid, bool swap) {
        constexpr int R = RAD::get(0);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                C a = {(T)0, (T)0};
                if (valid) a = gin[(long long)n * nstride];
                if (swap) { T t = a.x; a.x = a.y; a.y = t; }
                v[b * R + r] = a;
            }
        }
    }

    // ---- shared write after pass S -----------------------------------------
    template <int S>
    static B2F_HD void store_shared(const C* v, int p, int q, C* smem) {
        constexpr int R = RAD::get(S);
        constexpr int Ns = RAD::before(S);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = q + b * TP;
            const int base = (j / Ns) * (Ns * R) + (j % Ns);
#pragma unroll
            for (int r = 0; r < R; ++r) smem[SI::at(p, base + r * Ns)] = v[b * R + r];
        }
    }

    // ---- shared read before pass S -----------------------------------------
    template <int S>
    static B2F_HD void load_shared(C* v, int p, int q, const C* smem) {
        constexpr int R = RAD::get(S);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b * R + r] = smem[SI::at(p, q + b * TP + r * (N / R))];
        }
    }

    // ---- last pass store: y[j + r*N/R] (Ns*R == N so the output index of
    //      butterfly j < N/R collapses to j + r*Ns) ----------------------------
    static B2F_HD void store_global(const C* v, int q, C* __restrict__ gout, long long nstride,
                                    bool valid, bool swap, T scale) {
        constexpr int R = RAD::get(NPASS - 1);
        constexpr int NB = E / R;
        if (!valid) return;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                C a = v[b * R + r];
                a.x *= scale;
                a.y *= scale;
                if (swap) { T t = a.x; a.x = a.y; a.y = t; }
                gout[(long long)n * nstride] = a;
            }
        }
    }
    // ---- pass 0 load of a padded backward transform: the input holds tm.n modes ----
    static B2F_HD void load_global_padded(C* v, int q, const C* __restrict__ gin, long long nstride, bool valid,
                                          bool swap, const TruncMap& tm) {
        constexpr int R = RAD::get(0);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                // the load itself is unconditional (a dropped mode re-reads mode 0, a cache hit) so
                // that all E loads of the thread are in flight together; the value is selected after
                C a = {(T)0, (T)0};
                bool partner;
                const int m = tm.full(n, &partner);
                if (valid) a = gin[(long long)(m < 0 ? 0 : m) * nstride];
                const T f = m < 0 ? (T)0 : (tm.even() && m == tm.n / 2) ? (T)0.5 : (T)1;
                a.x *= f;
                a.y *= f;
                if (swap) { T t = a.x; a.x = a.y; a.y = t; }
                v[b * R + r] = a;
            }
        }
    }
    // ---- last pass store of a padded forward transform: only the kept modes are written;
    //      nyq[p] carries P[Np - N/2] to the thread that stores m = N/2 (even N).
    //      Phase A (before the barrier): publish the partner; phase B: store.
    static B2F_HD void store_truncated_publish(const C* v, int p, int q, C* nyq, const TruncMap& tm, T scale) {
        constexpr int R = RAD::get(NPASS - 1);
        constexpr int NB = E / R;
        if (!tm.even()) return;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                if (n == tm.np - tm.n / 2) nyq[p] = {v[b * R + r].x * scale, v[b * R + r].y * scale};
            }
        }
    }
    static B2F_HD void store_truncated(const C* v, int p, int q, C* __restrict__ gout, long long nstride, bool valid,
                                       bool swap, T scale, const C* nyq, const TruncMap& tm) {
        constexpr int R = RAD::get(NPASS - 1);
        constexpr int NB = E / R;
        if (!valid) return;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                bool partner;
                const int m = tm.full(n, &partner);
                if (m < 0 || partner) continue;
                C a = v[b * R + r];
                a.x *= scale;
                a.y *= scale;
                if (tm.even() && m == tm.n / 2) {
                    a.x += nyq[p].x;
                    a.y += nyq[p].y;
                }
                if (swap) { T t = a.x; a.x = a.y; a.y = t; }
                gout[(long long)m * nstride] = a;
            }
        }
    }
    // ---- last pass store into the owners' arrays (see PeerStore) ----------------
    static B2F_HD void store_peer(const C* v, int q, const PeerStore& ps, long long part, long long rest,
                                  bool valid, bool swap, T scale) {
        constexpr int R = RAD::get(NPASS - 1);
        constexpr int NB = E / R;
        if (!valid) return;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                C a = v[b * R + r];
                a.x *= scale;
                a.y *= scale;
                if (swap) { T t = a.x; a.x = a.y; a.y = t; }
                int len, start;
                const int w = ps.owner(n, &len, &start);
                C* dst = reinterpret_cast<C*>(ps.base[w]);
                dst[(part * len + (n - start)) * ps.stride + rest] = a;
            }
        }
    }
    // ---- real transforms: 2N reals <-> N+1 complex through the N-point complex
    // FFT of the packed pencil z[j] = x[2j] + i x[2j+1]  (replaces
    // fftw_plan_guru_dft_r2c / _c2r, /root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:57-67).
    // w[k] = exp(-2 pi i k / 2N), k < N.  X[N] lives in the extra slot tile_elems + p.
    //
    // r2c: with Z the transform of z (natural order in shared memory),
    //   X[k] = (Z[k] + conj Z[N-k])/2 - i w^k (Z[k] - conj Z[N-k])/2,   X[N] = Re Z[0] - Im Z[0]
    // keep > 0: only the first `keep` modes of the half spectrum are stored (padded transform);
    // for even keep the last one is made real and doubled (reference libfft.py:270-279)
    static B2F_HD void r2c_post(int p, int q, const C* smem, const C* __restrict__ w, C* __restrict__ gout,
                                long long out_ns, bool valid, T scale, int keep = 0) {
        if (!valid) return;
        const T h = (T)0.5 * scale;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = q + e * TP;
            const C a = smem[SI::at(p, k)];
            C b = smem[SI::at(p, k == 0 ? 0 : N - k)];
            b.y = -b.y;
            const C sm = a + b, d = a - b;
            const C t = cmul(w[k], d);
            C x = {(sm.x + t.y) * h, (sm.y - t.x) * h};
            if (keep > 0 && keep % 2 == 0 && k == keep - 1) {
                x.x *= (T)2;
                x.y = (T)0;
            }
            if (keep == 0 || k < keep) gout[(long long)k * out_ns] = x;
            if (k == 0 && (keep == 0 || N < keep)) {
                C xn = {(a.x - a.y) * scale, (T)0};
                if (keep > 0 && keep % 2 == 0 && N == keep - 1) xn.x *= (T)2;
                gout[(long long)N * out_ns] = xn;
            }
        }
    }

    // c2r: the pass-0 inputs  Z'[n] = (X[n] + conj X[N-n]) + i conj(w^n) (X[n] - conj X[N-n])
    // (twice the packed spectrum, so that the unnormalised result is FFTW's), already
    // re/im-swapped for the forward-kernel-as-backward trick.  The imaginary parts of
    // X[0] and X[N] are ignored, as FFTW does.
    static B2F_HD void c2r_pre(C* v, int p, int q, const C* smem, const C* __restrict__ w) {
        constexpr int R = RAD::get(0);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int n = q + b * TP + r * (N / R);
                C a = smem[SI::at(p, n)];
                C c = (n == 0) ? smem[SI::tile_elems + p] : smem[SI::at(p, N - n)];
                if (n == 0) {
                    a.y = (T)0;
                    c.y = (T)0;
                }
                c.y = -c.y;
                const C sm = a + c, d = a - c;
                C wc = w[n];
                wc.y = -wc.y;
                const C t = cmul(wc, d);
                v[b * R + r] = {sm.y + t.x, sm.x - t.y};   // (re, im) = (sm.x - t.y, sm.y + t.x), swapped
            }
        }
    }

    // ---- r2r kinds II and III of even length n = 2N through the real transforms above
    // (replaces fftw_plan_guru_r2r for FFTW_REDFT10 / REDFT01 / RODFT10 / RODFT01,
    // /root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:69-75).  Makhoul's permutation
    //     v[j] = x[2j],  v[n-1-j] = x[2j+1]                       (sigma below: v[j] = x[sigma(j)])
    // turns the DCT-II into the real DFT V of v followed by a quarter-wave twiddle,
    //     Y[k] = 2 Re(w4^k V[k]),  Y[n-k] = -2 Im(w4^k V[k]),   w4 = exp(-i pi / 2n),
    // and the DCT-III into its transpose: V[k] = (X[k] - i X[n-k]) conj(w4^k) (X[n] = 0), v = c2r(V),
    // y[sigma(j)] = v[j].  The sine kinds are the cosine kinds of the sign-alternated input with the
    // output reversed (II), or of the reversed input with the output sign-alternated (III): `flip`.
    static B2F_HD int r2r_sigma(int j) { return j < N ? 2 * j : 2 * (2 * N - 1 - j) + 1; }
    static B2F_HD int r2r_pos(int k, bool flip) { return flip ? 2 * N - 1 - k : k; }
    // kind II, pass-0 load: z[m] = v[2m] + i v[2m+1]
    static B2F_HD void r2r_load(C* v, int q, const T* __restrict__ gin, long long ns, bool valid, bool flip) {
        constexpr int R = RAD::get(0);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int m = q + b * TP + r * (N / R);
                const int i0 = r2r_sigma(2 * m), i1 = r2r_sigma(2 * m + 1);
                C a = {(T)0, (T)0};
                if (valid) {
                    a.x = gin[(long long)i0 * ns];
                    a.y = gin[(long long)i1 * ns];
                }
                if (flip) {
                    if (i0 & 1) a.x = -a.x;
                    if (i1 & 1) a.y = -a.y;
                }
                v[b * R + r] = a;
            }
        }
    }
    // kind II, after the N-point transform sits in shared memory in natural order: split as r2c_post,
    // twiddle, store two reals per mode
    static B2F_HD void r2r_post(int p, int q, const C* smem, const C* __restrict__ w, const C* __restrict__ w4,
                                T* __restrict__ gout, long long out_ns, bool valid, T scale, bool flip) {
        if (!valid) return;
        const T h = (T)0.5 * scale;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = q + e * TP;
            const C a = smem[SI::at(p, k)];
            C b = smem[SI::at(p, k == 0 ? 0 : N - k)];
            b.y = -b.y;
            const C sm = a + b, d = a - b;
            const C t = cmul(w[k], d);
            const C x = {(sm.x + t.y) * h, (sm.y - t.x) * h};          // V[k] (scaled)
            const C y = cmul(w4[k], x);
            gout[(long long)r2r_pos(k, flip) * out_ns] = (T)2 * y.x;
            if (k > 0) gout[(long long)r2r_pos(2 * N - k, flip) * out_ns] = (T)-2 * y.y;
            if (k == 0) {
                const T vn = (a.x - a.y) * scale;                       // V[N], real
                gout[(long long)r2r_pos(N, flip) * out_ns] = (T)2 * w4[N].x * vn;
            }
        }
    }
    // kind III: the half spectrum V[k], k = 0..N, built from the real input straight into shared memory
    // (where c2r_pre expects it: V[N] in the extra slot)
    static B2F_HD void r2r_fill(int p, int q, C* smem, const C* __restrict__ w4, const T* __restrict__ gin,
                                long long ns, bool valid, bool flip) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = q + e * TP;
            C u = {(T)0, (T)0};
            if (valid) {
                u.x = gin[(long long)r2r_pos(k, flip) * ns];
                u.y = k == 0 ? (T)0 : -gin[(long long)r2r_pos(2 * N - k, flip) * ns];
            }
            const C wc = {w4[k].x, -w4[k].y};
            smem[SI::at(p, k)] = cmul(u, wc);
        }
        if (q == 0) {
            C u = {(T)0, (T)0};
            if (valid) {
                u.x = gin[(long long)r2r_pos(N, flip) * ns];
                u.y = -u.x;
            }
            const C wc = {w4[N].x, -w4[N].y};
            smem[SI::tile_elems + p] = cmul(u, wc);
        }
    }
    // ---- r2r kinds I (FFTW_REDFT00 / RODFT00): the real DFT of the even / odd extension of length
    // L = 2N.  DCT-I of n = N + 1 points: e[j] = X[j] (j <= N), e[L-j] = X[j];  Y[k] = Re E[k], k = 0..N.
    // DST-I of n = N - 1 points: o[0] = o[N] = 0, o[j] = X[j-1], o[L-j] = -X[j-1];  Y[k-1] = -Im O[k], k = 1..N-1.
    static B2F_HD T r2r1_elem(const T* __restrict__ gin, long long ns, int j, bool sine) {
        if (!sine) return gin[(long long)(j <= N ? j : 2 * N - j) * ns];
        if (j == 0 || j == N) return (T)0;
        return j < N ? gin[(long long)(j - 1) * ns] : -gin[(long long)(2 * N - j - 1) * ns];
    }
    static B2F_HD void r2r1_load(C* v, int q, const T* __restrict__ gin, long long ns, bool valid, bool sine) {
        constexpr int R = RAD::get(0);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int m = q + b * TP + r * (N / R);
                C a = {(T)0, (T)0};
                if (valid) {
                    a.x = r2r1_elem(gin, ns, 2 * m, sine);
                    a.y = r2r1_elem(gin, ns, 2 * m + 1, sine);
                }
                v[b * R + r] = a;
            }
        }
    }
    static B2F_HD void r2r1_post(int p, int q, const C* smem, const C* __restrict__ w, T* __restrict__ gout,
                                 long long out_ns, bool valid, T scale, bool sine) {
        if (!valid) return;
        const T h = (T)0.5 * scale;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = q + e * TP;
            const C a = smem[SI::at(p, k)];
            C b = smem[SI::at(p, k == 0 ? 0 : N - k)];
            b.y = -b.y;
            const C sm = a + b, d = a - b;
            const C t = cmul(w[k], d);
            const C x = {(sm.x + t.y) * h, (sm.y - t.x) * h};          // spectrum of the extension at k (scaled)
            if (!sine) {
                gout[(long long)k * out_ns] = x.x;
                if (k == 0) gout[(long long)N * out_ns] = (a.x - a.y) * scale;
            } else if (k > 0) {
                gout[(long long)(k - 1) * out_ns] = -x.y;
            }
        }
    }
    // ---- r2r kinds IV (FFTW_REDFT11 / RODFT11) of even length n = 2N: ONE N-point complex transform,
    //   c[m] = (x[2m] + i x[n-1-2m]) exp(-i pi (4m+1) / 4n),  C = FFT_N(c),  d[k] = C[k] exp(-i pi k / n),
    //   Y[2k] = 2 Re d[k],  Y[n-1-2k] = -2 Im d[k];   the sine kind reverses the input and alternates the
    //   sign of the output.  t4 = build_dct4_twiddles.
    static B2F_HD void r2r4_load(C* v, int q, const T* __restrict__ gin, long long ns, bool valid, bool sine,
                                 const C* __restrict__ t4) {
        constexpr int R = RAD::get(0);
        constexpr int NB = E / R;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int m = q + b * TP + r * (N / R);
                C a = {(T)0, (T)0};
                if (valid) {
                    const T lo = gin[(long long)(2 * m) * ns], hi = gin[(long long)(2 * N - 1 - 2 * m) * ns];
                    a.x = sine ? hi : lo;
                    a.y = sine ? lo : hi;
                }
                v[b * R + r] = cmul(a, t4[m]);
            }
        }
    }
    static B2F_HD void r2r4_store(const C* v, int q, T* __restrict__ gout, long long out_ns, bool valid, T scale,
                                  bool sine, const C* __restrict__ t4) {
        constexpr int R = RAD::get(NPASS - 1);
        constexpr int NB = E / R;
        if (!valid) return;
        const T s2 = (T)2 * scale;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k = q + b * TP + r * (N / R);
                const C d = cmul(v[b * R + r], t4[N + k]);
                gout[(long long)(2 * k) * out_ns] = s2 * d.x;
                gout[(long long)(2 * N - 1 - 2 * k) * out_ns] = sine ? s2 * d.y : -s2 * d.y;
            }
        }
    }
    // kind III, last pass: (v.y, v.x) = (v[2m], v[2m+1]) of the c2r result -> y[sigma(j)] = v[j]
    static B2F_HD void r2r_store(const C* v, int q, T* __restrict__ gout, long long out_ns, bool valid, T scale,
                                 bool flip) {
        constexpr int R = RAD::get(NPASS - 1);
        constexpr int NB = E / R;
        if (!valid) return;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int m = q + b * TP + r * (N / R);
                const int i0 = r2r_sigma(2 * m), i1 = r2r_sigma(2 * m + 1);
                T y0 = v[b * R + r].y * scale, y1 = v[b * R + r].x * scale;
                if (flip) {
                    if (i0 & 1) y0 = -y0;
                    if (i1 & 1) y1 = -y1;
                }
                gout[(long long)i0 * out_ns] = y0;
                gout[(long long)i1 * out_ns] = y1;
            }
        }
    }
};

}  // namespace b2f
