// fft_core.cuh -- per-thread building blocks of the sm_100a Stockham kernels.
//
// Everything here is __host__ __device__ and free of CUDA built-ins so that
// tests/emu/emu_fft.cpp can step the exact per-thread code of a kernel on the
// CPU (thread by thread, barrier phase by barrier phase) before a GPU is ever
// involved.  Replaces the FFTW codelets behind fftw_plan_guru_dft
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:52-56).
#pragma once
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

// ---------------------------------------------------------------------------
// R-point forward DFT on registers, natural order in and out (R = 2..32).
// Decimation in time by two; all indices are compile-time constants after
// unrolling so v/e/o live in registers.
// ---------------------------------------------------------------------------
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
            cplx<T> t = mul_w32(o[k], k * (32 / R));
            v[k] = e[k] + t;
            v[k + R / 2] = e[k] - t;
        }
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
    using SI = SmemIndex<STRIDED, P, N, PS>;
    static constexpr int TP = N / E;           // threads per pencil
    static constexpr int THREADS = TP * P;
    static constexpr int NPASS = RAD::count;
    static_assert(RAD::product() == N, "radix schedule must multiply to N");
    static_assert(N % E == 0, "E must divide N");

    // thread -> (pencil in tile, slot in pencil)
    static B2F_HD int pencil_of(int tid) { return STRIDED ? tid % P : tid / TP; }
    static B2F_HD int slot_of(int tid) { return STRIDED ? tid / P : tid % TP; }

    template <int S>
    static B2F_HD void twiddle_dft(C* v, int q, const C* __restrict__ tw) {
        constexpr int R = RAD::get(S);
        constexpr int Ns = RAD::before(S);
        constexpr int NB = E / R;
        static_assert(E % R == 0, "radix must divide E");
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (S > 0) {
                const int j = q + b * TP;
                const int k = j % Ns;
                constexpr int tstride = N / (Ns * R);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const C w = tw[r * k * tstride];
                    v[b * R + r] = cmul(v[b * R + r], w);
                }
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
};

}  // namespace b2f
