// lengths.h -- which transform lengths have a Stockham kernel instance, and how a longer length is
// split into two that do (four-step).  Pure host code shared by capi.cu and the CPU emulator.
#pragma once
#include "fft_configs.h"
#include "rot_plan.h"

namespace b2f {

static inline bool is_pow2(long long n) { return n >= 2 && (n & (n - 1)) == 0; }
// 3 * 2^k, 3 <= n <= 6144: served by the radix-3/6/12/24 schedules
static inline bool is_mixed(long long n) { return n >= 3 && n <= B2F_MIXED_MAX_N && n % 3 == 0 && (n == 3 || is_pow2(n / 3)); }
// 5 * 2^k <= 1280 and 7 * 2^k <= 1792: the radix-5/10/20 and radix-7/14/28 schedules
static inline bool is_mixed57(long long n) {
    if (n >= 5 && n <= B2F_MIXED5_MAX_N && n % 5 == 0 && (n == 5 || is_pow2(n / 5))) return true;
    return n >= 7 && n <= B2F_MIXED7_MAX_N && n % 7 == 0 && (n == 7 || is_pow2(n / 7));
}
// lengths with a Stockham kernel instance
static inline bool is_stockham(long long n) { return (is_pow2(n) && n <= B2F_POW2_MAX_N) || is_mixed(n) || is_mixed57(n); }

// Four-step split of a c2c length beyond one tile: n = n1 * n2 with n1 a length of the rotating kernels
// (2^k, 64..2048: the second step reads whole n1-point rows and writes them transposed) and n2 any
// Stockham length.  With j = j1 + n1*j2 and k = n2*k1 + k2:
//   X[n2*k1 + k2] = sum_j1 W_n1^(j1 k1) * W_n^(j1 k2) * [ sum_j2 x[j1 + n1*j2] W_n2^(j2 k2) ]
//   step 1: n2-point transforms along j2 (stride n1) for every j1;  step 2: twiddle W_n^(j1 k2);
//   step 3: n1-point transforms along j1 for every k2, stored transposed (k1-major).
// The most balanced admissible pair is taken.  Replaces what FFTW's planner does for any N
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:49-56).
struct FourStep {
    long long n, n1, n2;
};
static inline bool fourstep_split(long long n, FourStep* fs) {
    long long best1 = 0, best_gap = -1;
    for (long long n1 = 64; n1 <= 2048; n1 *= 2) {
        if (rot_default((int)n1) < 0 || n % n1) continue;
        const long long n2 = n / n1;
        if (n2 < 2 || !is_stockham(n2)) continue;
        const long long gap = n1 > n2 ? n1 - n2 : n2 - n1;
        if (best_gap < 0 || gap < best_gap) {
            best_gap = gap;
            best1 = n1;
        }
    }
    if (!best1) return false;
    fs->n = n;
    fs->n1 = best1;
    fs->n2 = n / best1;
    return true;
}

}  // namespace b2f
