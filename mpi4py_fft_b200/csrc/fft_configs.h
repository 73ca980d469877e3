// fft_configs.h -- the tables of power-of-two kernel instances.
//
// One row = one compiled kernel:  X(N, VAR, E, P, PS, MINB, radices...)
//   N      transform length
//   VAR    variant id (0 = default; others selectable with
//          b2f_set_option("variant_contig" | "variant_strided", v) -- used by
//          tools/sweep.py to pick the default on hardware)
//   E      points held per thread            (threads per pencil = N/E)
//   P      pencils per CTA tile              (threads per CTA = P*N/E)
//   PS     log2 of the shared-memory pad period (one pad slot every 2^PS
//          points; 30 = none)
//   MINB   __launch_bounds__ min CTAs per SM (register cap)
// CONTIG tables serve the unit-stride axis (a tile is P rows of N points),
// STRIDED tables every other axis (a tile is N rows of P contiguous elements;
// P*itemsize is the contiguous run per row: it must be >= 128 B to stay off the
// translation-rate limit when rows are more than a 2 MiB page apart, see
// DESIGN.md "what the stride probe showed").
// The same tables drive the kernels (fft_pow2_*.cu) and the CPU emulator
// (tests/emu/emu_fft.cpp).
#pragma once

#define B2F_CONTIG_SMALL(X)               \
    X(2, 0, 2, 128, 30, 1, 2) \
    X(4, 0, 4, 128, 30, 1, 4) \
    X(8, 0, 8, 64, 30, 1, 8) \
    X(16, 0, 16, 32, 30, 1, 16) \
    X(32, 0, 8, 32, 3, 1, 8, 4) \
    X(64, 0, 8, 16, 3, 1, 8, 8) \
    X(128, 0, 16, 16, 4, 1, 16, 8) \
    X(256, 0, 16, 8, 4, 1, 16, 16)

#define B2F_STRIDED_SMALL(X)              \
    X(2, 0, 2, 128, 30, 1, 2) \
    X(4, 0, 4, 128, 30, 1, 4) \
    X(8, 0, 8, 64, 30, 1, 8) \
    X(16, 0, 16, 32, 30, 1, 16) \
    X(32, 0, 8, 32, 30, 1, 8, 4) \
    X(64, 0, 8, 16, 30, 1, 8, 8) \
    X(128, 0, 16, 16, 30, 1, 16, 8) \
    X(256, 0, 16, 8, 30, 1, 16, 16) \
    X(256, 1, 16, 16, 30, 1, 16, 16)

#define B2F_CONTIG_MID(X)                 \
    X(512, 0, 16, 4, 3, 1, 8, 8, 8) \
    X(512, 6, 32, 8, 5, 1, 32, 16) \
    X(1024, 0, 16, 1, 4, 1, 16, 8, 8) \
    X(1024, 3, 32, 4, 5, 1, 32, 32)

#define B2F_STRIDED_MID(X)                \
    X(512, 0, 16, 8, 30, 1, 8, 8, 8) \
    X(512, 3, 16, 16, 30, 1, 8, 8, 8) \
    X(1024, 0, 16, 8, 30, 1, 16, 8, 8) \
    X(1024, 1, 16, 4, 4, 2, 16, 8, 8)

#define B2F_CONTIG_LARGE(X)               \
    X(2048, 0, 16, 2, 4, 1, 16, 16, 8)    \
    X(2048, 1, 16, 1, 4, 1, 16, 16, 8)    \
    X(4096, 0, 16, 1, 4, 1, 16, 16, 16)   \
    X(8192, 0, 16, 1, 4, 1, 16, 8, 8, 8)

#define B2F_STRIDED_LARGE(X)              \
    X(2048, 0, 16, 2, 4, 1, 16, 16, 8)    \
    X(2048, 1, 16, 4, 4, 1, 16, 16, 8)    \
    X(4096, 0, 16, 2, 4, 1, 16, 16, 16)   \
    X(8192, 0, 16, 1, 4, 1, 16, 8, 8, 8)

// lengths 3 * 2^k (the sizes a 3/2-rule padded solver produces): radices from
// {3, 4, 6, 8, 12, 24}, 24 points per thread above n = 24
#define B2F_CONTIG_MIXED(X)               \
    X(3, 0, 3, 128, 30, 1, 3)             \
    X(6, 0, 6, 64, 30, 1, 6)              \
    X(12, 0, 12, 64, 30, 1, 12)           \
    X(24, 0, 24, 64, 30, 1, 24)           \
    X(48, 0, 24, 32, 3, 1, 8, 6)          \
    X(96, 0, 24, 32, 3, 1, 12, 8)         \
    X(192, 0, 24, 16, 3, 1, 24, 8)        \
    X(384, 0, 24, 8, 3, 1, 6, 8, 8)       \
    X(768, 0, 24, 4, 3, 1, 12, 8, 8)      \
    X(1536, 0, 24, 2, 3, 1, 24, 8, 8)     \
    X(3072, 0, 24, 1, 3, 1, 6, 8, 8, 8)   \
    X(6144, 0, 24, 1, 3, 1, 24, 8, 8, 4)

#define B2F_STRIDED_MIXED(X)              \
    X(3, 0, 3, 128, 30, 1, 3)             \
    X(6, 0, 6, 64, 30, 1, 6)              \
    X(12, 0, 12, 64, 30, 1, 12)           \
    X(24, 0, 24, 64, 30, 1, 24)           \
    X(48, 0, 24, 32, 30, 1, 8, 6)         \
    X(96, 0, 24, 32, 30, 1, 12, 8)        \
    X(192, 0, 24, 16, 30, 1, 24, 8)       \
    X(384, 0, 24, 8, 30, 1, 6, 8, 8)      \
    X(768, 0, 24, 8, 30, 1, 12, 8, 8)     \
    X(1536, 0, 24, 8, 30, 1, 24, 8, 8)    \
    X(3072, 0, 24, 4, 30, 1, 6, 8, 8, 8)  \
    X(6144, 0, 24, 2, 30, 1, 24, 8, 8, 4)

// lengths 5 * 2^k (<= 1280) and 7 * 2^k (<= 1792): radices {5, 10, 20} with 20 points per thread,
// {7, 14, 28} with 28 points per thread, then radix-4 / radix-2 passes (every radix of a schedule
// must divide E).  What FFTW plans for any N through its codelets
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:49-76) these lengths get as Stockham kernels
// instead of the chirp-z detour.
#define B2F_CONTIG_MIXED5(X) \
    X(5, 0, 5, 128, 30, 1, 5) \
    X(10, 0, 10, 64, 30, 1, 10) \
    X(20, 0, 20, 64, 30, 1, 20) \
    X(40, 0, 20, 32, 3, 1, 20, 2) \
    X(80, 0, 20, 32, 3, 1, 20, 4) \
    X(160, 0, 20, 16, 3, 1, 20, 4, 2) \
    X(320, 0, 20, 8, 3, 1, 20, 4, 4) \
    X(640, 0, 20, 4, 3, 1, 20, 4, 4, 2) \
    X(1280, 0, 20, 2, 3, 1, 20, 4, 4, 4)

#define B2F_CONTIG_MIXED7(X) \
    X(7, 0, 7, 128, 30, 1, 7) \
    X(14, 0, 14, 64, 30, 1, 14) \
    X(28, 0, 28, 32, 30, 1, 28) \
    X(56, 0, 28, 32, 3, 1, 28, 2) \
    X(112, 0, 28, 32, 3, 1, 28, 4) \
    X(224, 0, 28, 16, 3, 1, 28, 4, 2) \
    X(448, 0, 28, 8, 3, 1, 28, 4, 4) \
    X(896, 0, 28, 4, 3, 1, 28, 4, 4, 2) \
    X(1792, 0, 28, 2, 3, 1, 28, 4, 4, 4)

#define B2F_CONTIG_MIXED57(X) B2F_CONTIG_MIXED5(X) B2F_CONTIG_MIXED7(X)

#define B2F_STRIDED_MIXED5(X) \
    X(5, 0, 5, 128, 30, 1, 5) \
    X(10, 0, 10, 64, 30, 1, 10) \
    X(20, 0, 20, 64, 30, 1, 20) \
    X(40, 0, 20, 32, 30, 1, 20, 2) \
    X(80, 0, 20, 32, 30, 1, 20, 4) \
    X(160, 0, 20, 16, 30, 1, 20, 4, 2) \
    X(320, 0, 20, 8, 30, 1, 20, 4, 4) \
    X(640, 0, 20, 8, 30, 1, 20, 4, 4, 2) \
    X(1280, 0, 20, 8, 30, 1, 20, 4, 4, 4)

#define B2F_STRIDED_MIXED7(X) \
    X(7, 0, 7, 128, 30, 1, 7) \
    X(14, 0, 14, 64, 30, 1, 14) \
    X(28, 0, 28, 32, 30, 1, 28) \
    X(56, 0, 28, 32, 30, 1, 28, 2) \
    X(112, 0, 28, 32, 30, 1, 28, 4) \
    X(224, 0, 28, 16, 30, 1, 28, 4, 2) \
    X(448, 0, 28, 8, 30, 1, 28, 4, 4) \
    X(896, 0, 28, 8, 30, 1, 28, 4, 4, 2) \
    X(1792, 0, 28, 4, 30, 1, 28, 4, 4, 4)

#define B2F_STRIDED_MIXED57(X) B2F_STRIDED_MIXED5(X) B2F_STRIDED_MIXED7(X)

// TMA-staged strided kernels (fft_tma.cuh):
//   X(N, VAR, E, P, PS, STAGES, SPLIT, MINB, radices...)
//   STAGES  shared-memory stages the TMA engine fills ahead of the compute
//   SPLIT   1 = the exchange buffer holds one real component at a time
//   PS      pad period of the exchange buffer (rows narrower than 128 B need
//           log2(first radix); 30 = none)
//   MINB    min CTAs per SM  +  16 * OPT   (OPT bit 0: pass twiddles live in shared
//           memory; bit 1: cp.async requests of the next tile are spread over the
//           current tile's phases; bit 2: L2 prefetch of the tile after next; bit 3 (cp.async table only): results leave through
//           shared memory and TMA tensor stores) -- fft_tma.cuh
#define B2F_TMA_TABLE_A(X) \
    X(64, 0, 8, 16, 30, 2, 0, 1, 8, 8) \
    X(128, 0, 16, 16, 30, 2, 0, 1, 16, 8) \
    X(128, 1, 16, 8, 30, 2, 0, 1, 16, 8) \
    X(256, 0, 16, 8, 30, 2, 0, 1, 16, 16)

#define B2F_TMA_TABLE_B(X) \
    X(512, 0, 16, 8, 3, 1, 1, 2, 8, 8, 8) \
    X(512, 2, 16, 8, 30, 2, 0, 1, 8, 8, 8)

#define B2F_TMA_TABLE_C(X) \
    X(1024, 0, 32, 8, 5, 1, 1, 1, 32, 32) \
    X(1024, 6, 32, 8, 5, 1, 1, 17, 32, 32) \
    X(2048, 0, 16, 4, 4, 1, 1, 1, 16, 16, 8)

#define B2F_TMA_TABLE(X) B2F_TMA_TABLE_A(X) B2F_TMA_TABLE_B(X) B2F_TMA_TABLE_C(X)

// the same pipeline filled by cp.async instead of TMA (variant_tma = 100 + VAR)
#define B2F_CPA_TABLE_A(X) \
    X(256, 2, 16, 16, 30, 2, 0, 17, 16, 16) \
    X(256, 3, 16, 16, 30, 2, 0, 81, 16, 16) \
    X(192, 0, 24, 16, 30, 2, 0, 17, 24, 8) \
    X(384, 1, 24, 16, 30, 1, 0, 17, 6, 8, 8)

#define B2F_CPA_TABLE_B(X) \
    X(512, 0, 16, 8, 30, 2, 0, 1, 8, 8, 8) \
    X(512, 5, 32, 16, 5, 1, 1, 17, 32, 16) \
    X(768, 0, 24, 8, 30, 1, 0, 17, 12, 8, 8)

#define B2F_CPA_TABLE_C(X) \
    X(1024, 0, 32, 8, 5, 1, 1, 1, 32, 32) \
    X(1024, 2, 32, 8, 5, 1, 1, 17, 32, 32) \
    X(1024, 7, 32, 8, 5, 1, 1, 145, 32, 32) \
    X(2048, 0, 16, 4, 4, 1, 1, 1, 16, 16, 8) \
    X(2048, 2, 32, 4, 5, 1, 1, 1, 32, 8, 8)

#define B2F_CPA_TABLE(X) B2F_CPA_TABLE_A(X) B2F_CPA_TABLE_B(X) B2F_CPA_TABLE_C(X)

// rotating kernels (fft_rot.cuh): contiguous pencils in (bulk copies), rows of P
// elements out with the axes rotated.  X(N, VAR, E, P, PS, STAGES, SPLIT, MINB + 16*OPT, radices...)
// as B2F_TMA_TABLE, plus OPT bit 3 (MINB + 128): results leave through shared memory and TMA
// stores, OPT bit 2 (MINB + 64): cp.async loader instead of bulk copies, OPT bits 4-5 (MINB + 256, + 512): 2 or 4 independent warp groups per CTA, each with
// its own tile of P pencils; P * 16 bytes is the contiguous run of a STORE only (loads are
// whole pencils whatever P is), so narrow tiles with two stages or two CTAs per SM
// are candidates here.
#define B2F_ROT_TABLE_A(X) \
    X(64, 0, 8, 16, 30, 2, 0, 20, 8, 8) \
    X(128, 0, 16, 16, 30, 2, 0, 18, 16, 8) \
    X(256, 0, 16, 8, 30, 2, 0, 18, 16, 16) \
    X(512, 0, 16, 8, 30, 2, 0, 17, 8, 8, 8) \
    X(512, 1, 32, 16, 5, 1, 1, 17, 32, 16)

#define B2F_ROT_TABLE_B(X) \
    X(1024, 0, 32, 8, 5, 1, 1, 17, 32, 32) \
    X(1024, 2, 32, 8, 5, 1, 1, 145, 32, 32) \
    X(1024, 3, 32, 4, 5, 1, 1, 273, 32, 32) \
    X(2048, 0, 16, 4, 4, 1, 1, 1, 16, 16, 8)

#define B2F_ROT_TABLE(X) B2F_ROT_TABLE_A(X) B2F_ROT_TABLE_B(X)

// real transforms (r2c / c2r of even length 2N through the N-point schedule,
// fft_pow2.cuh fft_real_kernel):  X(N, E, P, PS, MINB, radices...), one row per N
#define B2F_REAL_CONTIG_POW2(X)          \
    X(2, 2, 128, 30, 1, 2)               \
    X(4, 4, 128, 30, 1, 4)               \
    X(8, 8, 64, 30, 1, 8)                \
    X(16, 16, 32, 30, 1, 16)             \
    X(32, 8, 32, 3, 1, 8, 4)             \
    X(64, 8, 16, 3, 1, 8, 8)             \
    X(128, 16, 16, 4, 1, 16, 8)          \
    X(256, 16, 8, 4, 1, 16, 16)          \
    X(512, 16, 4, 3, 1, 8, 8, 8)         \
    X(1024, 16, 2, 4, 1, 16, 8, 8)       \
    X(2048, 16, 2, 4, 1, 16, 16, 8)      \
    X(4096, 16, 1, 4, 1, 16, 16, 16)     \
    X(8192, 16, 1, 4, 1, 16, 8, 8, 8)

#define B2F_REAL_CONTIG_MIXED(X)         \
    X(3, 3, 128, 30, 1, 3)               \
    X(6, 6, 64, 30, 1, 6)                \
    X(12, 12, 64, 30, 1, 12)             \
    X(24, 24, 64, 30, 1, 24)             \
    X(48, 24, 32, 3, 1, 8, 6)            \
    X(96, 24, 32, 3, 1, 12, 8)           \
    X(192, 24, 16, 3, 1, 24, 8)          \
    X(384, 24, 8, 3, 1, 6, 8, 8)         \
    X(768, 24, 4, 3, 1, 12, 8, 8)        \
    X(1536, 24, 2, 3, 1, 24, 8, 8)       \
    X(3072, 24, 1, 3, 1, 6, 8, 8, 8)     \
    X(6144, 24, 1, 3, 1, 24, 8, 8, 4)

#define B2F_REAL_STRIDED_POW2(X)         \
    X(2, 2, 128, 30, 1, 2)               \
    X(4, 4, 128, 30, 1, 4)               \
    X(8, 8, 64, 30, 1, 8)                \
    X(16, 16, 32, 30, 1, 16)             \
    X(32, 8, 32, 30, 1, 8, 4)            \
    X(64, 8, 16, 30, 1, 8, 8)            \
    X(128, 16, 16, 30, 1, 16, 8)         \
    X(256, 16, 8, 30, 1, 16, 16)         \
    X(512, 16, 8, 30, 1, 8, 8, 8)        \
    X(1024, 16, 8, 30, 1, 16, 8, 8)      \
    X(2048, 16, 2, 4, 1, 16, 16, 8)      \
    X(4096, 16, 2, 4, 1, 16, 16, 16)     \
    X(8192, 16, 1, 4, 1, 16, 8, 8, 8)

#define B2F_REAL_STRIDED_MIXED(X)        \
    X(3, 3, 128, 30, 1, 3)               \
    X(6, 6, 64, 30, 1, 6)                \
    X(12, 12, 64, 30, 1, 12)             \
    X(24, 24, 64, 30, 1, 24)             \
    X(48, 24, 32, 30, 1, 8, 6)           \
    X(96, 24, 32, 30, 1, 12, 8)          \
    X(192, 24, 16, 30, 1, 24, 8)         \
    X(384, 24, 8, 30, 1, 6, 8, 8)        \
    X(768, 24, 8, 30, 1, 12, 8, 8)       \
    X(1536, 24, 8, 30, 1, 24, 8, 8)      \
    X(3072, 24, 4, 30, 1, 6, 8, 8, 8)    \
    X(6144, 24, 2, 30, 1, 24, 8, 8, 4)

// real transforms of length 2N, N = 5 * 2^k or 7 * 2^k (rows of B2F_*_MIXED57 without the VAR column)
#define B2F_REAL_CONTIG_MIXED5(X) \
    X(5, 5, 128, 30, 1, 5) \
    X(10, 10, 64, 30, 1, 10) \
    X(20, 20, 64, 30, 1, 20) \
    X(40, 20, 32, 3, 1, 20, 2) \
    X(80, 20, 32, 3, 1, 20, 4) \
    X(160, 20, 16, 3, 1, 20, 4, 2) \
    X(320, 20, 8, 3, 1, 20, 4, 4) \
    X(640, 20, 4, 3, 1, 20, 4, 4, 2) \
    X(1280, 20, 2, 3, 1, 20, 4, 4, 4)

#define B2F_REAL_CONTIG_MIXED7(X) \
    X(7, 7, 128, 30, 1, 7) \
    X(14, 14, 64, 30, 1, 14) \
    X(28, 28, 32, 30, 1, 28) \
    X(56, 28, 32, 3, 1, 28, 2) \
    X(112, 28, 32, 3, 1, 28, 4) \
    X(224, 28, 16, 3, 1, 28, 4, 2) \
    X(448, 28, 8, 3, 1, 28, 4, 4) \
    X(896, 28, 4, 3, 1, 28, 4, 4, 2) \
    X(1792, 28, 2, 3, 1, 28, 4, 4, 4)

#define B2F_REAL_CONTIG_MIXED57(X) B2F_REAL_CONTIG_MIXED5(X) B2F_REAL_CONTIG_MIXED7(X)

#define B2F_REAL_STRIDED_MIXED5(X) \
    X(5, 5, 128, 30, 1, 5) \
    X(10, 10, 64, 30, 1, 10) \
    X(20, 20, 64, 30, 1, 20) \
    X(40, 20, 32, 30, 1, 20, 2) \
    X(80, 20, 32, 30, 1, 20, 4) \
    X(160, 20, 16, 30, 1, 20, 4, 2) \
    X(320, 20, 8, 30, 1, 20, 4, 4) \
    X(640, 20, 8, 30, 1, 20, 4, 4, 2) \
    X(1280, 20, 4, 30, 1, 20, 4, 4, 4)

#define B2F_REAL_STRIDED_MIXED7(X) \
    X(7, 7, 128, 30, 1, 7) \
    X(14, 14, 64, 30, 1, 14) \
    X(28, 28, 32, 30, 1, 28) \
    X(56, 28, 32, 30, 1, 28, 2) \
    X(112, 28, 32, 30, 1, 28, 4) \
    X(224, 28, 16, 30, 1, 28, 4, 2) \
    X(448, 28, 8, 30, 1, 28, 4, 4) \
    X(896, 28, 8, 30, 1, 28, 4, 4, 2) \
    X(1792, 28, 4, 30, 1, 28, 4, 4, 4)

#define B2F_REAL_STRIDED_MIXED57(X) B2F_REAL_STRIDED_MIXED5(X) B2F_REAL_STRIDED_MIXED7(X)

#define B2F_REAL_CONTIG(X) B2F_REAL_CONTIG_POW2(X) B2F_REAL_CONTIG_MIXED(X) B2F_REAL_CONTIG_MIXED57(X)
#define B2F_REAL_STRIDED(X) B2F_REAL_STRIDED_POW2(X) B2F_REAL_STRIDED_MIXED(X) B2F_REAL_STRIDED_MIXED57(X)
#define B2F_CONTIG_ALL(X) B2F_CONTIG_SMALL(X) B2F_CONTIG_MID(X) B2F_CONTIG_LARGE(X) B2F_CONTIG_MIXED(X) B2F_CONTIG_MIXED57(X)
#define B2F_STRIDED_ALL(X) B2F_STRIDED_SMALL(X) B2F_STRIDED_MID(X) B2F_STRIDED_LARGE(X) B2F_STRIDED_MIXED(X) B2F_STRIDED_MIXED57(X)

#define B2F_POW2_MAX_N 8192
#define B2F_MIXED_MAX_N 6144
#define B2F_MIXED5_MAX_N 1280
#define B2F_MIXED7_MAX_N 1792
