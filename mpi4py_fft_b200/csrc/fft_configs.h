// fft_configs.h -- the one table of power-of-two kernel instances.
//
// X(N, VAR, E, PC, PSC, PST, PSS, MINB, radices...)
//   N      transform length
//   VAR    variant id (0 = default; others selectable with b2f_set_option("variant", v))
//   E      points held per thread            (threads per pencil = N/E)
//   PC     pencils per CTA, contiguous axis  PSC  log2 pad period there (30 = none)
//   PST    pencils per CTA, strided axis     PSS  log2 pad period there (30 = none)
//   MINB   __launch_bounds__ min CTAs per SM (register cap)
// The same table drives the kernels (fft_pow2_*.cu) and the CPU emulator
// (tests/emu/emu_fft.cpp).
#pragma once

#define B2F_POW2_TABLE_SMALL(X)                    \
    X(2, 0, 2, 128, 30, 128, 30, 1, 2)             \
    X(4, 0, 4, 128, 30, 128, 30, 1, 4)             \
    X(8, 0, 8, 64, 30, 64, 30, 1, 8)               \
    X(16, 0, 16, 32, 30, 32, 30, 1, 16)            \
    X(32, 0, 8, 32, 3, 32, 30, 1, 8, 4)            \
    X(64, 0, 8, 16, 3, 16, 30, 1, 8, 8)            \
    X(128, 0, 16, 16, 4, 16, 30, 1, 16, 8)         \
    X(256, 0, 16, 8, 4, 8, 30, 1, 16, 16)

#define B2F_POW2_TABLE_MID(X)                      \
    X(512, 0, 8, 4, 3, 4, 3, 2, 8, 8, 8)           \
    X(512, 1, 16, 4, 3, 8, 30, 1, 8, 8, 8)         \
    X(512, 2, 8, 2, 3, 8, 30, 1, 8, 8, 8)          \
    X(1024, 0, 16, 4, 4, 4, 4, 2, 16, 8, 8)        \
    X(1024, 1, 16, 2, 4, 8, 30, 1, 16, 8, 8)       \
    X(1024, 2, 16, 1, 4, 2, 4, 1, 16, 8, 8)

#define B2F_POW2_TABLE_LARGE(X)                    \
    X(2048, 0, 16, 2, 4, 2, 4, 1, 16, 16, 8)       \
    X(2048, 1, 16, 1, 4, 4, 4, 1, 16, 16, 8)       \
    X(4096, 0, 16, 1, 4, 2, 4, 1, 16, 16, 16)      \
    X(8192, 0, 16, 1, 4, 1, 4, 1, 16, 8, 8, 8)

#define B2F_POW2_TABLE_ALL(X) \
    B2F_POW2_TABLE_SMALL(X) B2F_POW2_TABLE_MID(X) B2F_POW2_TABLE_LARGE(X)

#define B2F_POW2_MAX_N 8192
