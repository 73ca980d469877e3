"""DCT / DST of kinds II and III as Stockham real transforms (csrc/fft_core.cuh r2r_load / r2r_post /
r2r_fill / r2r_store: Makhoul's permutation + the n/2-point complex schedule + a quarter-wave twiddle)
stepped on the CPU and compared with scipy's definitions == FFTW's REDFT10 / REDFT01 / RODFT10 /
RODFT01 (/root/reference/mpi4py_fft/fftw/xfftn.py:14-36): every Stockham family, unit-stride and
strided, ragged tiles, in place, fused scale."""
import ctypes as C

import numpy as np
import pytest
import scipy.fft as sfft

KINDS = {5: ('dct', 2), 4: ('dct', 3), 9: ('dst', 2), 8: ('dst', 3), 6: ('dct', 4), 10: ('dst', 4)}


@pytest.mark.parametrize('n', [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 6, 12, 24, 48, 96, 192, 384, 768,
                               10, 20, 40, 80, 160, 14, 28, 56, 112])
@pytest.mark.parametrize('prec', [8, 4])
def test_r2r_kinds_2_3_4(emu, n, prec):
    emu.emu_fft_r2r.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_double]
    rt = np.float64 if prec == 8 else np.float32
    tol = 4e-15 if prec == 8 else 4e-6
    rng = np.random.default_rng(n)
    for outer, inner in ((3, 1), (2, 5), (1, 17)) if n <= 512 else ((2, 1), (1, 3)):
        x = rng.random((outer, n, inner)).astype(rt)
        for kind, (fam, typ) in KINDS.items():
            ref = getattr(sfft, fam)(x.astype(np.float64), type=typ, axis=1) * 0.25
            y = np.full_like(x, np.nan)
            assert emu.emu_fft_r2r(prec, kind, n, outer, inner, x.ctypes.data, y.ctypes.data, C.c_double(0.25)) == 0
            assert np.abs(y - ref).max() <= tol * np.abs(ref).max() * max(1, np.log2(n)), (n, prec, kind, outer, inner)
            # in place
            z = x.copy()
            assert emu.emu_fft_r2r(prec, kind, n, outer, inner, z.ctypes.data, z.ctypes.data, C.c_double(0.25)) == 0
            assert np.array_equal(z, y), (n, prec, kind, 'in place')


@pytest.mark.parametrize('N', [2, 4, 8, 16, 64, 256, 1024, 12, 96, 40, 56])
@pytest.mark.parametrize('prec', [8, 4])
def test_r2r_kinds_1(emu, N, prec):
    """DCT-I of N + 1 points (Chebyshev grids 2^k + 1) and DST-I of N - 1 points: the real transform of the
    even / odd extension of length 2N, read through an index map (r2r1_load / r2r1_post)"""
    emu.emu_fft_r2r.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_double]
    rt = np.float64 if prec == 8 else np.float32
    tol = 4e-15 if prec == 8 else 4e-6
    rng = np.random.default_rng(N)
    for kind, fam, n in ((3, 'dct', N + 1), (7, 'dst', N - 1)):
        if n < 2:
            continue
        for outer, inner in ((3, 1), (2, 5)):
            x = rng.random((outer, n, inner)).astype(rt)
            ref = getattr(sfft, fam)(x.astype(np.float64), type=1, axis=1) * 0.5
            y = np.full_like(x, np.nan)
            assert emu.emu_fft_r2r(prec, kind, n, outer, inner, x.ctypes.data, y.ctypes.data, C.c_double(0.5)) == 0
            assert np.abs(y - ref).max() <= tol * np.abs(ref).max() * max(1, np.log2(N)), (N, prec, kind, outer, inner)
            z = x.copy()
            assert emu.emu_fft_r2r(prec, kind, n, outer, inner, z.ctypes.data, z.ctypes.data, C.c_double(0.5)) == 0
            assert np.array_equal(z, y), (N, prec, kind, 'in place')
