"""Four-step split of c2c lengths beyond one tile (csrc/lengths.h, capi.cu STEP_FOURSTEP) stepped on the CPU:
the library's own split rule, the strided / rotating kernels' per-thread code for the two transforms, the twiddle
in between -- against numpy.fft for unit-stride and strided axes, forward and backward."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize('n,expect', [(16384, (128, 128)), (32768, (128, 256)), (12288, (128, 96)), (10240, (128, 80)),
                                      (65536, (256, 256)), (7168, (64, 112))])
@pytest.mark.parametrize('prec', [8, 4])
def test_fourstep_is_fft(emu, n, expect, prec):
    emu.emu_fourstep.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_double, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    ct = np.complex128 if prec == 8 else np.complex64
    tol = 5e-15 if prec == 8 else 5e-6
    rng = np.random.default_rng(n % 1000)
    for outer, inner in ((2, 1), (1, 3)) if n <= 16384 else ((1, 1),):
        x = (rng.random((outer, n, inner)) + 1j * rng.random((outer, n, inner))).astype(ct)
        for swap in (0, 1):
            y = np.full_like(x, np.nan)
            w = np.full_like(x, np.nan)
            n1, n2 = C.c_longlong(), C.c_longlong()
            xin = x.copy()
            rc = emu.emu_fourstep(prec, n, outer, inner, xin.ctypes.data, y.ctypes.data, w.ctypes.data, C.c_double(0.5), swap,
                                  C.byref(n1), C.byref(n2))
            assert rc == 0
            assert (n1.value, n2.value) == expect
            assert np.array_equal(xin, x)
            x64 = x.astype(np.complex128)
            ref = (np.fft.ifft(x64, axis=1) * n if swap else np.fft.fft(x64, axis=1)) * 0.5
            assert np.abs(y - ref).max() <= tol * np.abs(ref).max() * np.log2(n), (n, prec, outer, inner, swap)


def test_fourstep_leaves_other_lengths_alone(emu):
    emu.emu_fourstep.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_double, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    x = np.zeros(8, dtype=np.complex128)
    a, b = C.c_longlong(), C.c_longlong()
    for n in (8192, 1024, 6144, 8191, 64 * 13 * 11):      # kernels of their own / no admissible pair
        assert emu.emu_fourstep(8, n, 1, 1, x.ctypes.data, x.ctypes.data, x.ctypes.data, C.c_double(1.0), 0,
                                C.byref(a), C.byref(b)) == -1
