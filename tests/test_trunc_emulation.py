"""Dealiasing folded into the transform (csrc/fft_core.cuh TruncMap: truncating last pass of the
forward transform, padding first pass of the backward one) stepped on the CPU by tests/emu and
compared with the oracle's restatement of the reference rule (oracle/pfft_oracle.py
truncate_forward / pad_backward <- reference libfft.py:263-311): c2c and r2c / c2r, even and odd
kept extents (the Nyquist rule differs), unit-stride and strided layouts, 3/2-rule sizes."""
import ctypes as C

import numpy as np
import pytest

import pfft_oracle as O


def _emu(emu):
    emu.emu_fft_trunc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p,
                                  C.c_void_p, C.c_double]
    return emu


CASES = [(96, 64), (96, 63), (192, 128), (384, 256), (384, 255), (768, 512), (48, 32), (12, 8), (24, 17)]


@pytest.mark.parametrize('n_pad,n_keep', CASES)
@pytest.mark.parametrize('prec', [8, 4])
@pytest.mark.parametrize('outer,inner', [(3, 1), (2, 5)])
def test_c2c_truncating_store_and_padding_load(emu, n_pad, n_keep, prec, outer, inner):
    emu = _emu(emu)
    ct = np.complex128 if prec == 8 else np.complex64
    tol = 5e-15 if prec == 8 else 5e-6
    rng = np.random.default_rng(n_pad + n_keep)
    x = (rng.random((outer, n_pad, inner)) + 1j * rng.random((outer, n_pad, inner))).astype(ct)
    y = np.full((outer, n_keep, inner), np.nan, dtype=ct)
    scale = 1.0 / n_pad
    assert emu.emu_fft_trunc(prec, -1, n_pad, n_keep, outer, inner, x.ctypes.data, y.ctypes.data, C.c_double(scale)) == 0
    ref = O.truncate_forward(np.fft.fft(x.astype(np.complex128), axis=1), 1, n_keep, False) * scale
    assert np.abs(y - ref).max() <= tol * max(1.0, np.abs(ref).max()), (n_pad, n_keep)
    # backward: truncated spectrum in, padded physical block out (unnormalised)
    t = (rng.random((outer, n_keep, inner)) + 1j * rng.random((outer, n_keep, inner))).astype(ct)
    z = np.full((outer, n_pad, inner), np.nan, dtype=ct)
    assert emu.emu_fft_trunc(prec, 1, n_pad, n_keep, outer, inner, t.ctypes.data, z.ctypes.data, C.c_double(1.0)) == 0
    refb = np.fft.ifft(O.pad_backward(t.astype(np.complex128), 1, n_pad, False), axis=1) * n_pad
    assert np.abs(z - refb).max() <= tol * np.abs(refb).max(), (n_pad, n_keep)


@pytest.mark.parametrize('n_pad,n_keep', [(96, 64), (192, 128), (384, 256), (768, 512), (96, 62), (24, 16), (12, 8),
                                           (128, 86), (64, 42)])
@pytest.mark.parametrize('prec', [8, 4])
@pytest.mark.parametrize('outer,inner', [(3, 1), (2, 5)])
def test_real_transforms_with_truncated_half_spectrum(emu, n_pad, n_keep, prec, outer, inner):
    """n_keep = logical kept length; the half spectrum keeps n_keep // 2 + 1 modes (libfft.py:401-406)"""
    emu = _emu(emu)
    rt, ct = (np.float64, np.complex128) if prec == 8 else (np.float32, np.complex64)
    tol = 5e-15 if prec == 8 else 5e-6
    keep = n_keep // 2 + 1
    rng = np.random.default_rng(n_pad * 3 + n_keep)
    x = rng.random((outer, n_pad, inner)).astype(rt)
    y = np.full((outer, keep, inner), np.nan, dtype=ct)
    scale = 1.0 / n_pad
    assert emu.emu_fft_trunc(prec, -2, n_pad, keep, outer, inner, x.ctypes.data, y.ctypes.data, C.c_double(scale)) == 0
    ref = O.truncate_forward(np.fft.rfft(x.astype(np.float64), axis=1), 1, keep, True) * scale
    assert np.abs(y - ref).max() <= tol * max(1.0, np.abs(ref).max()), (n_pad, n_keep)
    t = (rng.random((outer, keep, inner)) + 1j * rng.random((outer, keep, inner))).astype(ct)
    z = np.full((outer, n_pad, inner), np.nan, dtype=rt)
    assert emu.emu_fft_trunc(prec, 2, n_pad, keep, outer, inner, t.ctypes.data, z.ctypes.data, C.c_double(1.0)) == 0
    refb = np.fft.irfft(O.pad_backward(t.astype(np.complex128), 1, n_pad // 2 + 1, True), n=n_pad, axis=1) * n_pad
    assert np.abs(z - refb).max() <= tol * np.abs(refb).max(), (n_pad, n_keep)
