"""Rotating kernels (csrc/fft_rot.cuh) stepped on the CPU: every row of
B2F_ROT_TABLE as a single step (in [B][I][O][n] -> out [B][O][n][I], ragged
last tile, forward / backward, fused scale), and the three-step schedule of a
3-axis stage (csrc/rot_plan.h, the library's own build_rotation) against
numpy.fft.fftn."""
import ctypes as C

import numpy as np
import pytest


def step(emu, prec, n, var, B, I, O, swap, scale, seed=0):
    rng = np.random.default_rng(seed)
    ct = np.complex128 if prec == 8 else np.complex64
    x = (rng.random((B, I, O, n)) + 1j * rng.random((B, I, O, n))).astype(ct)
    y = np.full((B, O, n, I), np.nan, dtype=ct)
    rc = emu.emu_rot_step(prec, n, var, B, I, O, x.ctypes.data, y.ctypes.data, C.c_double(scale), swap)
    if rc != 0:
        return None
    x64 = x.astype(np.complex128)
    ref = (np.fft.ifft(x64, axis=3) * n if swap else np.fft.fft(x64, axis=3)) * scale
    ref = ref.transpose(0, 2, 3, 1)
    return np.abs(y - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize('n', [64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize('prec', [8, 4])
def test_rot_step_variants(emu, n, prec):
    emu.emu_rot_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                 C.c_void_p, C.c_double, C.c_int]
    tol = 2e-15 if prec == 8 else 2e-6
    found = 0
    for var in range(8):
        for B, I, O in ((1, 19, 2), (2, 8, 1)) if n <= 512 else ((1, 11, 1),):
            for swap in (0, 1):
                err = step(emu, prec, n, var, B, I, O, swap, 1.0 / n if swap else 1.0)
                if err is None:
                    continue
                found += 1
                assert err < tol, (n, prec, var, B, I, O, swap, err)
    assert found >= 2


@pytest.mark.parametrize('shape,axes', [((64, 128, 64), (0, 1, 2)), ((2, 64, 64, 128), (3, 1, 2)),
                                         ((128, 64, 64), (2, 0, 1))])
@pytest.mark.parametrize('prec', [8, 4])
def test_rot_schedule_is_fftn(emu, shape, axes, prec):
    """in -> out -> scratch -> out: natural layout back after three rotations"""
    emu.emu_rot_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_double, C.c_int]
    ct = np.complex128 if prec == 8 else np.complex64
    tol = 1e-14 if prec == 8 else 1e-5
    rng = np.random.default_rng(1)
    x = (rng.random(shape) + 1j * rng.random(shape)).astype(ct)
    sizes = (C.c_longlong * len(shape))(*shape)
    ax = (C.c_int * 3)(*axes)
    # engine: -1 = bulk-copy staged rotating kernels, 1000 = the strided register-path kernels
    # with whole pencils in (FftParams::in_istride)
    for swap, var in ((0, -1), (1, -1), (0, 1000), (1, 1000)):
        y = np.full(shape, np.nan, dtype=ct)
        w = np.full(shape, np.nan, dtype=ct)
        xin = x.copy()
        scale = 0.5
        rc = emu.emu_rot_plan(prec, len(shape), sizes, ax, var, xin.ctypes.data, y.ctypes.data, w.ctypes.data,
                              C.c_double(scale), swap)
        assert rc == 0
        assert np.array_equal(xin, x)          # the input survives
        x64 = x.astype(np.complex128)
        nd = len(shape)
        tax = tuple(range(nd - 3, nd))
        ref = (np.fft.ifftn(x64, axes=tax) * np.prod(shape[-3:]) if swap else np.fft.fftn(x64, axes=tax)) * scale
        assert np.abs(y - ref).max() / np.abs(ref).max() < tol


def test_rot_schedule_refuses_other_axes(emu):
    emu.emu_rot_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_double, C.c_int]
    x = np.zeros((64, 64, 64, 64), dtype=np.complex128)
    sizes = (C.c_longlong * 4)(64, 64, 64, 64)
    assert emu.emu_rot_plan(8, 4, sizes, (C.c_int * 3)(0, 1, 2), -1, x.ctypes.data, x.ctypes.data, x.ctypes.data,
                            C.c_double(1.0), 0) == -1      # not the last three axes
    sizes = (C.c_longlong * 3)(64, 96, 64)
    assert emu.emu_rot_plan(8, 3, sizes, (C.c_int * 3)(0, 1, 2), -1, x.ctypes.data, x.ctypes.data, x.ctypes.data,
                            C.c_double(1.0), 0) == -1      # 96 has no rotating kernel
