"""Every row of csrc/fft_configs.h stepped on the CPU (tests/emu/emu_fft.cpp runs
the kernels' own per-thread phase functions from csrc/fft_core.cuh) against
numpy: contiguous and strided layouts, ragged tiles, forward and backward
(swap trick), fused scale, in place."""
import numpy as np
import pytest

SIZES = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192,
         3, 6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 6144,        # 2^k and 3 * 2^k
         5, 10, 20, 40, 80, 160, 320, 640, 1280, 7, 14, 28, 56, 112, 224, 448, 896, 1792]   # 5 * 2^k, 7 * 2^k


def run(emu, prec, n, var, outer, inner, swap, scale, inplace, seed=0):
    rng = np.random.default_rng(seed)
    ct = np.complex128 if prec == 8 else np.complex64
    x = (rng.random((outer, n, inner)) + 1j * rng.random((outer, n, inner))).astype(ct)
    tw = np.exp(-2j * np.pi * np.arange(n) / n).astype(ct)
    xin = x.copy()
    y = xin if inplace else np.full_like(x, np.nan)
    rc = emu.emu_fft_pow2(prec, n, var, outer, inner, xin.ctypes.data, y.ctypes.data, tw.ctypes.data, scale, swap)
    if rc != 0:
        return None
    x64 = x.astype(np.complex128)
    ref = (np.fft.ifft(x64, axis=1) * n if swap else np.fft.fft(x64, axis=1)) * scale
    return np.abs(y - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize('n', SIZES)
@pytest.mark.parametrize('prec', [8, 4])
def test_all_variants(emu, n, prec):
    tol = 2e-15 if prec == 8 else 2e-6
    found = 0
    for var in range(8):
        for outer, inner in ((3, 1), (2, 5), (1, 17), (1, 1)):
            if n >= 2048 and (outer, inner) == (1, 17):
                inner = 9
            for swap in (0, 1):
                err = run(emu, prec, n, var, outer, inner, swap, 1.0 / n if swap else 1.0, inplace=bool(swap))
                if err is None:
                    continue
                found += 1
                assert err < tol, (n, prec, var, outer, inner, swap, err)
    assert found >= (8 if n & (n - 1) == 0 else 6)      # at least one row, four geometries, both directions


def run_tma(emu, prec, n, var, outer, inner, swap, scale, inplace, seed=0):
    rng = np.random.default_rng(seed)
    ct = np.complex128 if prec == 8 else np.complex64
    x = (rng.random((outer, n, inner)) + 1j * rng.random((outer, n, inner))).astype(ct)
    xin = x.copy()
    y = xin if inplace else np.full_like(x, np.nan)
    rc = emu.emu_fft_tma(prec, n, var, outer, inner, xin.ctypes.data, y.ctypes.data, scale, swap)
    if rc != 0:
        return None
    x64 = x.astype(np.complex128)
    ref = (np.fft.ifft(x64, axis=1) * n if swap else np.fft.fft(x64, axis=1)) * scale
    return np.abs(y - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize('n', [64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize('prec', [8, 4])
def test_tma_staged_variants(emu, n, prec):
    """fft_tma.cuh: stage read, split / unsplit exchange, ragged last tile"""
    tol = 2e-15 if prec == 8 else 2e-6
    found = 0
    for var in range(10):
        for outer, inner in ((2, 8), (1, 19)):
            for swap in (0, 1):
                err = run_tma(emu, prec, n, var, outer, inner, swap, 1.0 / n if swap else 1.0, inplace=bool(swap))
                if err is None:
                    continue
                found += 1
                assert err < tol, (n, prec, var, outer, inner, swap, err)
    assert found >= 4


@pytest.mark.parametrize('nreal', [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384,
                                   6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 6144, 12288,
                                   10, 20, 40, 80, 160, 320, 640, 1280, 2560, 14, 28, 56, 112, 224, 448, 896, 1792, 3584])
@pytest.mark.parametrize('prec', [8, 4])
def test_real_transform_kernels(emu, nreal, prec):
    """fft_real_body (r2c / c2r of even length 2N through the N-point schedule plus
    the split / merge pass) against numpy.fft.rfft / irfft: contiguous and strided,
    ragged tiles, fused scale; c2r ignores the imaginary parts of X[0] and X[N] as
    FFTW does (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:57-67)"""
    rt = np.float64 if prec == 8 else np.float32
    ct = np.complex128 if prec == 8 else np.complex64
    tol = 3e-15 if prec == 8 else 2e-6
    rng = np.random.default_rng(nreal)
    geos = ((3, 1), (2, 5), (1, 17)) if nreal <= 2048 else ((2, 1), (1, 3))
    for outer, inner in geos:
        x = rng.random((outer, nreal, inner)).astype(rt)
        y = np.full((outer, nreal // 2 + 1, inner), np.nan, dtype=ct)
        assert emu.emu_fft_real(prec, nreal, 1, outer, inner, x.ctypes.data, y.ctypes.data, 0.5) == 0
        ref = np.fft.rfft(x.astype(np.float64), axis=1) * 0.5
        assert np.abs(y - ref).max() / np.abs(ref).max() < tol, (nreal, prec, outer, inner, 'r2c')
        X = (rng.random(y.shape) + 1j * rng.random(y.shape)).astype(ct)
        z = np.full((outer, nreal, inner), np.nan, dtype=rt)
        assert emu.emu_fft_real(prec, nreal, 2, outer, inner, X.copy().ctypes.data, z.ctypes.data, 1.0 / nreal) == 0
        ref = np.fft.irfft(X.astype(np.complex128), n=nreal, axis=1)
        assert np.abs(z - ref).max() / np.abs(ref).max() < tol, (nreal, prec, outer, inner, 'c2r')


def _chirp_case(emu, dft_ref, prec, kind, n, outer, inner, seed=0):
    rt = np.float64 if prec == 8 else np.float32
    ct = np.complex128 if prec == 8 else np.complex64
    rng = np.random.default_rng(seed)
    n_in = n // 2 + 1 if kind == 2 else n
    n_out = n // 2 + 1 if kind == -2 else n
    in_complex = kind in (-1, 1, 2)
    out_complex = kind in (-1, 1, -2)
    x = rng.random((outer, n_in, inner))
    if in_complex:
        x = x + 1j * rng.random((outer, n_in, inner))
    x = x.astype(ct if in_complex else rt)
    y = np.full((outer, n_out, inner), np.nan, dtype=ct if out_complex else rt)
    rc = emu.emu_chirpz(prec, kind, n, outer, inner, x.ctypes.data, y.ctypes.data, 0.25)
    assert rc == 0, (kind, n, rc)
    ref = np.empty(y.shape, dtype=complex if out_complex else float)
    for o in range(outer):
        for i in range(inner):
            ref[o, :, i] = dft_ref(kind, n, x[o, :, i].astype(complex if in_complex else float)) * 0.25
    return np.abs(y - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize('kind', [-1, 1, -2, 2, 3, 4, 5, 6, 7, 8, 9, 10])
@pytest.mark.parametrize('prec', [8, 4])
def test_chirpz_all_kinds_any_length(emu, dft_ref, kind, prec):
    """chirpz.cuh (pre-chirp, M-point FFT, filter, M-point FFT, post-chirp) for every
    transform kind the reference can ask FFTW for (fftw_planxfftn.c:49-76), odd /
    prime / composite lengths, against the O(n^2) long double restatement of the
    FFTW definitions (oracle/dft_ref.c); contiguous and strided with ragged tiles"""
    tol = 2e-14 if prec == 8 else 2e-5
    for n in (2, 3, 5, 6, 7, 12, 13, 31, 33, 64, 100, 127, 191):
        if kind == 3 and n < 2:
            continue
        for outer, inner in ((2, 1), (1, 5)):
            err = _chirp_case(emu, dft_ref, prec, kind, n, outer, inner, seed=n)
            assert err < tol, (kind, prec, n, outer, inner, err)


@pytest.mark.parametrize('n', [384, 1000, 1536, 2187, 4096])
def test_chirpz_large_lengths(emu, dft_ref, n):
    """3/2-rule sizes and the largest length that fits one tile (M = 8192)"""
    for kind in (-1, 5):
        err = _chirp_case(emu, dft_ref, 8, kind, n, 1, 1 if n > 1000 else 3, seed=1)
        assert err < 5e-14, (kind, n, err)
