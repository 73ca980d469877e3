"""Fused stage + redistribution (csrc/fft_core.cuh PeerStore / TileFFT::store_peer)
stepped on the CPU: every rank of a group transforms the split axis of its block
with the kernels' own per-thread code and the last pass stores each point into
the owner's array.  Expected result = numpy FFT along that axis followed by the
reference's Alltoallw (/root/reference/mpi4py_fft/mpifft.py:70-74,
pencil.py:182-183)."""
import ctypes

import numpy as np
import pytest

from mpi4py_fft_b200.pencil import _blockdist

# group-local shape, axisS (transformed + split), axisD (gathered), ranks, staged kernel?
CASES = [
    ((6, 5, 64), 2, 1, 2, 0),      # contiguous FFT axis, destination gathers the axis before it (mode 0)
    ((6, 5, 64), 2, 0, 3, 0),      # mode 0 with a middle extent M = 5, uneven split of both axes
    ((3, 64, 10), 1, 0, 2, 0),     # strided FFT axis, mode 0
    ((3, 64, 10), 1, 2, 2, 0),     # mode 1: gathered axis after the transformed one
    ((64, 7, 6), 0, 2, 4, 0),      # mode 1 with M = 7
    ((64, 7, 6), 0, 1, 3, 0),
    ((2, 3, 128, 5, 9), 2, 4, 2, 0),
    ((2, 9, 128, 5, 3), 2, 1, 4, 0),
    ((3, 64, 16), 1, 0, 2, 1),     # TMA/cp.async-staged kernel flavour
    ((64, 4, 24), 0, 2, 3, 1),
    ((8, 32, 1024), 2, 1, 4, 0),   # C3 stage-0 class
]


def split_blocks(g, axis, p):
    out = []
    for r in range(p):
        n, s = _blockdist(g.shape[axis], p, r)
        sl = [slice(None)] * g.ndim
        sl[axis] = slice(s, s + n)
        out.append(np.ascontiguousarray(g[tuple(sl)]))
    return out


@pytest.mark.parametrize('prec', [8, 4])
@pytest.mark.parametrize('case', CASES)
def test_fused_stage_and_transfer(emu, case, prec):
    shape, axS, axD, p, staged = case
    ct = np.complex128 if prec == 8 else np.complex64
    rng = np.random.default_rng(9)
    g = (rng.random(shape) + 1j * rng.random(shape)).astype(ct)
    n = shape[axS]
    shp = (ctypes.c_longlong * len(shape))(*shape)
    for swap in (0, 1):
        g64 = g.astype(np.complex128)
        full = (np.fft.ifft(g64, axis=axS) * n if swap else np.fft.fft(g64, axis=axS)) / n
        expect = split_blocks(full, axS, p)            # destination arrays: split along the transformed axis
        src = split_blocks(g, axD, p)                  # source arrays: split along the axis to gather
        dst = [np.full(e.shape, np.nan, dtype=ct) for e in expect]
        ptrs = (ctypes.c_void_p * p)(*[d.ctypes.data for d in dst])
        for r in range(p):
            rc = emu.emu_fft_scatter(prec, len(shape), shp, axS, axD, p, r, 0, staged, src[r].ctypes.data, ptrs,
                                     1.0 / n, swap)
            assert rc == 0, (case, rc)
        tol = 2e-15 if prec == 8 else 2e-6
        for r in range(p):
            err = np.abs(dst[r] - expect[r]).max() / np.abs(full).max()
            assert err < tol, (case, prec, swap, r, err)


# group-local shape, axisS, axisD, ranks, how the stage is cut: 'inner', 'outer' or 'rows'
# ('rows' = ranges of the last axis of a block whose FIRST axis is transformed, re-viewed)
CHUNKED = [
    ((6, 64, 40), 1, 0, 2, 'inner'),
    ((6, 64, 40), 1, 2, 3, 'outer'),
    ((8, 5, 128), 2, 1, 2, 'outer'),
    ((64, 12, 40), 0, 1, 3, 'rows'),
    ((64, 12, 40), 0, 2, 2, 'rows'),
    ((4, 1024, 24), 1, 0, 4, 'inner'),
]


@pytest.mark.parametrize('case', CHUNKED)
def test_fused_stage_in_pieces(emu, case):
    """the pipelined redistribution's producer: the fused stage launched in pieces
    (PeerStore ooff / ioff / vstride, shifted base pointers) stores exactly what
    the whole launch stores -- the library's parameter edits restated in the
    emulator entry point, the kernels' own per-thread code underneath"""
    shape, axS, axD, p, how = case
    rng = np.random.default_rng(3)
    g = rng.random(shape) + 1j * rng.random(shape)
    n = shape[axS]
    shp = (ctypes.c_longlong * len(shape))(*shape)
    full = np.fft.fft(g, axis=axS)
    expect = split_blocks(full, axS, p)
    src = split_blocks(g, axD, p)
    dst = [np.full(e.shape, np.nan, dtype=complex) for e in expect]
    ptrs = (ctypes.c_void_p * p)(*[d.ctypes.data for d in dst])
    for r in range(p):
        ss = src[r].shape
        if how == 'outer':
            extent, mode, view = int(np.prod(ss[:axS])), 2, (0, 0)
        elif how == 'inner':
            extent, mode, view = int(np.prod(ss[axS + 1:])), 1, (0, 0)
        else:
            extent, mode, view = ss[-1], 1, (int(np.prod(ss[1:-1])), ss[-1])
        cuts = sorted(set([0, extent // 3, extent // 3 + 1, extent]))
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            rc = emu.emu_fft_scatter_chunk(8, len(shape), shp, axS, axD, p, r, src[r].ctypes.data, ptrs, 1.0, 0,
                                           mode, lo, hi - lo, view[0], view[1])
            assert rc == 0, (case, r, lo, hi, rc)
    for r in range(p):
        err = np.abs(dst[r] - expect[r]).max() / np.abs(full).max()
        assert err < 2e-15, (case, r, err)
