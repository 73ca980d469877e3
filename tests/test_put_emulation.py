"""The peer-memory redistribution's index map (csrc/transfer_put.h, the code the
put kernel runs per 16-byte unit) replayed on the CPU for every rank of a group:
the result must be exactly what the reference's Alltoallw over subarray datatypes
produces (/root/reference/mpi4py_fft/pencil.py:12-29,182-183,200-201), i.e. block
i of A along axisA lands in B of rank i at this rank's block of axisB."""
import ctypes

import numpy as np
import pytest

from mpi4py_fft_b200.pencil import _blockdist

CASES = [
    # group-local shape, axisA, axisB, ranks
    ((6, 7, 5), 2, 1, 3),
    ((9, 4, 3), 1, 0, 2),
    ((5, 8, 3, 7), 3, 1, 4),
    ((10, 6), 1, 0, 4),
    ((10, 6), 0, 1, 3),
    ((4, 16, 8), 2, 1, 2),      # C3 T0 shape class (even split, wide vectors)
    ((16, 4, 8), 1, 0, 4),      # C3 T1 shape class
    ((3, 5, 4, 6, 2), 1, 3, 5),
    ((8, 8, 8), 0, 2, 8),
]


def blocks_of(g, shape, axis_full, axis_split, p):
    """per-rank arrays: full along axis_full, this rank's block along axis_split"""
    out = []
    for r in range(p):
        n, s = _blockdist(shape[axis_split], p, r)
        sl = [slice(None)] * len(shape)
        sl[axis_split] = slice(s, s + n)
        out.append(np.ascontiguousarray(g[tuple(sl)]))
    return out


@pytest.mark.parametrize('dtype', ['f4', 'f8', 'c8', 'c16', 'u1'])
@pytest.mark.parametrize('case', CASES)
def test_put_is_alltoallw(emu, case, dtype):
    shape, axisA, axisB, p = case
    dt = np.dtype(dtype)
    rng = np.random.default_rng(5)
    g = rng.integers(0, 250, size=shape).astype(dt)
    if dt.kind == 'c':
        g = g + 1j * rng.integers(0, 250, size=shape).astype(dt)
    A = blocks_of(g, shape, axisA, axisB, p)     # aligned on axisA: split along axisB
    Bx = blocks_of(g, shape, axisB, axisA, p)    # aligned on axisB: split along axisA
    shp = (ctypes.c_longlong * len(shape))(*shape)
    for direction, src_blocks, dst_expect, axS, axD in ((0, A, Bx, axisA, axisB), (1, Bx, A, axisB, axisA)):
        dst = [np.full_like(d, 251) for d in dst_expect]
        ptrs = (ctypes.c_void_p * p)(*[d.ctypes.data for d in dst])
        vec = ctypes.c_int(0)
        for r in range(p):
            rc = emu.emu_put(len(shape), shp, dt.itemsize, axS, axD, p, r, src_blocks[r].ctypes.data, ptrs,
                             ctypes.byref(vec))
            assert rc == 0
            assert vec.value in (1, 2, 4, 8, 16)
        for r in range(p):
            assert np.array_equal(dst[r], dst_expect[r]), (case, dtype, direction, r)


def test_put_uses_wide_vectors_for_headline_shapes(emu):
    """C3's two transfers at reduced size: rows are multiples of 16 bytes"""
    for shape, axS, axD, p in (((8, 16, 32), 2, 1, 2), ((8, 32, 16), 1, 0, 4)):
        g = np.zeros(shape, dtype='c16')
        src = blocks_of(g, shape, axS, axD, p)
        dst = blocks_of(g, shape, axD, axS, p)
        ptrs = (ctypes.c_void_p * p)(*[d.ctypes.data for d in dst])
        vec = ctypes.c_int(0)
        shp = (ctypes.c_longlong * 3)(*shape)
        assert emu.emu_put(3, shp, 16, axS, axD, p, 0, src[0].ctypes.data, ptrs, ctypes.byref(vec)) == 0
        assert vec.value == 16
