"""Host-side index maps of the PRODUCT (mpi4py_fft_b200: Subcomm, Pencil, PFFT
planning, transfer geometry in Python and in the C ABI) against the reference's
fixtures -- bit-exact, every rank evaluated as a virtual rank.  CPU only: no
device memory is touched by planning."""
import os

import numpy as np
import pytest

import mpi4py_fft_b200 as B
from mpi4py_fft_b200 import PFFT, Pencil, Subcomm, COMM_WORLD, Compute_dims
from mpi4py_fft_b200.comm import virtual_world
from mpi4py_fft_b200.pencil import _blockdist, transfer_geometry
from mpi4py_fft_b200.distarray import DistArray
from mpi4py_fft_b200 import fftw
import pfft_oracle as O
from conftest import case_kwargs
from test_oracle_golden import CASES


def test_blockdist_matches_oracle():
    for N in range(1, 40):
        for p in range(1, min(N, 9) + 1):
            got = [_blockdist(N, p, r) for r in range(p)]
            assert got == [O.blockdist(N, p, r) for r in range(p)]
            assert sum(n for n, _ in got) == N and got[0][1] == 0


def test_compute_dims(layouts):
    for key, val in layouts['_compute_dims'].items():
        n, d = key.split(':')
        assert Compute_dims(int(n), int(d)) == val
    assert Compute_dims(8, [0, 0, 1]) == [4, 2, 1]
    assert Compute_dims(12, [0, 3]) == [4, 3]
    assert Compute_dims(16, [4, 2, 0, 1]) == [4, 2, 2, 1]


def test_doc_goldens(layouts):
    for n in (4, 6):
        with virtual_world(n, 0):
            assert [c.Get_size() for c in Subcomm(COMM_WORLD, [0, 0, 1])] == layouts['_doc_subcomm_p%d' % n][0]
    for r in range(4):
        with virtual_world(4, r):
            s = Subcomm(COMM_WORLD, [0, 0, 1, 0])
            p0 = Pencil(s, (8, 8, 8, 8), 2)
            p1 = p0.pencil(0)
            assert [list(p0.subshape), list(p1.subshape)] == layouts['_doc_pencil_8x4_p4'][r]


@pytest.mark.parametrize('name', CASES)
def test_pfft_plan_bit_exact(layouts, name):
    case = layouts[name]
    n = case['meta']['nranks']
    for r, ref in enumerate(case['ranks']):
        with virtual_world(n, r):
            kw = case_kwargs(case['meta'])
            fft = PFFT(COMM_WORLD, **kw)
            assert [c.Get_size() for c in fft.subcomm] == ref['subcomm_sizes']
            assert [c.Get_rank() for c in fft.subcomm] == ref['subcomm_ranks']
            assert [list(a) for a in fft.axes] == ref['axes']
            assert list(fft.global_shape(False)) == ref['input_shape']
            assert list(fft.global_shape(True)) == ref['output_shape']
            assert list(fft.shape(False)) == ref['local_shape_in']
            assert list(fft.shape(True)) == ref['local_shape_out']
            assert [[s.start, s.stop] for s in fft.local_slice(False)] == ref['local_slice_in']
            assert [[s.start, s.stop] for s in fft.local_slice(True)] == ref['local_slice_out']
            assert fft.dtype(False).char == ref['dtype_in'] and fft.dtype(True).char == ref['dtype_out']
            assert fft.dimensions == len(ref['input_shape'])
            for st, rst in zip(fft.xfftn, ref['stages']):
                assert list(st.axes) == rst['axes']
                assert list(st.forward.input_shape) == rst['in_shape']
                assert st.forward.input_dtype.char == rst['in_dtype']
                assert list(st.forward.output_shape) == rst['out_shape']
                assert st.forward.output_dtype.char == rst['out_dtype']
            assert len(fft.transfer) == len(ref['transfers']) == len(fft.xfftn) - 1
            for tr, rtr in zip(fft.transfer, ref['transfers']):
                assert (tr.axisA, tr.axisB) == (rtr['axisA'], rtr['axisB'])
                assert list(tr.subshapeA) == rtr['subshapeA'] and list(tr.subshapeB) == rtr['subshapeB']
                assert list(tr.shape) == rtr['shape']
                assert tr.comm.Get_size() == rtr['group_size'] and tr.comm.Get_rank() == rtr['group_rank']
                assert tr.dtype.char == rtr['dtype']
            for key, pen in (('pencil_in', fft.pencil[0]), ('pencil_out', fft.pencil[1])):
                assert list(pen.subshape) == ref[key]['subshape']
                assert list(pen.substart) == ref[key]['substart']
                assert pen.axis == ref[key]['axis']
            # structural invariants the reference's tests assert (tests/test_mpifft.py:144-164)
            assert fft.forward.input_pencil.subshape == tuple(fft.shape(False))
            assert fft.backward.input_pencil.subshape == tuple(fft.shape(True))
            assert all(fft.forward.input_pencil.substart[a] == 0 for a in fft.axes[-1])
            assert fft.backward.input_pencil.substart[fft.axes[0][-1]] == 0


@pytest.mark.parametrize('name', ['c3_c2c_16_p8_pencil', 'uneven_r2c_12_13_14_p4', 'uneven_c2c_13_12_11_p6_axes201',
                                  'c5_c2c_8x4_p8_grid42'])
def test_transfer_geometry_python_vs_cabi(layouts, name):
    """per-peer counts/offsets: Python restatement == b2f_transfer_geometry ==
    the balanced blocks the reference's subarray datatypes select"""
    from mpi4py_fft_b200._lib import TransferHandle
    case = layouts[name]
    n = case['meta']['nranks']
    for r, ref in enumerate(case['ranks']):
        with virtual_world(n, r):
            fft = PFFT(COMM_WORLD, **case_kwargs(case['meta']))
            for tr in fft.transfer:
                geo = tr.geometry
                h = TransferHandle(tr.comm, tr.shape, tr.dtype.itemsize, tr.subshapeA, tr.axisA,
                                   tr.subshapeB, tr.axisB, exchange=False)
                cg = h.geometry()
                for key in ('send_counts', 'send_offsets', 'recv_counts', 'recv_offsets'):
                    assert cg[key] == geo[key], key
                p = tr.comm.Get_size()
                NA, NB = tr.shape[tr.axisA], tr.shape[tr.axisB]
                assert geo['blocksA'] == [(O.blockdist(NA, p, i)[1], O.blockdist(NA, p, i)[0]) for i in range(p)]
                assert sum(geo['send_counts']) == int(np.prod(tr.subshapeA))
                assert sum(geo['recv_counts']) == int(np.prod(tr.subshapeB))
                h.destroy()


def test_collapse_and_grid_variants():
    """grid wildcards, Subcomm / CART communicators as `comm`, slab keyword
    (reference mpifft.py:259-290, tests/test_mpifft.py:124-133)"""
    for r in range(8):
        with virtual_world(8, r):
            a = PFFT(COMM_WORLD, (16, 16, 16), dtype='D')
            b = PFFT(Subcomm(COMM_WORLD, [0, 0, 1]), (16, 16, 16), dtype='D')
            cart = COMM_WORLD.Create_cart(Compute_dims(8, [0, 0, 1]))
            c = PFFT(Subcomm(cart), (16, 16, 16), dtype='D')
            d = PFFT(COMM_WORLD, (16, 16, 16), dtype='D', grid=(4, 2))
            for f in (b, c, d):
                assert f.shape(True) == a.shape(True) and f.local_slice(False) == a.local_slice(False)
            s = PFFT(COMM_WORLD, (16, 16, 16), dtype='d', grid=(-1,), collapse=True)
            assert s.axes == ((0,), (1, 2)) and len(s.transfer) == 1
            s2 = PFFT(COMM_WORLD, (16, 16, 16), dtype='d', slab=True)
            assert [c_.Get_size() for c_ in s2.subcomm] == [8, 1, 1]


def test_planner_shapes_and_normalization():
    """output shape / dtype / normalisation rules of the planners
    (reference xfftn.py:228-239, 306-326, 763-816) -- specs only, no device"""
    from mpi4py_fft_b200.devarray import ArraySpec
    U = ArraySpec((6, 7, 8), 'd')
    f = fftw.rfftn(U, axes=(0, 2))
    assert f.output_shape == (6, 7, 5) and f.output_dtype == np.dtype('D') and f.get_normalization() == 1.0 / 48
    b = fftw.irfftn(ArraySpec((6, 7, 5), 'D'), s=(6, 8), axes=(0, 2))
    assert b.output_shape == (6, 7, 8) and b.output_dtype == np.dtype('d')
    assert fftw.irfftn(ArraySpec((4,), 'D')).output_shape == (6,)
    assert fftw.irfftn(ArraySpec((4,), 'D'), s=(7,)).output_shape == (7,)
    c = fftw.fftn(ArraySpec((4, 5), 'F'), axes=(1,))
    assert c.output_dtype == np.dtype('F') and c.get_normalization() == 0.2
    for typ, expect in ((1, 1.0 / (2 * 6)), (2, 1.0 / 14), (3, 1.0 / 14), (4, 1.0 / 14)):
        assert np.isclose(fftw.dctn(ArraySpec((7,), 'd'), type=typ).get_normalization(), expect)
    assert np.isclose(fftw.dstn(ArraySpec((7,), 'd'), type=1).get_normalization(), 1.0 / 16)
    assert fftw.get_normalization([fftw.FFTW_REDFT00, fftw.FFTW_RODFT00], (5, 6), (0, 1)) == 1.0 / (8 * 14)
    assert fftw.dct_type[2] == fftw.FFTW_REDFT10 and fftw.idct_type[2] == fftw.FFTW_REDFT01
    assert fftw.dst_type[3] == fftw.FFTW_RODFT01 and fftw.idst_type[3] == fftw.FFTW_RODFT10


def test_buffer_layout_of_chain():
    """which buffer every intermediate lives in (Transform._layout): single GPU
    c2c needs no work buffer; 8 ranks pencil uses two"""
    with virtual_world(1, 0):
        f = PFFT(COMM_WORLD, (32, 32, 32), dtype='D')
        assert f.forward._plan['a'] == ['IN', 'OUT', 'OUT'] and f.forward._plan['b'] == ['OUT', 'OUT', 'OUT']
        assert f.backward._plan['b'] == ['OUT', 'OUT', 'OUT']
        r = PFFT(COMM_WORLD, (32, 32, 32), dtype='d')
        assert r.forward._plan['b'] == ['OUT', 'OUT', 'OUT']
        assert r.backward._plan['a'][0] == 'IN' and r.backward._plan['b'][-1] == 'OUT'
        assert r.backward._plan['b'][0].startswith('W') and r.backward._plan['a'][2].startswith('W')
    with virtual_world(8, 5):
        # peer-memory transfers (default): a transfer only ever writes a plan-owned
        # window, the last stage runs out of place into the caller's array
        f = PFFT(COMM_WORLD, (32, 32, 32), dtype='D')
        lay = f.forward._plan
        assert lay['trivial'] == [False, False] and lay['windowed']
        assert lay['a'] == ['IN', 'W0', 'W1'] and lay['b'] == ['W1', 'W0', 'OUT']
        assert f.backward._plan['a'] == ['IN', 'W0', 'W1'] and f.backward._plan['b'] == ['W1', 'W0', 'OUT']
        assert f._buffers.need == {'W0': 16 * 8 * 32 * 16, 'W1': 16 * 8 * 32 * 16}
        # NCCL path: the last transfer may write the caller's array directly
        os.environ['B2F_P2P'] = '0'
        try:
            f = PFFT(COMM_WORLD, (32, 32, 32), dtype='D')
        finally:
            del os.environ['B2F_P2P']
        lay = f.forward._plan
        assert not lay['windowed']
        assert lay['a'][0] == 'IN' and lay['b'][0] != lay['a'][1] and lay['a'][2] == 'OUT' and lay['b'][2] == 'OUT'


def test_distarray_metadata_needs_no_device():
    """DistArray allocates device memory, so only its planning inputs are
    checked here: newDistArray's shape/pencil come from the PFFT"""
    with virtual_world(4, 2):
        f = PFFT(COMM_WORLD, (16, 14, 12), dtype='d')
        assert f.global_shape(True) == (16, 14, 7)
        assert f.pencil[True].axis == 0 and f.pencil[False].axis == 2


def test_pipeline_chunk_plans():
    """which axis the pipelined redistribution cuts (Transform._pipeline): the last
    axis where neither stage transforms it (inner ranges; re-viewed rows when other
    axes lie between), else the first axis (outer ranges); r2c/c2r stages and
    stages followed by another redistribution are not pipelined"""
    os.environ['B2F_PIPELINE'] = '4'
    try:
        with virtual_world(2, 0):
            odd = PFFT(COMM_WORLD, (64, 48, 96), dtype='D').forward._pipeline(1)
            assert [c[0][1:3] for c in odd] == [(0, 16), (16, 32), (48, 16), (64, 32)]     # cuts at multiples of 16
        with virtual_world(8, 3):
            f = PFFT(COMM_WORLD, (1024, 1024, 1024), dtype='D')       # C3: grid [4, 2, 1]
            assert [tuple(s.forward.input_shape) for s in f.xfftn] == [(256, 512, 1024), (256, 1024, 512), (1024, 256, 512)]
            assert f.forward._pipeline(0) is None                     # stage 1 feeds another redistribution
            fw = f.forward._pipeline(1)
            # producer: inner ranges of axis 2; consumer: 256 rows 512 apart, ranges of the same axis
            assert fw == [((1, 128 * j, 128, 0, 0), (1, 128 * j, 128, 256, 512)) for j in range(4)]
            assert f.backward._pipeline(0) is None
            bw = f.backward._pipeline(1)
            # backward: the consumer transforms the last axis, so the first axis is cut (outer ranges)
            assert bw == [((2, 64 * j, 64, 0, 0), (2, 64 * j * 512, 64 * 512, 0, 0)) for j in range(4)]
        with virtual_world(2, 1):
            f = PFFT(COMM_WORLD, (1024, 1024, 1024), dtype='D')       # slab [2, 1, 1]
            fw = f.forward._pipeline(1)
            assert fw == [((1, 256 * j, 256, 0, 0), (1, 256 * j, 256, 512, 1024)) for j in range(4)]
            bw = f.backward._pipeline(0)
            # producer transforms axis 0 of (1024, 512, 1024): re-viewed rows; consumer: plain inner ranges
            assert bw == [((1, 256 * j, 256, 512, 1024), (1, 256 * j, 256, 0, 0)) for j in range(4)]
            r = PFFT(COMM_WORLD, (64, 64, 64), dtype='d')
            assert r.backward._pipeline(0) is not None                # c2c consumer
            assert r.forward._pipeline(1) is not None
    finally:
        del os.environ['B2F_PIPELINE']


@pytest.mark.parametrize('nranks', [1, 2, 3, 4])
def test_padded_plan_matrix_vs_oracle(nranks):
    """the padded half of the reference's PFFT test matrix (tests/test_mpifft.py:179-236:
    shapes from (12, 13), padding 1.5 on every axis, the listed axes variants, slab
    and pencil grids): physical / spectral global shapes, every rank's local slices
    and the per-stage shapes of the product equal the oracle's restatement of
    mpifft.py:247-253 + libfft.py:424-434 (itself pinned against fixtures of the
    unmodified reference)"""
    from itertools import product
    import pfft_oracle as O
    sizes = (12, 13)
    checked = 0
    for dim, allaxes, grids in ((2, [None, (-1,), (-2,), (-1, -2), (-2, -1), (-1, 0), (0, -1), ((0,), (1,))], (None,)),
                                (3, [None, ((0,), (1,), (2,)), ((0,), (-2,), (-1,))], ((-1,), None))):
        for shape in product(*([sizes] * dim)):
            for dtype in 'dD':
                for grid in grids:
                    for axes in allaxes:
                        kw = dict(shape=shape, axes=axes, dtype=dtype, padding=[1.5] * dim, grid=grid)
                        try:
                            orc = O.OraclePFFT(nranks, shape, axes=axes, dtype=dtype, grid=grid, padding=[1.5] * dim)
                        except AssertionError:
                            continue        # more ranks than a distributed extent can hold (the reference skips these too)
                        lay = orc.layout()
                        for r in range(nranks):
                            with virtual_world(nranks, r):
                                fft = PFFT(COMM_WORLD, **{k: v for k, v in kw.items() if v is not None})
                                assert list(fft.global_shape(False)) == lay['input_shape'], kw
                                assert list(fft.global_shape(True)) == lay['output_shape'], kw
                                assert fft.dtype(True).char == lay['output_dtype']
                                assert fft.local_slice(False) == orc.local_slice(r, False), kw
                                assert fft.local_slice(True) == orc.local_slice(r, True), kw
                                for st, ost in zip(fft.xfftn, lay['ranks'][r]['stages']):
                                    assert list(st.axes) == ost['axes']
                                    assert list(st.forward.input_shape) == ost['in_subshape'], kw
                                    assert list(st.forward.output_shape) == ost['out_subshape'], kw
                                assert len(fft.axes) == len(fft.xfftn) == len(fft.transfer) + 1
                        checked += 1
    assert checked >= 40
