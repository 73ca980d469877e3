"""PFFT / DistArray / Transfer on the device against the oracle and the
reference's fixtures.

  * one rank: the product end to end (PFFT.forward/backward through the C ABI);
  * several ranks on ONE GPU: every rank of a golden multi-rank case is played
    in turn -- the product's own plans, stage kernels and pack/unpack kernels
    run on the device, only the wire (NCCL) is replaced by device copies of the
    packed segments.  Results are compared with the gathered output of the
    unmodified reference (tests/golden/values.npz);
  * full-size properties at BASELINE config C2 (512^3 complex128): spot values
    against a direct DFT, round trip, linearity, Parseval.
"""
from itertools import product

import numpy as np
import pytest

import pfft_oracle as O
from conftest import case_kwargs

pytestmark = pytest.mark.gpu

TOL = {'d': 1e-12, 'f': 1e-5}


def relerr(a, ref):
    return np.abs(np.asarray(a) - ref).max() / max(np.abs(ref).max(), 1e-300)


def rand(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    dtype = np.dtype(dtype)
    x = rng.random(shape)
    if dtype.char in 'FD':
        x = x + 1j * rng.random(shape)
    return x.astype(dtype)


@pytest.fixture(scope='module')
def B():
    import torch
    import mpi4py_fft_b200 as B
    torch.cuda.set_device(0)
    return B


# ---------------------------------------------------------------------------
# one rank
# ---------------------------------------------------------------------------
def test_single_rank_matches_reference_fixture(B, layouts, values):
    """golden case c2c_32_p1 has no stored values (size); use the 2-rank input of
    c1 on one rank: the global result must not depend on the decomposition"""
    g = values['c1_c2c_16_p2__input']
    fft = B.PFFT(B.COMM_WORLD, g.shape, dtype='D')
    u = B.newDistArray(fft, False)
    u[...] = g
    uh = fft.forward(u)
    assert relerr(uh, values['c1_c2c_16_p2__forward']) < 1e-12
    ub = fft.backward(uh)
    assert relerr(ub, g) < 1e-12
    assert np.array_equal(np.asarray(u), g)          # caller's input preserved


@pytest.mark.parametrize('dt', ['f', 'd', 'F', 'D'])
def test_single_rank_matrix(B, dt):
    """the reference's PFFT test matrix (tests/test_mpifft.py:53-177) on one rank:
    dims 2-4, sizes (12, 13) plus powers of two, axes variants, collapse; values
    against the transform of the global array and round trip"""
    tol = TOL[dt.lower()]
    axes_by_dim = {2: [None, (-1,), (-2,), (-1, -2), (-2, -1), (-1, 0), (0, -1), ((0,), (1,))],
                   3: [None, ((0,), (1, 2)), ((0,), (-2, -1))],
                   4: [None, ((0,), (1,), (2,), (3,)), ((0,), (1, 2, 3)), ((0,), (1,), (2, 3))]}
    for dim in (2, 3, 4):
        sizes = (12, 13) if dim < 4 else (6, 5)
        shapes = list(product(*([sizes] * dim))) + [(16,) * dim, (8, 32, 4, 2)[:dim]]
        if dim == 4:
            shapes = shapes[::5]
        for shape in shapes:
            for collapse in (False, True):
                for axes in axes_by_dim[dim]:
                    fft = B.PFFT(B.COMM_WORLD, shape, axes=axes, dtype=dt, collapse=collapse)
                    g = rand(shape, dt, 11)
                    u = B.newDistArray(fft, False)
                    u[...] = g
                    uh = fft.forward(u)
                    g64 = g.astype('D' if dt in 'FD' else 'd')
                    ref = O.expected_forward(g64, axes)
                    assert tuple(uh.shape) == ref.shape == fft.global_shape(True)
                    assert relerr(uh, ref) < tol, (shape, axes, collapse, dt)
                    ub = B.newDistArray(fft, False)
                    ub = fft.backward(uh, ub)
                    assert relerr(ub, g64) < 10 * tol, (shape, axes, collapse, dt)
                    # implicit arrays (tests/test_mpifft.py:173-176)
                    fft.forward.input_array[...] = g
                    fft.forward()
                    fft.backward()
                    assert relerr(fft.backward.output_array, g64) < 10 * tol
                    fft.destroy()


def test_single_rank_r2r(B):
    """tests/test_mpifft.py:35-51: DCT-III on (1,2), DST-III on (3,4), Fourier on 0"""
    import functools
    fftw = B.fftw
    N = (5, 6, 7, 8, 9)
    tr = {(1, 2): (functools.partial(fftw.dctn, type=3), functools.partial(fftw.idctn, type=3)),
          (3, 4): (functools.partial(fftw.dstn, type=3), functools.partial(fftw.idstn, type=3))}
    fft = B.PFFT(B.COMM_WORLD, N, axes=((0,), (1, 2), (3, 4)), grid=(-1,), transforms=tr)
    g = rand(N, 'd', 3)
    A = B.newDistArray(fft, False)
    A[...] = g
    Bh = fft.forward(A)
    ref = O.expected_forward(g, ((0,), (1, 2), (3, 4)), {(1, 2): ('dct', 3), (3, 4): ('dst', 3)})
    assert relerr(Bh, ref) < 1e-12
    C = fft.backward(Bh)
    assert relerr(C, g) < 1e-12


def test_distarray_basics(B):
    """tests/test_darray.py of the reference, one rank"""
    from mpi4py_fft_b200 import DistArray, newDistArray, PFFT
    z = DistArray((8, 6, 4), dtype=float, val=2.0)
    assert z.global_shape == (8, 6, 4) and z.alignment == 2 and z.rank == 0 and z.dimensions == 3
    assert z.local_slice() == (slice(0, 8), slice(0, 6), slice(0, 4))
    assert float(np.asarray(z).sum()) == 2.0 * 8 * 6 * 4
    v = DistArray((3, 8, 6, 4), rank=1, val=1.0)
    assert v.pencil.shape == (8, 6, 4) and v.rank == 1
    assert isinstance(v[0], DistArray) and v[0].rank == 0 and v[0].shape == (8, 6, 4)
    assert not isinstance(v[0, 1], DistArray)
    w = v * 2.0 + v
    assert isinstance(w, DistArray) and float(np.asarray(w).max()) == 3.0
    fft = PFFT(B.COMM_WORLD, darray=z)
    assert fft.axes == ((0,), (1,), (2,))
    z1 = z.redistribute(0)
    assert z1 is z and z1.alignment == 0
    u = newDistArray(fft, False, rank=1)
    assert u.shape == (3, 8, 6, 4)
    with pytest.raises(NotImplementedError):
        z.write('x.h5')


# ---------------------------------------------------------------------------
# all ranks of a multi-rank case on one device
# ---------------------------------------------------------------------------
class Played(object):
    """The product's PFFT for every rank of a job (planned under virtual_world),
    executed stage by stage; a transfer packs on every rank, moves the packed
    segments the way the all-to-all would, and unpacks on every rank."""

    def __init__(self, B, nranks, kw):
        from mpi4py_fft_b200.comm import virtual_world
        from mpi4py_fft_b200._lib import TransferHandle
        self.B, self.n = B, nranks
        self.ffts, self.handles = [], []
        for r in range(nranks):
            with virtual_world(nranks, r):
                f = B.PFFT(B.COMM_WORLD, **kw)
                self.ffts.append(f)
                self.handles.append([TransferHandle(t.comm, t.shape, t.dtype.itemsize, t.subshapeA, t.axisA,
                                                    t.subshapeB, t.axisB, exchange=False) for t in f.transfer])

    def _exchange(self, ti, arrays, backward):
        import torch
        B, n = self.B, self.n
        trs = [f.transfer[ti] for f in self.ffts]
        direction = 1 if backward else 0
        packed, geos = [], []
        for r in range(n):
            t = trs[r]
            src = arrays[r]
            buf = B.fftw.aligned(src.shape, dtype=src.dtype)
            self.handles[r][ti].pack(direction, src, buf)
            packed.append(buf.tensor.reshape(-1))
            geos.append(t.geometry)
        out = []
        for r in range(n):
            t = trs[r]
            group = t.comm.ranks
            me = t.comm.Get_rank()
            shape = t.subshapeA if backward else t.subshapeB
            recv = torch.empty(int(np.prod(shape)), dtype=packed[r].dtype, device=packed[r].device)
            rc = geos[r]['send_counts' if backward else 'recv_counts']
            ro = geos[r]['send_offsets' if backward else 'recv_offsets']
            for j, peer in enumerate(group):
                sc = geos[peer]['recv_counts' if backward else 'send_counts']
                so = geos[peer]['recv_offsets' if backward else 'send_offsets']
                assert sc[me] == rc[j]
                recv[ro[j]:ro[j] + rc[j]] = packed[peer][so[me]:so[me] + sc[me]]
            dst = B.fftw.aligned(shape, dtype=arrays[r].dtype)
            self.handles[r][ti].unpack(direction, B.DeviceArray(recv), dst)
            out.append(dst)
        return out

    def _stage_then_transfer(self, cur, stage_of, ti, backward):
        """stage on every rank followed by transfer ``ti`` (None: no transfer).  In
        ``self.mode`` 'fused' the stage's last pass stores straight into the arrays of
        the group's ranks (b2f_execute_scatter, no ordering needed on one device:
        every rank's launch is on the same stream); 'put' runs the put kernel;
        'pack' moves packed segments by hand."""
        from mpi4py_fft_b200.devarray import device_ptr
        B = self.B
        nxt = []
        trs = [f.transfer[ti] for f in self.ffts] if ti is not None else None
        direction = 1 if backward else 0
        p2p = trs is not None and self.mode in ('fused', 'put') and trs[0].comm.Get_size() > 1
        recv = None
        if p2p:
            recv = [B.fftw.aligned(t.subshapeA if backward else t.subshapeB, dtype=t.dtype, fill=0) for t in trs]
        self.fused_count = getattr(self, 'fused_count', 0)
        for r, f in enumerate(self.ffts):
            st = stage_of(f)
            dst = B.fftw.aligned(st.output_shape, dtype=st.output_dtype)
            if p2p:
                ptrs = [device_ptr(recv[w]) for w in trs[r].comm.ranks]
                if self.mode == 'fused' and st.can_scatter(self.handles[r][ti], direction):
                    st.run_scatter(cur[r], dst, None, self.handles[r][ti], direction, ptrs_sync_off(ptrs))
                    self.fused_count += 1
                else:
                    st.run(cur[r], dst)
                    self.handles[r][ti].put(direction, dst, ptrs)
            else:
                st.run(cur[r], dst)
            nxt.append(dst)
        if p2p:
            return recv
        if trs is not None:
            return self._exchange(ti, nxt, backward=backward)
        return nxt

    mode = 'pack'

    def forward(self, blocks):
        B = self.B
        cur = []
        for r, f in enumerate(self.ffts):
            a = B.fftw.aligned(blocks[r].shape, dtype=blocks[r].dtype)
            a[...] = blocks[r]
            cur.append(a)
        nst = len(self.ffts[0].xfftn)
        for i in range(nst):
            cur = self._stage_then_transfer(cur, lambda f: f.xfftn[i].forward, i if i + 1 < nst else None, False)
        return cur

    def backward(self, blocks):
        cur = list(blocks)
        nst = len(self.ffts[0].xfftn)
        for i in range(nst - 1, -1, -1):
            cur = self._stage_then_transfer(cur, lambda f: f.xfftn[i].backward, i - 1 if i > 0 else None, True)
        return cur


class ptrs_sync_off(list):
    """peer pointer list that tells _Stage.run_scatter to skip the group barriers
    (all ranks of the job share one stream in these tests)"""
    sync = False


MULTI = ['c1_c2c_16_p2', 'c3_c2c_16_p8_pencil', 'c4_r2c_16_p8_slab', 'c4_r2c_16_p8_slab_collapse',
         'c5_c2c_8x4_p8_grid42', 'uneven_r2c_12_13_14_p4', 'uneven_c2c_13_12_11_p6_axes201',
         'uneven_c2c_7_9_p3_2d', 'c2c_4d_nested_p4', 'r2c_3d_nested_collapse_p4',
         'pad_c2c_8_p4_3half', 'pad_r2c_8_12_10_p4_3half', 'pad_c2c_9_7_p2_mixed', 'pad_r2c_10_9_8_p1']


@pytest.mark.parametrize('mode', ['pack', 'put', 'fused'])
@pytest.mark.parametrize('name', MULTI)
def test_all_ranks_played_on_one_gpu(B, layouts, values, name, mode):
    """every golden multi-rank case with all its ranks on one device, through the
    three implementations of the redistribution (pack/unpack by hand, put kernel,
    stage fused with the redistribution)"""
    case = layouts[name]
    n = case['meta']['nranks']
    kw = case_kwargs(case['meta'])
    g = values[name + '__input']
    tol = TOL[g.dtype.char.lower()]
    job = Played(B, n, kw)
    job.mode = mode
    ranks = case['ranks']
    blocks = [np.ascontiguousarray(g[tuple(slice(a, b) for a, b in ranks[r]['local_slice_in'])]) for r in range(n)]
    out = job.forward(blocks)
    ref = values[name + '__forward']
    for r in range(n):
        sl = tuple(slice(a, b) for a, b in ranks[r]['local_slice_out'])
        assert tuple(out[r].shape) == tuple(ranks[r]['local_shape_out'])
        assert np.abs(np.asarray(out[r]) - ref[sl]).max() <= tol * max(1.0, np.abs(ref).max()), (name, r)
    back = job.backward(out)
    refb = values[name + '__backward']      # == the input, except for padded (lossy) transforms
    for r in range(n):
        slin = tuple(slice(a, b) for a, b in ranks[r]['local_slice_in'])
        assert np.abs(np.asarray(back[r]) - refb[slin]).max() <= 10 * tol, (name, r)


SERIAL_PAD = ['spad_c_even', 'spad_c_odd', 'spad_c_first_axis', 'spad_r_even_half', 'spad_r_odd_half',
              'spad_r_evenhalf2', 'spad_r_first_axis']


@pytest.mark.parametrize('name', SERIAL_PAD)
def test_padded_serial_stage_vs_reference(B, layouts, values, name):
    """libfft.FFT(padding=...) on the device against the reference's own outputs
    (fixtures made by oracle/make_golden.py from the unmodified libfft.FFT): the
    truncation / zero padding and its Nyquist rule (libfft.py:263-311), complex and
    real, odd and even extents, both normalisation switches"""
    import pfft_oracle as O
    meta = layouts['_' + name]
    x, y, z = values[name + '__input'], values[name + '__forward'], values[name + '__backward']
    f = B.FFT(meta['shape'], axes=(meta['axis'],), dtype=meta['dtype'], padding=meta['padding'])
    assert list(f.forward.output_shape) == meta['trunc_shape']
    u = B.fftw.aligned(meta['shape'], dtype=meta['dtype'])
    u[...] = x
    got = np.asarray(f.forward(u)).copy()
    assert np.abs(got - y).max() <= 1e-12 * max(1.0, np.abs(y).max())
    raw = np.asarray(f.forward(u, normalize=False)).copy()
    assert np.abs(raw - y * meta['shape'][meta['axis']]).max() <= 1e-11 * max(1.0, np.abs(raw).max())
    v = B.fftw.aligned(meta['trunc_shape'], dtype=meta['trunc_dtype'])
    v[...] = y
    back = np.asarray(f.backward(v)).copy()
    assert np.abs(back - z).max() <= 1e-12 * max(1.0, np.abs(z).max())
    # host arrays in and out, as scripts written for the reference pass them
    out = np.zeros(meta['trunc_shape'], dtype=meta['trunc_dtype'])
    f.forward(x, out)
    assert np.abs(out - y).max() <= 1e-12 * max(1.0, np.abs(y).max())
    assert np.abs(got - O.padded_stage_forward(x, meta['axis'], meta['padding'])).max() <= 1e-12


@pytest.mark.parametrize('itemsize_dtype', ['f', 'd', 'F', 'D'])
def test_pack_unpack_kernels_vs_numpy(B, itemsize_dtype):
    """pack == concatenation of the axis blocks in C order (what the reference's
    subarray datatypes describe, pencil.py:12-29), unpack is its inverse; odd
    extents exercise the 4/8/16-byte paths"""
    from mpi4py_fft_b200.comm import virtual_world
    from mpi4py_fft_b200._lib import TransferHandle
    from mpi4py_fft_b200.pencil import _blockdist
    dt = np.dtype(itemsize_dtype)
    for shape, axisA, axisB, p in (((6, 7, 5), 2, 1, 3), ((9, 4, 3), 1, 0, 2), ((5, 8, 3, 7), 3, 1, 4), ((10, 6), 1, 0, 4)):
        for rank in range(p):
            nB, sB = _blockdist(shape[axisB], p, rank)
            nA, sA = _blockdist(shape[axisA], p, rank)
            subA = list(shape)
            subA[axisB] = nB
            subB = list(shape)
            subB[axisA] = nA

            class FakeComm(object):
                ranks = tuple(range(p))

                def Get_size(self):
                    return p

                def Get_rank(self):
                    return rank
            h = TransferHandle(FakeComm(), shape, dt.itemsize, subA, axisA, subB, axisB, exchange=False)
            for direction, sub, axis in ((0, subA, axisA), (1, subB, axisB)):
                x = rand(sub, dt, 4)
                src = B.fftw.aligned(sub, dtype=dt)
                src[...] = x
                packed = B.fftw.aligned(sub, dtype=dt)
                h.pack(direction, src, packed)
                expect = np.concatenate([np.ascontiguousarray(
                    np.take(x, range(_blockdist(shape[axis], p, i)[1], sum(_blockdist(shape[axis], p, i))), axis=axis)).ravel()
                    for i in range(p)])
                assert np.array_equal(np.asarray(packed).ravel(), expect), (shape, axis, direction)
                # unpack of the other direction scatters back into the same layout
                dst = B.fftw.aligned(sub, dtype=dt, fill=0)
                h.unpack(1 - direction, packed, dst)
                assert np.array_equal(np.asarray(dst), x)
            h.destroy()


@pytest.mark.parametrize('itemsize_dtype', ['f', 'd', 'F', 'D'])
def test_put_kernel_is_alltoallw(B, itemsize_dtype):
    """the peer-memory put kernel, every rank of a group played on one device
    (the peers' windows are plain local arrays): block i of A along axisA lands in
    B of rank i at the sender's block of axisB (reference pencil.py:12-29,182-183),
    and the backward direction is its inverse"""
    from mpi4py_fft_b200._lib import TransferHandle
    from mpi4py_fft_b200.devarray import device_ptr
    from mpi4py_fft_b200.pencil import _blockdist
    dt = np.dtype(itemsize_dtype)
    for shape, axisA, axisB, p in (((6, 7, 5), 2, 1, 3), ((9, 4, 3), 1, 0, 2), ((5, 8, 3, 7), 3, 1, 4),
                                   ((10, 6), 1, 0, 4), ((8, 64, 32), 2, 1, 2), ((8, 32, 64), 1, 0, 4),
                                   ((33, 40, 24), 0, 2, 8)):
        g = rand(shape, dt, 11)

        def blocks(axis_split):
            out = []
            for r in range(p):
                n, s0 = _blockdist(shape[axis_split], p, r)
                sl = [slice(None)] * len(shape)
                sl[axis_split] = slice(s0, s0 + n)
                out.append(np.ascontiguousarray(g[tuple(sl)]))
            return out
        A, Bx = blocks(axisB), blocks(axisA)
        handles = []
        for rank in range(p):
            class FakeComm(object):
                ranks = tuple(range(p))
                _r = rank

                def Get_size(self):
                    return p

                def Get_rank(self):
                    return self._r
            handles.append(TransferHandle(FakeComm(), shape, dt.itemsize, A[rank].shape, axisA, Bx[rank].shape, axisB,
                                          exchange=False))
        for direction, src_np, dst_np in ((0, A, Bx), (1, Bx, A)):
            src = []
            for r in range(p):
                a = B.fftw.aligned(src_np[r].shape, dtype=dt)
                a[...] = src_np[r]
                src.append(a)
            dst = [B.fftw.aligned(d.shape, dtype=dt, fill=0) for d in dst_np]
            ptrs = [device_ptr(d) for d in dst]
            for r in range(p):
                handles[r].put(direction, src[r], ptrs)
            for r in range(p):
                assert np.array_equal(np.asarray(dst[r]), dst_np[r]), (shape, axisA, axisB, p, direction, r)
        for h in handles:
            h.destroy()


@pytest.mark.parametrize('dtype', ['D', 'F'])
def test_fused_stage_and_transfer_kernels(B, dtype):
    """b2f_execute_scatter, every rank of a group played on one device: FFT along
    the split axis with the last pass storing into the owners' arrays == numpy FFT
    followed by the reference's Alltoallw (mpifft.py:70-74, pencil.py:182-183);
    covers the contiguous, register-strided and staged (TMA / cp.async) kernels,
    both PeerStore modes, uneven splits and a two-axis stage"""
    from mpi4py_fft_b200._lib import TransferHandle, Plan
    from mpi4py_fft_b200.devarray import device_ptr
    from mpi4py_fft_b200.pencil import _blockdist
    dt = np.dtype(dtype)
    prec = 8 if dtype == 'D' else 4
    tol = 1e-12 if dtype == 'D' else 1e-5
    cases = [((6, 5, 64), (2,), 1, 2), ((6, 5, 64), (2,), 0, 3), ((3, 64, 10), (1,), 0, 2), ((3, 64, 10), (1,), 2, 2),
             ((64, 7, 6), (0,), 2, 4), ((2, 9, 128, 5, 3), (2,), 1, 4), ((4, 256, 48), (1,), 0, 2),
             ((512, 6, 40), (0,), 2, 3), ((8, 32, 1024), (2,), 1, 4), ((16, 1024, 24), (1,), 0, 4),
             ((6, 16, 64), (1, 2), 0, 2)]
    for shape, axes, axD, p in cases:
        axS = axes[-1]
        g = rand(shape, dt, 21)

        def blocks(arr, axis):
            out = []
            for r in range(p):
                n, s0 = _blockdist(shape[axis], p, r)
                sl = [slice(None)] * len(shape)
                sl[axis] = slice(s0, s0 + n)
                out.append(np.ascontiguousarray(arr[tuple(sl)]))
            return out
        for kind in (-1, 1):
            g64 = g.astype(np.complex128)
            full = np.fft.fftn(g64, axes=axes) if kind == -1 else np.fft.ifftn(g64, axes=axes) * np.prod([shape[a] for a in axes])
            expect = blocks(full, axS)
            src_np = blocks(g, axD)
            dst = [B.fftw.aligned(e.shape, dtype=dt, fill=0) for e in expect]
            ptrs = [device_ptr(d) for d in dst]
            for r in range(p):
                class FakeComm(object):
                    ranks = tuple(range(p))
                    _r = r

                    def Get_size(self):
                        return p

                    def Get_rank(self):
                        return self._r
                # transfer A -> B with axisA = axS: A blocks are split along axD, B blocks along axS
                h = TransferHandle(FakeComm(), shape, dt.itemsize, src_np[r].shape, axS, expect[r].shape, axD,
                                   exchange=False)
                plan = Plan(src_np[r].shape, src_np[r].shape, axes, [kind] * len(axes), prec)
                assert plan.can_scatter(h, 0)
                a = B.fftw.aligned(src_np[r].shape, dtype=dt)
                a[...] = src_np[r]
                work = B.fftw.aligned(src_np[r].shape, dtype=dt)
                plan.execute_scatter(device_ptr(a), device_ptr(work), 1.0, h, 0, ptrs, sync=False)
                plan.destroy()
                h.destroy()
            for r in range(p):
                err = np.abs(np.asarray(dst[r]) - expect[r]).max() / np.abs(full).max()
                assert err < tol, (shape, axes, axD, p, kind, r, err)


@pytest.mark.parametrize('dtype', ['D', 'F'])
def test_partial_stage_launches(B, dtype):
    """b2f_execute_chunk / b2f_execute_scatter_chunk: a one-axis stage launched in
    pieces (ranges of the inner index, of the last array axis through a re-viewed
    block, or of the outer index) writes exactly what the whole launch writes; the
    scatter form stores the pieces into the owners' arrays.  These are the launches
    the pipelined redistribution is made of."""
    from mpi4py_fft_b200._lib import TransferHandle, Plan
    from mpi4py_fft_b200.devarray import device_ptr
    from mpi4py_fft_b200.pencil import _blockdist
    dt = np.dtype(dtype)
    prec = 8 if dtype == 'D' else 4
    tol = 1e-12 if dtype == 'D' else 1e-5
    for shape, axis in (((6, 64, 40), 1), ((3, 1024, 24), 1), ((5, 512), 1), ((256, 12, 40), 0), ((1024, 3, 24), 0),
                        ((4, 384, 16), 1)):
        x = rand(shape, dt, 5)
        a = B.fftw.aligned(shape, dtype=dt)
        a[...] = x
        ref = np.fft.fft(x.astype('D'), axis=axis)
        plan = Plan(shape, shape, (axis,), [-1], prec)
        outer = int(np.prod(shape[:axis]))
        inner = int(np.prod(shape[axis + 1:]))
        last = shape[-1]
        # outer ranges
        if outer > 1:
            out = B.fftw.aligned(shape, dtype=dt, fill=0)
            cuts = [0, outer // 2, outer]
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                plan.execute_chunk(device_ptr(a), device_ptr(out), 1.0, 2, lo, hi - lo)
            assert relerr(out, ref) < tol, ('outer', shape)
        if inner > 1:
            # inner ranges
            out = B.fftw.aligned(shape, dtype=dt, fill=0)
            cuts = [0, inner // 3, inner // 3 + 5, inner]
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                plan.execute_chunk(device_ptr(a), device_ptr(out), 1.0, 1, lo, hi - lo, grid_cap=40)
            assert relerr(out, ref) < tol, ('inner', shape)
        if axis == 0 and len(shape) == 3:
            # ranges of the last axis: rows = the middle axis, row pitch = the last extent
            out = B.fftw.aligned(shape, dtype=dt, fill=0)
            cuts = [0, 8, last - 3, last]
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                plan.execute_chunk(device_ptr(a), device_ptr(out), 1.0, 1, lo, hi - lo, view_outer=shape[1],
                                   view_ostride=last)
            assert relerr(out, ref) < tol, ('last axis', shape)
        plan.destroy()
    # scatter form: FFT along axS of each rank's block, pieces stored into the owners' arrays
    for shape, axS, axD, p, mode in (((8, 64, 48), 1, 0, 2, 1), ((8, 64, 48), 1, 2, 4, 2), ((6, 5, 128), 2, 1, 3, 2),
                                     ((16, 1024, 24), 1, 0, 4, 1), ((64, 12, 40), 0, 1, 3, 3), ((512, 8, 24), 0, 1, 2, 3)):
        g = rand(shape, dt, 9)
        full = np.fft.fft(g.astype('D'), axis=axS)

        def blocks(arr, ax):
            out = []
            for r in range(p):
                n, s0 = _blockdist(shape[ax], p, r)
                sl = [slice(None)] * len(shape)
                sl[ax] = slice(s0, s0 + n)
                out.append(np.ascontiguousarray(arr[tuple(sl)]))
            return out
        expect, src_np = blocks(full, axS), blocks(g, axD)
        dst = [B.fftw.aligned(e.shape, dtype=dt, fill=0) for e in expect]
        ptrs = [device_ptr(d) for d in dst]
        for r in range(p):
            class FakeComm(object):
                ranks = tuple(range(p))
                _r = r

                def Get_size(self):
                    return p

                def Get_rank(self):
                    return self._r
            h = TransferHandle(FakeComm(), shape, dt.itemsize, src_np[r].shape, axS, expect[r].shape, axD, exchange=False)
            plan = Plan(src_np[r].shape, src_np[r].shape, (axS,), [-1], prec)
            a = B.fftw.aligned(src_np[r].shape, dtype=dt)
            a[...] = src_np[r]
            sshape = src_np[r].shape
            if mode == 3:      # ranges of the last axis of a block whose FIRST axis is transformed (re-viewed rows)
                cuts = [0, 7, sshape[-1]]
                for lo, hi in zip(cuts[:-1], cuts[1:]):
                    plan.execute_scatter_chunk(device_ptr(a), 1.0, h, 0, ptrs, 0, 1, lo, hi - lo,
                                               view_outer=int(np.prod(sshape[1:-1])), view_ostride=sshape[-1])
            else:
                extent = int(np.prod(sshape[axS + 1:])) if mode == 1 else int(np.prod(sshape[:axS]))
                cuts = [0, extent // 2 + 1, extent]
                for lo, hi in zip(cuts[:-1], cuts[1:]):
                    plan.execute_scatter_chunk(device_ptr(a), 1.0, h, 0, ptrs, 0, mode, lo, hi - lo)
            plan.destroy()
            h.destroy()
        for r in range(p):
            err = np.abs(np.asarray(dst[r]) - expect[r]).max() / np.abs(full).max()
            assert err < tol, ('scatter', shape, axS, axD, p, mode, r, err)


# ---------------------------------------------------------------------------
# BASELINE config C2 at full size: properties that do not need a full oracle run
# ---------------------------------------------------------------------------
def _direct_dft_point(g, k):
    """one output point of the normalised 3-D DFT by three tensor contractions"""
    n0, n1, n2 = g.shape
    w0 = np.exp(-2j * np.pi * k[0] * np.arange(n0) / n0)
    w1 = np.exp(-2j * np.pi * k[1] * np.arange(n1) / n1)
    w2 = np.exp(-2j * np.pi * k[2] * np.arange(n2) / n2)
    return np.einsum('i,j,k,ijk->', w0, w1, w2, g, optimize=True) / g.size


def test_full_size_c2_properties(B):
    import torch
    N = 512
    shape = (N, N, N)
    fft = B.PFFT(B.COMM_WORLD, shape, dtype='D')
    gen = torch.Generator(device='cuda')
    gen.manual_seed(0)
    u = B.newDistArray(fft, False)
    u.tensor.copy_(torch.view_as_complex(torch.rand(shape + (2,), dtype=torch.float64, device='cuda', generator=gen)))
    uh = fft.forward(u)
    # (1) spot values against a direct evaluation of the definition (oracle, host)
    g = np.asarray(u)
    for k in [(0, 0, 0), (1, 2, 3), (511, 0, 256), (37, 400, 129)]:
        got = complex(uh.tensor[k].item())
        assert abs(got - _direct_dft_point(g, k)) < 1e-12, k
    # (2) Parseval: sum |u|^2 / N^3 == sum |u_hat|^2  (forward is normalised)
    e_phys = float((u.tensor.abs() ** 2).sum().item()) / g.size
    e_spec = float((uh.tensor.abs() ** 2).sum().item())
    assert abs(e_phys - e_spec) < 1e-10 * e_phys
    # (3) round trip
    back = B.newDistArray(fft, False)
    fft.backward(uh, back)
    assert float((back.tensor - u.tensor).abs().max().item()) < 1e-12
    # (4) linearity: F(a u + b v) == a F(u) + b F(v)
    v = B.newDistArray(fft, False)
    v.tensor.copy_(torch.view_as_complex(torch.rand(shape + (2,), dtype=torch.float64, device='cuda', generator=gen)))
    uh_copy = uh.tensor.clone()
    vh = fft.forward(v).tensor.clone()
    w = B.newDistArray(fft, False)
    w.tensor.copy_(2.5 * u.tensor - 0.5j * v.tensor)
    wh = fft.forward(w)
    assert float((wh.tensor - (2.5 * uh_copy - 0.5j * vh)).abs().max().item()) < 1e-12


@pytest.mark.parametrize('p', [2, 3, 4, 8, 16])
def test_flag_barrier_kernel_index_logic(B, p):
    """b2f_transfer_set_flags + b2f_transfer_exchange_p2p with every rank of a group played on
    one device, ONE AFTER THE OTHER (no rank ever waits for a kernel that has not been launched:
    the counters a rank polls are pre-set as if its peers had arrived).  Checks what the flag
    kernel publishes -- this rank's arrival count in slot [rank] of every peer's array, nothing
    in its own -- and that barrier / put / barrier delivers the blocks (the ordering MPI_Alltoallw's
    blocking semantics give the reference, pencil.py:182,200).  The cross-process behaviour is
    covered by tests/test_gpu_multi.py on real ranks."""
    import torch
    from mpi4py_fft_b200._lib import TransferHandle
    from mpi4py_fft_b200.devarray import device_ptr
    from mpi4py_fft_b200.pencil import _blockdist
    shape, axisA, axisB = (2 * p, 3 * p, 40), 1, 0
    g = rand(shape, 'D', 5)

    def blocks(axis_split):
        out = []
        for r in range(p):
            n, s0 = _blockdist(shape[axis_split], p, r)
            sl = [slice(None)] * len(shape)
            sl[axis_split] = slice(s0, s0 + n)
            out.append(np.ascontiguousarray(g[tuple(sl)]))
        return out
    A, Bx = blocks(axisB), blocks(axisA)
    flags = torch.zeros((p, 16), dtype=torch.int64, device='cuda')
    handles = []
    for rank in range(p):
        class FakeComm(object):
            ranks = tuple(range(p))
            _r = rank

            def Get_size(self):
                return p

            def Get_rank(self):
                return self._r
        h = TransferHandle(FakeComm(), shape, 16, A[rank].shape, axisA, Bx[rank].shape, axisB, exchange=False)
        h.set_flags([flags[j].data_ptr() for j in range(p)])
        handles.append(h)
    src = []
    for r in range(p):
        a = B.fftw.aligned(A[r].shape, dtype='D')
        a[...] = A[r]
        src.append(a)
    dst = [B.fftw.aligned(d.shape, dtype='D', fill=0) for d in Bx]
    ptrs = [device_ptr(d) for d in dst]
    big = 1 << 40
    for rounds in (1, 2):
        for r in range(p):
            flags[r, :] = big                      # "every peer has arrived" as far as rank r can tell
            torch.cuda.synchronize()
            handles[r].exchange_p2p(0, src[r], ptrs)          # flag barrier, put kernel, flag barrier
            torch.cuda.synchronize()
        f = flags.cpu().numpy()
        for j in range(p):
            for i in range(p):
                if i == j:
                    continue
                # slot [i] of rank j's array holds rank i's count: 2 barriers per exchange -- unless rank j
                # played after rank i and overwrote its own row with the stand-in value
                assert f[j, i] in (2 * rounds, big), (p, rounds, j, i, f[j, i])
        last = p - 1
        assert all(f[j, last] == 2 * rounds for j in range(p - 1))     # nobody played after the last rank
        for r in range(p):
            assert np.array_equal(np.asarray(dst[r]), Bx[r]), (p, r)
    for h in handles:
        h.destroy()


@pytest.mark.parametrize('shape,axis,dtype', [((96, 40, 24), 0, 'D'), ((20, 96, 24), 1, 'D'), ((20, 30, 96), 2, 'D'),
                                               ((20, 30, 96), 2, 'd'), ((40, 96, 6), 1, 'F'), ((12, 20, 192), 2, 'f'),
                                               ((384, 10, 8), 0, 'D'), ((6, 10, 128), 2, 'd'), ((24, 95, 8), 1, 'D'),
                                               ((128, 6, 10), 0, 'D'), ((6, 160, 10), 1, 'F'), ((4, 6, 1024), 2, 'D')])
def test_dealiasing_folded_into_the_transform(B, shape, axis, dtype, monkeypatch):
    """padded stages at any single-tile Stockham length (3 * 2^k for the 3/2 rule) run as ONE kernel: the forward transform's last pass
    writes only the kept modes, the backward transform's first pass reads them
    (b2f_plan_set_truncation) -- same values as the reference rule (libfft.py:263-311, restated in
    oracle/pfft_oracle.py and pinned on the reference's fixtures) and as the two-pass form."""
    import pfft_oracle as O
    from mpi4py_fft_b200 import _lib
    tol = 1e-12 if dtype in 'dD' else 1e-5
    pf = 1.5
    x = rand(shape, dtype, 31)
    f = B.FFT(shape, axes=(axis,), dtype=dtype, padding=pf)
    u = B.fftw.aligned(shape, dtype=dtype)
    u[...] = x
    n0 = _lib.launch_count()
    got = np.asarray(f.forward(u)).copy()
    launches = _lib.launch_count() - n0
    n = shape[axis]
    def has_kernel(m, families):      # 2^k, 3 * 2^k, 5 * 2^k, 7 * 2^k
        for f in families:
            if m % f == 0 and m // f >= 1 and (m // f) & (m // f - 1) == 0 and (f == 1 and m >= 2 or f > 1):
                return True
        return False
    # real transforms: every Stockham length of n / 2; c2c: the 2^k and 3 * 2^k families
    stockham = has_kernel(n // 2, (1, 3, 5, 7)) if (dtype in 'df' and n % 2 == 0) else has_kernel(n, (1, 3))
    fused = bool(f.forward._fused_plan())
    assert fused == stockham, (shape, axis, dtype)
    if fused:
        assert launches == 1, launches
    x64 = x.astype('D' if dtype in 'FD' else 'd')
    ref = O.padded_stage_forward(x64, axis, pf)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max())
    y = rand(ref.shape, ref.dtype, 32)
    v = B.fftw.aligned(ref.shape, dtype=got.dtype)
    v[...] = y
    back = np.asarray(f.backward(v)).copy()
    refb = O.padded_stage_backward(y.astype(got.dtype).astype(np.complex128), axis, n, dtype in 'df')
    assert np.abs(back - refb).max() <= tol * max(1.0, np.abs(refb).max())
    assert np.array_equal(np.asarray(v), y.astype(got.dtype))           # the truncated input survives
    # two-pass form (transform + b2f_pad_truncate): same values
    monkeypatch.setenv('B2F_FUSED_PAD', '0')
    f2 = B.FFT(shape, axes=(axis,), dtype=dtype, padding=pf)
    assert not f2.forward._fused_plan()
    got2 = np.asarray(f2.forward(u))
    assert np.abs(got2 - got).max() <= 10 * tol * max(1.0, np.abs(ref).max())
    f.destroy()
    f2.destroy()


def test_padded_pfft_uses_fused_stages(B):
    """PFFT(padding=1.5) on 64^3 -> 96^3: every stage is a padded 3 * 2^k stage, each one launch"""
    import pfft_oracle as O
    from mpi4py_fft_b200 import _lib
    shape = (64, 64, 64)
    fft = B.PFFT(B.COMM_WORLD, shape, dtype='D', padding=[1.5, 1.5, 1.5])
    u = B.newDistArray(fft, False)
    assert tuple(u.shape) == (96, 96, 96)
    x = rand((96, 96, 96), 'D', 8)
    u[...] = x
    n0 = _lib.launch_count()
    uh = fft.forward(u)
    assert _lib.launch_count() - n0 == 3
    ref = x
    for ax in (2, 1, 0):
        ref = O.padded_stage_forward(ref, ax, 1.5)
    assert np.abs(np.asarray(uh) - ref).max() < 1e-12
    n0 = _lib.launch_count()
    ub = fft.backward(uh)
    assert _lib.launch_count() - n0 == 3
    refb = ref
    for ax in (0, 1, 2):
        refb = O.padded_stage_backward(refb, ax, 96, False)
    assert np.abs(np.asarray(ub) - refb).max() < 1e-12
    fft.destroy()


def test_full_size_c3_closed_form_values(B):
    """BASELINE.json config 3 on one GPU (3D c2c 1024^3 complex128): EVERY point of the forward
    transform against a closed form -- the input is a sum of plane waves (bench.py plane_waves /
    fill_plane_waves, phases reduced exactly in integers), whose normalised spectrum is a_m at k_m
    and zero elsewhere -- then the round trip.  The same check runs inside bench.py at every N."""
    import importlib.util
    import os
    import torch
    from conftest import ROOT
    from mpi4py_fft_b200.devarray import as_tensor
    if torch.cuda.get_device_properties(0).total_memory < 100 * 2 ** 30:
        pytest.skip("needs ~70 GiB of HBM")
    spec = importlib.util.spec_from_file_location('b2f_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    shape = (1024, 1024, 1024)
    fft = B.PFFT(B.COMM_WORLD, shape, dtype='D')
    u = B.newDistArray(fft, False)
    ks, amps = bench.plane_waves(shape, False)
    bench.fill_plane_waves(torch, as_tensor(u), fft.local_slice(False), shape, False, ks, amps)
    keep = as_tensor(u)[:4].clone()
    uh = fft.forward(u)
    back = B.newDistArray(fft, False)
    fft.backward(uh, back)
    err = bench.spectrum_error(torch, as_tensor(uh), fft.local_slice(True), ks, amps, False)
    assert err < 1e-12, err
    assert float((as_tensor(back)[:4] - keep).abs().max().item()) < 1e-12
    assert float((as_tensor(back)[-4:] - as_tensor(u)[-4:]).abs().max().item()) < 1e-12
    fft.destroy()
