"""Serial transforms on the device (through the C ABI: b2f_planxfftn /
b2f_execute) against the oracle -- numpy/scipy pocketfft and the plain-C
restatement oracle/dft_ref.c.  Mirrors the reference's serial tests
(tests/test_fftw.py:32-138, tests/test_libfft.py:23-135): all planners, dims
1-3, every axis, f/d precision, r2r kinds 1-4, round trips -- plus exact value
checks and the power-of-two sizes the Stockham kernels serve.

Tolerances (north_star): 1e-12 (fp64) / 1e-5 (fp32) on transform values,
relative to max|reference| for unnormalised results.
"""
import functools
from itertools import product

import numpy as np
import pytest
import scipy.fft as sfft

import pfft_oracle as O

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


def test_docstring_vectors_on_device(B):
    """/root/reference/mpi4py_fft/fftw/xfftn.py:85-88,155-158,220-223,293-301,381-384"""
    fftw = B.fftw
    A = fftw.aligned(4, dtype='D')
    plan = fftw.fftn(A)
    A[:] = np.array([1, 2, 3, 4], dtype='D')
    assert np.allclose(np.asarray(plan()), [10, -2 + 2j, -2, -2 - 2j], atol=1e-14)
    assert plan.input_array is A and plan.output_array is plan()
    iplan = fftw.ifftn(A)
    assert np.allclose(np.asarray(iplan()), [10, -2 - 2j, -2, -2 + 2j], atol=1e-14)
    R = fftw.aligned(4, dtype='d')
    rplan = fftw.rfftn(R)
    R[:] = np.array([1., 2, 3, 4])
    assert np.allclose(np.asarray(rplan()), [10, -2 + 2j, -2], atol=1e-14)
    c2r = fftw.irfftn(A)
    assert np.allclose(np.asarray(c2r()), [15., -4., 0., -1., 0., -4.], atol=1e-13)
    c2r7 = fftw.irfftn(A, s=(7,))
    assert np.allclose(np.asarray(c2r7()), [19., -5.04891734, -0.30797853, -0.64310413, -0.64310413,
                                            -0.30797853, -5.04891734])
    dct = fftw.dctn(R)
    assert np.allclose(np.asarray(dct()), [20., -6.30864406, 0., -0.44834153])
    assert np.isclose(dct.get_normalization(), 1.0 / 8)


@pytest.mark.parametrize('n', [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192,
                               3, 6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 6144,
                               5, 10, 20, 40, 80, 160, 320, 640, 1280, 7, 14, 28, 56, 112, 224, 448, 896, 1792])
@pytest.mark.parametrize('dt', ['D', 'F'])
def test_stockham_pow2_c2c(B, n, dt):
    """contiguous axis, strided axes with ragged tiles, forward/backward, fused
    normalisation, in place -- every variant of fft_configs.h; lengths 2^k,
    3 * 2^k (radices 3, 6, 12, 24), 5 * 2^k (5, 10, 20) and 7 * 2^k (7, 14, 28)"""
    from mpi4py_fft_b200 import _lib
    tol = TOL[dt.lower()]
    shapes = [(5, n), (3, n, 7), (n, 33)] if n <= 2048 else [(2, n), (n, 5)]
    try:
        for variant in range(8):
            _lib.set_option('variant', variant)
            for shape in shapes:
                axis = shape.index(n)
                x = rand(shape, dt, seed=n + axis)
                U = B.fftw.aligned(shape, dtype=dt)
                U[...] = x
                fwd = B.fftw.fftn(U, axes=(axis,))
                y = fwd(normalize=True)
                ref = np.fft.fft(x.astype('D'), axis=axis) / n
                assert relerr(y, ref) < tol, ('fwd', n, dt, variant, shape)
                assert 'stockham' in fwd.plan().describe()
                bck = B.fftw.ifftn(fwd.output_array, axes=(axis,), output_array=U)
                z = bck()
                assert relerr(z, x.astype('D')) < tol, ('bwd', n, dt, variant, shape)
                # in place
                V = B.fftw.aligned(shape, dtype=dt)
                V[...] = x
                B.fftw.fftn(V, axes=(axis,), output_array=V)()
                assert relerr(V, ref * n) < tol, ('inplace', n, dt, variant, shape)
    finally:
        _lib.set_option('variant', 0)


@pytest.mark.parametrize('n', [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384,
                               6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 6144, 12288,
                               10, 20, 40, 80, 160, 320, 640, 1280, 2560, 14, 28, 56, 112, 224, 448, 896, 1792, 3584])
@pytest.mark.parametrize('dt', ['d', 'f'])
def test_stockham_real_transforms(B, n, dt):
    """r2c / c2r of even power-of-two length through the n/2-point Stockham kernel
    plus the split / merge pass (fft_real_kernel): contiguous and strided axis,
    ragged tiles, fused normalisation, multi-axis stage, and the dense-matrix path
    (real_engine=1) as a second witness on the same input"""
    from mpi4py_fft_b200 import _lib
    tol = TOL[dt]
    shapes = [(5, n), (3, n, 7), (n, 33)] if n <= 2048 else [(2, n), (n, 3)]
    for shape in shapes:
        axis = shape.index(n)
        x = rand(shape, dt, seed=n + axis)
        U = B.fftw.aligned(shape, dtype=dt)
        U[...] = x
        fwd = B.fftw.rfftn(U, axes=(axis,))
        assert 'stockham-real' in fwd.plan().describe()
        y = np.asarray(fwd(normalize=True)).copy()
        ref = np.fft.rfft(x.astype('d'), axis=axis) / n
        assert relerr(y, ref) < tol, ('r2c', n, dt, shape)
        bck = B.fftw.irfftn(fwd.output_array, s=(n,), axes=(axis,), output_array=U)
        z = bck()
        assert relerr(z, x.astype('d')) < tol, ('c2r', n, dt, shape)
        if n <= 512:
            _lib.set_option('real_engine', 1)
            _lib.set_option('generic_engine', 1)
            try:
                dense = B.fftw.rfftn(U, axes=(axis,))
                U[...] = x
                assert 'dense' in dense.plan().describe()
                assert relerr(dense(normalize=True), y) < tol
                _lib.set_option('generic_engine', 2)
                chirp = B.fftw.rfftn(U, axes=(axis,))
                U[...] = x
                assert 'chirpz' in chirp.plan().describe()
                assert relerr(chirp(normalize=True), y) < tol
            finally:
                _lib.set_option('real_engine', 0)
                _lib.set_option('generic_engine', 0)
    if n <= 1024:
        # two-axis stage: r2c along the last listed axis, then c2c; and back
        shape = (6, 16, n)
        x = rand(shape, dt, seed=3)
        U = B.fftw.aligned(shape, dtype=dt)
        U[...] = x
        fwd = B.fftw.rfftn(U, axes=(1, 2))
        y = fwd(normalize=True)
        assert relerr(y, np.fft.rfftn(x.astype('d'), axes=(1, 2)) / (16 * n)) < tol
        bck = B.fftw.irfftn(fwd.output_array, s=(16, n), axes=(1, 2), output_array=U)
        assert relerr(bck(), x.astype('d')) < tol


@pytest.mark.parametrize('n', [64, 128, 256, 512, 1024, 2048, 192, 384, 768])
@pytest.mark.parametrize('dt', ['D', 'F'])
def test_tma_staged_strided_c2c(B, n, dt):
    """strided axes through the TMA-staged persistent kernel (fft_tma.cuh), every
    variant: several tiles per CTA, ragged last tile, outer > 1, backward,
    fused normalisation, in place; then the automatic engine choice on a layout
    TMA cannot describe (odd inner extent of complex64: 8-byte row pitch)."""
    from mpi4py_fft_b200 import _lib
    tol = TOL[dt.lower()]
    big = 148 * 2 * 8 * 3 + 6 if n <= 512 else 148 * 8 + 6      # > one tile per persistent CTA
    shapes = [(n, big), (3, n, 40), (2, n, 26)]
    try:
        _lib.set_option('strided_engine', 2)
        _lib.set_option('variant_strict', 1)
        served = 0
        for variant in list(range(10)) + list(range(100, 108)):     # 100.. = cp.async-loaded flavour
            _lib.set_option('variant_tma', variant)
            for shape in shapes:
                axis = shape.index(n)
                x = rand(shape, dt, seed=n + axis)
                U = B.fftw.aligned(shape, dtype=dt)
                U[...] = x
                fwd = B.fftw.fftn(U, axes=(axis,))
                try:
                    y = fwd(normalize=True)
                except Exception as exc:
                    assert 'invalid' in str(exc).lower(), exc     # variant not built for this n
                    break
                ref = np.fft.fft(x.astype('D'), axis=axis) / n
                assert relerr(y, ref) < tol, ('fwd', n, dt, variant, shape)
                bck = B.fftw.ifftn(fwd.output_array, axes=(axis,), output_array=U)
                z = bck()
                assert relerr(z, x.astype('D')) < tol, ('bwd', n, dt, variant, shape)
                V = B.fftw.aligned(shape, dtype=dt)
                V[...] = x
                B.fftw.fftn(V, axes=(axis,), output_array=V)()
                assert relerr(V, ref * n) < tol, ('inplace', n, dt, variant, shape)
                served += 1
        assert served >= 3
        if n >= 256 and n & (n - 1) == 0:
            # the cp.async flavour also serves an odd inner extent (no descriptor involved)
            _lib.set_option('variant_tma', 102 if n == 256 else 100)     # a cp.async row that exists for this n
            shape = (2, n, 21)
            x = rand(shape, dt, seed=7)
            U = B.fftw.aligned(shape, dtype=dt)
            U[...] = x
            y = B.fftw.fftn(U, axes=(1,))()
            assert relerr(y, np.fft.fft(x.astype('D'), axis=1)) < tol
        _lib.set_option('variant_strict', 0)
        _lib.set_option('variant_tma', 0)
        _lib.set_option('strided_engine', 0)
        shape = (n, 33)
        x = rand(shape, dt, seed=1)
        U = B.fftw.aligned(shape, dtype=dt)
        U[...] = x
        y = B.fftw.fftn(U, axes=(0,))()
        assert relerr(y, np.fft.fft(x.astype('D'), axis=0)) < tol
    finally:
        _lib.set_option('variant_strict', 0)
        _lib.set_option('variant_tma', 0)
        _lib.set_option('strided_engine', 0)


@pytest.mark.parametrize('dt', ['d', 'f'])
def test_all_planners_small_sizes(B, dt):
    """the reference's own matrix: dims 1-3, sizes (7, 8, 10), every axes
    combination (tests/test_fftw.py:36-103) -- values against numpy, and the
    r2c->c2r / c2c round trips the reference checks"""
    fftw = B.fftw
    tol = TOL[dt]
    for dim in (1, 2, 3):
        for shape in product(*([(7, 8, 10)] * dim)):
            if dim == 3 and shape[0] != 7:
                continue
            allaxes = tuple(reversed(range(dim)))
            for i in range(dim):
                axes = tuple(reversed(allaxes[:i + 1]))
                # r2c -> c2r
                x = rand(shape, dt, 1)
                A = fftw.aligned(shape, dtype=dt)
                A[...] = x
                r2c = fftw.rfftn(A, axes=axes)
                Bh = r2c()
                ref = np.fft.rfftn(x.astype('d'), axes=axes)
                assert relerr(Bh, ref) < tol, ('r2c', shape, axes)
                c2r = fftw.irfftn(r2c.output_array, s=np.take(shape, axes), axes=axes, output_array=A)
                back = c2r(normalize=True)
                assert relerr(back, x.astype('d')) < 10 * tol, ('c2r', shape, axes)
                # c2c
                z = rand(shape, dt.upper(), 2)
                C = fftw.aligned(shape, dtype=dt.upper())
                C[...] = z
                c2c = fftw.fftn(C, axes=axes)
                D = c2c()
                assert relerr(D, np.fft.fftn(z.astype('D'), axes=axes)) < tol, ('c2c', shape, axes)
                ic2c = fftw.ifftn(c2c.output_array, axes=axes, output_array=C)
                assert relerr(ic2c(normalize=True), z.astype('D')) < 10 * tol


@pytest.mark.parametrize('dt', ['d', 'f'])
def test_r2r_all_kinds(B, dt, dft_ref):
    """DCT/DST types 1-4 against scipy (as tests/test_fftw.py:106-138) and
    against the C restatement of FFTW's definitions; mixed kinds per axis"""
    fftw = B.fftw
    tol = 10 * TOL[dt]
    for n in (5, 8, 13):
        x = rand((n,), dt, n)
        A = fftw.aligned((n,), dtype=dt)
        for typ in (1, 2, 3, 4):
            for fam, planner, iplanner in (('dct', fftw.dctn, fftw.idctn), ('dst', fftw.dstn, fftw.idstn)):
                A[...] = x
                p = planner(A, type=typ)
                y = np.asarray(p())
                assert relerr(y, getattr(sfft, fam)(x.astype('d'), type=typ)) < tol, (fam, typ, n)
                kind = (fftw.dct_type if fam == 'dct' else fftw.dst_type)[typ]
                assert relerr(y, dft_ref(kind, n, x.astype('d'))) < tol
                ip = iplanner(p.output_array, type=typ, output_array=A)
                assert relerr(ip(normalize=True), x.astype('d')) < tol, ('inv', fam, typ, n)
    # several axes, a different kind on each (tests/test_fftw.py:119-133)
    shape = (6, 7, 5)
    x = rand(shape, dt, 3)
    kinds = [fftw.FFTW_REDFT10, fftw.FFTW_RODFT01, fftw.FFTW_REDFT00]
    A = fftw.aligned(shape, dtype=dt)
    A[...] = x
    out = fftw.aligned(shape, dtype=dt)
    plan = fftw.FFT(A, out, axes=(0, 1, 2), kind=kinds, normalization=fftw.get_normalization(kinds, shape, (0, 1, 2)))
    y = plan()
    ref = O.serial_transform(x.astype('d'), (0, 1, 2), kinds)
    assert relerr(y, ref) < tol
    inv = fftw.FFT(out, A, axes=(0, 1, 2), kind=[fftw.inverse[k] for k in kinds],
                   normalization=plan.get_normalization())
    assert relerr(inv(normalize=True), x.astype('d')) < tol


@pytest.mark.parametrize('n', [3, 5, 6, 7, 9, 12, 13, 24, 100])
def test_generic_lengths_c2c(B, n, dft_ref):
    """non powers of two (the reference's tests use 5..13) on every axis"""
    for shape, axis in (((4, n), 1), ((n, 6), 0), ((3, n, 5), 1)):
        z = rand(shape, 'D', n)
        U = B.fftw.aligned(shape, dtype='D')
        U[...] = z
        p = B.fftw.fftn(U, axes=(axis,))
        y = np.asarray(p())
        assert relerr(y, np.fft.fft(z, axis=axis)) < 1e-12
        expect = 'stockham' if n in (3, 5, 6, 7, 12, 24) else 'dense-matrix' if n <= 32 else 'chirpz'
        assert expect in p.plan().describe()
    z = rand((n,), 'D', 1)
    U = B.fftw.aligned((n,), dtype='D')
    U[...] = z
    assert relerr(B.fftw.fftn(U)(), dft_ref(-1, n, z)) < 1e-12


@pytest.mark.parametrize('dt', ['d', 'f'])
def test_chirpz_every_kind_any_length(B, dt, dft_ref):
    """chirp-z kernels forced for every length (generic_engine=2): c2c both signs,
    r2c, c2r (odd and even lengths), the eight r2r kinds, contiguous and strided,
    against numpy / scipy, the O(n^2) restatement of the FFTW definitions and the
    dense-matrix path; plus the sizes a 3/2-rule padded solver uses"""
    import scipy.fft as sfft
    from mpi4py_fft_b200 import _lib
    fftw = B.fftw
    tol = TOL[dt] * (5 if dt == 'f' else 1)
    cd = dt.upper()
    _lib.set_option('generic_engine', 2)
    _lib.set_option('stockham', 0)          # 6, 12, 384, 1536 have Stockham kernels: keep them on the chirp-z path here
    try:
        for n in (5, 6, 7, 12, 13, 30, 100, 127, 384, 1000, 1536):
            for shape, axis in (((3, n), 1), ((n, 5), 0)):
                z = rand(shape, cd, n)
                U = fftw.aligned(shape, dtype=cd)
                U[...] = z
                pf = fftw.fftn(U, axes=(axis,))
                assert 'chirpz' in pf.plan().describe()
                y = np.asarray(pf()).copy()
                assert relerr(y, np.fft.fft(z.astype('D'), axis=axis)) < tol, ('c2c', n, shape)
                pb = fftw.ifftn(pf.output_array, axes=(axis,), output_array=U)
                assert relerr(pb(normalize=True), z.astype('D')) < tol, ('c2c inverse', n, shape)
                x = rand(shape, dt, n + 1)
                R = fftw.aligned(shape, dtype=dt)
                R[...] = x
                pr = fftw.rfftn(R, axes=(axis,))
                assert 'chirpz' in pr.plan().describe()
                assert relerr(pr(), np.fft.rfft(x.astype('d'), axis=axis)) < tol, ('r2c', n, shape)
                pc = fftw.irfftn(pr.output_array, s=(n,), axes=(axis,), output_array=R)
                assert relerr(pc(normalize=True), x.astype('d')) < tol, ('c2r', n, shape)
                if n <= 384:
                    for typ in (1, 2, 3, 4):
                        for fam, planner in (('dct', fftw.dctn), ('dst', fftw.dstn)):
                            R[...] = x
                            pp = planner(R, axes=(axis,), type=typ)
                            assert 'chirpz' in pp.plan().describe()
                            ref = getattr(sfft, fam)(x.astype('d'), type=typ, axis=axis)
                            assert relerr(pp(), ref) < tol, (fam, typ, n, shape)
        # definitions (long double, O(n^2)) and the dense path as second witnesses
        n = 45
        x = rand((n,), dt, 7)
        R = fftw.aligned((n,), dtype=dt)
        for kind in range(3, 11):
            R[...] = x
            out = fftw.aligned((n,), dtype=dt)
            got = np.asarray(fftw.FFT(R, out, axes=(0,), kind=kind)()).copy()
            assert relerr(got, dft_ref(kind, n, x.astype('d'))) < tol, kind
            _lib.set_option('generic_engine', 1)
            dense = fftw.FFT(R, out, axes=(0,), kind=kind)
            assert 'dense' in dense.plan().describe()
            assert relerr(dense(), got) < tol, kind
            _lib.set_option('generic_engine', 2)
    finally:
        _lib.set_option('generic_engine', 0)
        _lib.set_option('stockham', 1)


@pytest.mark.parametrize('backendless', [True])
def test_libfft_class(B, backendless):
    """libfft.FFT as in the reference's tests/test_libfft.py: forward normalised,
    backward not, round trip; given arrays are used directly"""
    from mpi4py_fft_b200.libfft import FFT
    for dt in 'fFdD':
        tol = TOL[dt.lower()]
        for dim in (1, 2, 3):
            for shape in product(*([(7, 8, 9)] * dim)):
                if dim == 3 and shape[1] != 8:
                    continue
                for axes in [None, (dim - 1,), tuple(range(dim))]:
                    fft = FFT(shape, axes, dtype=dt)
                    x = rand(shape, dt, 5)
                    A = fft.forward.input_array
                    A[...] = x
                    Bh = fft.forward()
                    ax = tuple(range(dim)) if axes is None else axes
                    x64 = x.astype(dt.upper() if dt in 'FD' else 'd').astype('D' if dt in 'FD' else 'd')
                    ref = (np.fft.fftn(x64, axes=ax) if dt in 'FD' else np.fft.rfftn(x64, axes=ax)) / np.prod(np.take(shape, ax))
                    assert relerr(Bh, ref) < tol, (dt, shape, axes)
                    C = fft.backward()
                    assert relerr(C, x64) < 10 * tol
                    # explicit arrays: used in place of the owned ones
                    A2 = B.fftw.aligned_like(A)
                    A2[...] = x
                    out = B.fftw.aligned_like(Bh)
                    r = fft.forward(A2, out)
                    assert r is out and relerr(out, ref) < tol
                    assert np.array_equal(np.asarray(A2), x)       # input preserved
                    # host arrays are staged
                    h = np.zeros(Bh.shape, dtype=Bh.dtype)
                    fft.forward(x, h)
                    assert relerr(h, ref) < tol
                    # normalize flags swapped (tests/test_mpifft.py:246-251 idea)
                    un = fft.forward(A2, normalize=False)
                    assert relerr(np.asarray(un) / np.prod(np.take(shape, ax)), ref) < tol
                    fft.destroy()


def test_errors(B):
    from mpi4py_fft_b200._lib import B200FFTError
    with pytest.raises((B200FFTError, RuntimeError)):
        U = B.fftw.aligned((3, 5000), dtype='D')     # non-pow2 beyond the dense-matrix limit
        B.fftw.fftn(U, axes=(1,))()
    with pytest.raises(NotImplementedError):
        B.fftw.hfftn(None)
    with pytest.raises(AssertionError):
        U = B.fftw.aligned((4, 4), dtype='d')
        B.fftw.fftn(U)


@pytest.mark.parametrize('n', [4, 8, 64, 256, 1024, 4096, 96, 768, 160, 1280, 112, 896])
@pytest.mark.parametrize('dt', ['d', 'f'])
def test_stockham_r2r_kinds_2_3_4(B, n, dt):
    """DCT / DST of kinds II, III and IV (FFTW_REDFT10 / REDFT01 / REDFT11 and the RODFT ones,
    /root/reference/mpi4py_fft/fftw/xfftn.py:14-36) of even length as Stockham transforms
    (II / III: Makhoul permutation + n/2-point complex schedule + quarter-wave twiddle; IV: n/2-point
    transform of pre-twiddled pairs) against scipy:
    contiguous and strided axes, ragged tiles, in place, inverse pairs, and the chirp-z kernels
    (r2r_engine=1) as a second witness on the same input"""
    from mpi4py_fft_b200 import _lib
    fftw = B.fftw
    tol = TOL[dt] * (10 if dt == 'f' else 1) * max(1.0, np.log2(n) / 4)
    shapes = [(5, n), (3, n, 7), (n, 33)] if n <= 1024 else [(2, n), (n, 3)]
    for shape in shapes:
        axis = shape.index(n)
        x = rand(shape, dt, seed=n + axis)
        for typ in (2, 3, 4):
            for fam, planner, iplanner in (('dct', fftw.dctn, fftw.idctn), ('dst', fftw.dstn, fftw.idstn)):
                A = fftw.aligned(shape, dtype=dt)
                A[...] = x
                p = planner(A, axes=(axis,), type=typ)
                assert 'stockham-r2r' in p.plan().describe(), (n, fam, typ)
                y = np.asarray(p()).copy()
                ref = getattr(sfft, fam)(x.astype('d'), type=typ, axis=axis)
                assert relerr(y, ref) < tol, (fam, typ, n, shape)
                ip = iplanner(p.output_array, axes=(axis,), type=typ, output_array=A)
                assert relerr(ip(normalize=True), x.astype('d')) < tol, ('inverse', fam, typ, n, shape)
                # in place
                A[...] = x
                q = planner(A, axes=(axis,), type=typ, output_array=A)
                assert relerr(q(), ref) < tol, ('in place', fam, typ, n, shape)
                if n <= 1024 and shape is shapes[0]:
                    _lib.set_option('r2r_engine', 1)
                    try:
                        A[...] = x
                        c = planner(A, axes=(axis,), type=typ)
                        assert 'stockham' not in c.plan().describe()
                        assert relerr(c(), y) < tol
                    finally:
                        _lib.set_option('r2r_engine', 0)


@pytest.mark.parametrize('N', [4, 8, 64, 256, 1024, 96, 160, 112])
@pytest.mark.parametrize('dt', ['d', 'f'])
def test_stockham_r2r_kinds_1(B, N, dt):
    """DCT-I of N + 1 points (the Chebyshev grids 2^k + 1) and DST-I of N - 1 points (FFTW_REDFT00 / RODFT00)
    as the real transform of the even / odd extension of length 2N, read through an index map"""
    fftw = B.fftw
    tol = TOL[dt] * (10 if dt == 'f' else 1) * max(1.0, np.log2(N) / 4)
    for fam, planner, iplanner, n in (('dct', fftw.dctn, fftw.idctn, N + 1), ('dst', fftw.dstn, fftw.idstn, N - 1)):
        for shape in ((5, n), (3, n, 7), (n, 9)):
            axis = shape.index(n)
            x = rand(shape, dt, seed=n + axis)
            A = fftw.aligned(shape, dtype=dt)
            A[...] = x
            p = planner(A, axes=(axis,), type=1)
            assert 'stockham-r2r' in p.plan().describe(), (n, fam)
            y = np.asarray(p()).copy()
            ref = getattr(sfft, fam)(x.astype('d'), type=1, axis=axis)
            assert relerr(y, ref) < tol, (fam, n, shape)
            ip = iplanner(p.output_array, axes=(axis,), type=1, output_array=A)
            assert relerr(ip(normalize=True), x.astype('d')) < tol, ('inverse', fam, n, shape)


@pytest.mark.parametrize('n', [16384, 32768, 65536, 12288, 10240, 7168, 1048576])
@pytest.mark.parametrize('dt', ['D', 'F'])
def test_fourstep_lengths_beyond_one_tile(B, n, dt):
    """c2c lengths without a single-tile kernel (2^k > 8192, other lengths > 4096): the four-step split
    n = n1 * n2 (csrc/lengths.h) -- strided n2-point transforms, twiddle, n1-point transforms stored
    transposed (the rotating kernel on the contiguous axis) -- against numpy; FFTW plans any N
    (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:49-56)"""
    tol = TOL[dt.lower()] * 4
    shapes = [(2, n), (n, 3), (2, n, 2)] if n <= 65536 else [(1, n)]
    for shape in shapes:
        axis = shape.index(n)
        x = rand(shape, dt, seed=axis)
        U = B.fftw.aligned(shape, dtype=dt)
        U[...] = x
        fwd = B.fftw.fftn(U, axes=(axis,))
        assert 'stockham-fourstep' in fwd.plan().describe()
        y = np.asarray(fwd(normalize=True)).copy()
        ref = np.fft.fft(x.astype('D'), axis=axis) / n
        assert relerr(y, ref) < tol, ('fwd', n, dt, shape)
        assert np.array_equal(np.asarray(U), x)
        bck = B.fftw.ifftn(fwd.output_array, axes=(axis,), output_array=U)
        assert relerr(bck(), x.astype('D')) < tol, ('bwd', n, dt, shape)
        V = B.fftw.aligned(shape, dtype=dt)
        V[...] = x
        B.fftw.fftn(V, axes=(axis,), output_array=V)()
        assert relerr(V, ref * n) < tol, ('in place', n, dt, shape)
    if n == 16384:
        # inside a multi-axis stage
        shape = (n, 64)
        x = rand(shape, dt, seed=9)
        U = B.fftw.aligned(shape, dtype=dt)
        U[...] = x
        y = B.fftw.fftn(U, axes=(0, 1))()
        assert relerr(y, np.fft.fftn(x.astype('D'))) < tol
