"""The oracle (oracle/pfft_oracle.py, oracle/dft_ref.c) pinned against the
reference: fixtures produced by the unmodified reference (tests/golden/, see
oracle/make_golden.py) and the known-answer vectors in the reference's
docstrings.  CPU only."""
import numpy as np
import pytest
import scipy.fft as sfft

import pfft_oracle as O
from conftest import case_kwargs

CASES = ['c1_c2c_16_p2', 'c3_c2c_16_p8_pencil', 'c3_c2c_16_p4_pencil', 'c4_r2c_16_p8_slab',
         'c4_r2c_16_p8_slab_collapse', 'c5_c2c_8x4_p8_grid42', 'uneven_r2c_12_13_14_p4',
         'uneven_c2c_13_12_11_p6_axes201', 'uneven_c2c_7_9_p3_2d', 'r2c_doc_128_p4_axes201',
         'c2c_4d_nested_p4', 'r2c_3d_nested_collapse_p4', 'c2c_32_p1',
         'pad_c2c_8_p4_3half', 'pad_r2c_8_12_10_p4_3half', 'pad_c2c_9_7_p2_mixed', 'pad_r2c_10_9_8_p1']
PADDED = [c for c in CASES if c.startswith('pad_')]
SERIAL_PAD = ['spad_c_even', 'spad_c_odd', 'spad_c_first_axis', 'spad_r_even_half', 'spad_r_odd_half',
              'spad_r_evenhalf2', 'spad_r_first_axis']


def make_oracle(meta):
    kw = case_kwargs(meta)
    return O.OraclePFFT(meta['nranks'], kw.pop('shape'), **kw)


@pytest.mark.parametrize('name', CASES)
def test_oracle_layout_bit_exact(layouts, name):
    """grid, per-stage pencils, per-transfer geometry and local slices of every
    rank equal the reference's (mpifft.py:259-337, pencil.py:277-354)"""
    case = layouts[name]
    orc = make_oracle(case['meta'])
    lay = orc.layout()
    for r, ref in enumerate(case['ranks']):
        assert lay['dims'] == ref['subcomm_sizes']
        assert lay['ranks'][r]['coords'] == ref['subcomm_ranks']
        assert lay['axes'] == ref['axes']
        assert lay['input_shape'] == ref['input_shape']
        assert lay['output_shape'] == ref['output_shape']
        assert lay['output_dtype'] == ref['dtype_out']
        assert [[s.start, s.stop] for s in orc.local_slice(r, False)] == ref['local_slice_in']
        assert [[s.start, s.stop] for s in orc.local_slice(r, True)] == ref['local_slice_out']
        for st, rst in zip(lay['ranks'][r]['stages'], ref['stages']):
            assert st['axes'] == rst['axes']
            assert st['in_subshape'] == rst['in_shape']
        assert len(lay['ranks'][r]['transfers']) == len(ref['transfers'])
        for tr, rtr in zip(lay['ranks'][r]['transfers'], ref['transfers']):
            for key in ('axisA', 'axisB', 'subshapeA', 'subshapeB', 'group_size', 'group_rank'):
                assert tr[key] == rtr[key], key


@pytest.mark.parametrize('name', [c for c in CASES if c not in ('c3_c2c_16_p4_pencil', 'r2c_doc_128_p4_axes201', 'c2c_32_p1')])
def test_oracle_values_match_reference(layouts, values, name):
    case = layouts[name]
    orc = make_oracle(case['meta'])
    g = values[name + '__input']
    fwd = orc.gather(orc.forward(orc.scatter(g)), True)
    ref = values[name + '__forward']
    tol = 1e-6 if g.dtype.char in 'fF' else 1e-14
    assert np.abs(fwd - ref).max() <= tol * max(1.0, np.abs(ref).max())
    bwd = orc.gather(orc.backward(orc.scatter(ref, True)), False)
    assert np.abs(bwd - values[name + '__backward']).max() <= 10 * tol
    if name in PADDED:
        return      # truncation is lossy: no round trip, no plain-DFT identity
    # and both agree with the transform of the undistributed array
    kw = case_kwargs(case['meta'])
    direct = O.expected_forward(g, kw.get('axes'))
    assert np.abs(fwd - direct).max() <= tol * max(1.0, np.abs(direct).max())
    assert np.abs(bwd - g).max() <= 10 * tol


@pytest.mark.parametrize('name', SERIAL_PAD)
def test_oracle_padded_stage_matches_reference(layouts, values, name):
    """the restated truncation / padding (libfft.py:263-311) against the reference's
    own libfft.FFT(padding=...) outputs, complex and real, odd and even extents"""
    meta = layouts['_' + name]
    x, y, z = values[name + '__input'], values[name + '__forward'], values[name + '__backward']
    got = O.padded_stage_forward(x, meta['axis'], meta['padding'])
    assert list(got.shape) == meta['trunc_shape']
    assert np.abs(got - y).max() <= 1e-14 * max(1.0, np.abs(y).max())
    back = O.padded_stage_backward(y, meta['axis'], meta['shape'][meta['axis']], meta['dtype'] in 'fd')
    assert np.abs(back - z).max() <= 1e-13 * max(1.0, np.abs(z).max())


def test_doc_layout_goldens(layouts):
    """pencil.py:55-62,254-263 and distarray.py:271-274 of the reference"""
    assert layouts['_doc_subcomm_p4'][0] == [2, 2, 1]
    assert layouts['_doc_subcomm_p6'][0] == [3, 2, 1]
    assert O.subcomm_dims(4, [0, 0, 1]) == [2, 2, 1]
    assert O.subcomm_dims(6, [0, 0, 1]) == [3, 2, 1]
    for r in range(4):
        assert layouts['_doc_pencil_8x4_p4'][r] == [[4, 4, 8, 8], [8, 4, 4, 8]]
    dims = O.subcomm_dims(4, [0, 0, 1, 0])
    for r in range(4):
        c = np.unravel_index(r, dims)
        p0 = O.VPencil(dims, c, (8, 8, 8, 8), 2)
        assert list(p0.subshape) == [4, 4, 8, 8] and list(p0.pencil(0).subshape) == [8, 4, 4, 8]
    expect = [[[0, 16], [0, 7], [0, 6]], [[0, 16], [0, 7], [6, 12]], [[0, 16], [7, 14], [0, 6]], [[0, 16], [7, 14], [6, 12]]]
    assert layouts['_doc_distarray_local_slice_p4'] == expect
    for key, val in layouts['_compute_dims'].items():
        n, d = key.split(':')
        assert O.compute_dims(int(n), [0] * int(d)) == val


def test_reference_docstring_vectors(dft_ref):
    """known answers printed in /root/reference/mpi4py_fft/fftw/xfftn.py
    (:85-88 fftn, :155-158 ifftn, :220-223 rfftn, :293-301 irfftn, :381-384 dctn,
    :453-456 idctn, :525-528 dstn, :597-600 idstn) -- checked for the scipy
    oracle and for the C restatement."""
    a = np.array([1, 2, 3, 4.0])
    assert np.allclose(O.serial_transform(a.astype(complex), (0,), [O.FORWARD]), [10, -2 + 2j, -2, -2 - 2j])
    assert np.allclose(dft_ref(-1, 4, a.astype(complex)), [10, -2 + 2j, -2, -2 - 2j])
    assert np.allclose(O.serial_transform(a.astype(complex), (0,), [O.BACKWARD]), [10, -2 - 2j, -2, -2 + 2j])
    assert np.allclose(dft_ref(1, 4, a.astype(complex)), [10, -2 - 2j, -2, -2 + 2j])
    assert np.allclose(O.serial_transform(a, (0,), [O.R2C]), [10, -2 + 2j, -2])
    assert np.allclose(dft_ref(-2, 4, a), [10, -2 + 2j, -2])
    c = np.array([1, 2, 3, 4], dtype=complex)
    # irfftn docstring (n = 6 default, and s = (7,))
    assert np.allclose(O.serial_c2r(c, (0,), (6,)), [15., -4., 0., -1., 0., -4.])
    assert np.allclose(dft_ref(2, 6, c), [15., -4., 0., -1., 0., -4.])
    r7 = O.serial_c2r(c, (0,), (7,))
    assert np.allclose(r7, [19., -5.04891734, -0.30797853, -0.64310413, -0.64310413, -0.30797853, -5.04891734])
    assert np.allclose(dft_ref(2, 7, c), r7)
    # dctn type 2 / idctn / dstn / idstn vectors
    assert np.allclose(O.serial_transform(a, (0,), [O.REDFT10]), [20., -6.30864406, 0., -0.44834153])
    assert np.allclose(dft_ref(O.REDFT10, 4, a), [20., -6.30864406, 0., -0.44834153])
    assert np.allclose(O.serial_transform(a, (0,), [O.REDFT01]), sfft.dct(a, type=3))
    assert np.allclose(O.serial_transform(a, (0,), [O.RODFT10]), sfft.dst(a, type=2))
    assert np.allclose(dft_ref(O.RODFT10, 4, a), sfft.dst(a, type=2))
    assert np.allclose(dft_ref(O.RODFT01, 4, a), sfft.dst(a, type=3))


@pytest.mark.parametrize('n', [2, 5, 8, 12, 13, 16])
def test_c_restatement_vs_scipy(dft_ref, n):
    """the reference pins r2r kinds against scipy (tests/test_fftw.py:106-117)"""
    rng = np.random.default_rng(n)
    x = rng.random(n)
    z = rng.random(n) + 1j * rng.random(n)
    assert np.allclose(dft_ref(-1, n, z), np.fft.fft(z), atol=1e-13)
    assert np.allclose(dft_ref(1, n, z), np.fft.ifft(z) * n, atol=1e-13)
    assert np.allclose(dft_ref(-2, n, x), np.fft.rfft(x), atol=1e-13)
    assert np.allclose(dft_ref(2, n, np.fft.rfft(x)), x * n, atol=1e-12)
    for kind, (fam, typ) in O._R2R_SCIPY.items():
        assert np.allclose(dft_ref(kind, n, x), getattr(sfft, fam)(x, type=typ), atol=1e-12), kind


def test_r2r_oracle_roundtrip_5d():
    """the reference's test_r2r case (tests/test_mpifft.py:35-51): 5-D, DCT-III on
    (1,2), DST-III on (3,4), Fourier on 0, slab; forward o backward == identity"""
    orc = O.OraclePFFT(4, (5, 6, 7, 8, 9), axes=((0,), (1, 2), (3, 4)), grid=(-1,), dtype='d',
                       transforms={(1, 2): ('dct', 3), (3, 4): ('dst', 3)})
    g = np.random.default_rng(0).random((5, 6, 7, 8, 9))
    f = orc.forward(orc.scatter(g))
    assert orc.output_shape == (3, 6, 7, 8, 9)
    b = orc.gather(orc.backward(f), False)
    assert np.abs(b - g).max() < 1e-13
    direct = O.expected_forward(g, ((0,), (1, 2), (3, 4)), {(1, 2): ('dct', 3), (3, 4): ('dst', 3)})
    assert np.abs(orc.gather(f, True) - direct).max() < 1e-14
