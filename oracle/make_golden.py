"""Generate tests/golden/* by running the UNMODIFIED reference (staged under
oracle/_ref by oracle/build_ref.py) on virtual ranks (oracle/fakempi).

    python oracle/make_golden.py          # build container only

Two fixture files:
  layouts.json  -- for each case and each rank, everything the index maps
                   decide: process grid, per-stage pencil subshape/substart/axis,
                   per-transfer subshapes/axes/group size and rank, local
                   slices, global shapes, dtypes.  Bit-exact contract.
  values.npz    -- for the small cases, the seeded global input, the gathered
                   forward output of the reference and its backward result.
The reference runs with its own ``numpy`` serial backend (FFTW is not
installable here; see oracle/build_ref.py), so "values" pin the reference's
orchestration + pocketfft arithmetic; FFTW agrees with pocketfft to ~1e-16.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'fakempi'))
sys.path.insert(0, os.path.join(HERE, '_ref'))

from mpi4py import MPI                                    # noqa: E402  (the fake)
from mpi4py_fft import PFFT, newDistArray, DistArray     # noqa: E402  (the reference)
from mpi4py_fft.pencil import Subcomm, Pencil            # noqa: E402
import scipy.fft as sfft                                  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# name -> (nranks, kwargs of PFFT, save values?)
CASES = {
    'c1_c2c_16_p2':        (2, dict(shape=(16, 16, 16), dtype='D'), True),
    'c3_c2c_16_p8_pencil': (8, dict(shape=(16, 16, 16), dtype='D'), True),
    'c3_c2c_16_p4_pencil': (4, dict(shape=(16, 16, 16), dtype='D'), False),
    'c4_r2c_16_p8_slab':   (8, dict(shape=(16, 16, 16), dtype='f', grid=(-1,)), True),
    'c4_r2c_16_p8_slab_collapse': (8, dict(shape=(16, 16, 16), dtype='d', grid=(-1,), collapse=True), True),
    'c5_c2c_8x4_p8_grid42': (8, dict(shape=(8, 8, 8, 8), dtype='D', grid=(4, 2)), True),
    'uneven_r2c_12_13_14_p4': (4, dict(shape=(12, 13, 14), dtype='d'), True),
    'uneven_c2c_13_12_11_p6_axes201': (6, dict(shape=(13, 12, 11), dtype='D', axes=(2, 0, 1)), True),
    'uneven_c2c_7_9_p3_2d': (3, dict(shape=(7, 9), dtype='D'), True),
    'r2c_doc_128_p4_axes201': (4, dict(shape=(128, 128, 128), dtype='d', axes=(2, 0, 1)), False),
    'c2c_4d_nested_p4': (4, dict(shape=(6, 8, 5, 7), dtype='D', axes=((0,), (1,), (2, 3))), True),
    'r2c_3d_nested_collapse_p4': (4, dict(shape=(12, 13, 8), dtype='d', axes=((0,), (1, 2)), collapse=True), True),
    'c2c_32_p1': (1, dict(shape=(32, 32, 32), dtype='D'), False),
    # padded (dealiased) transforms, mpifft.py:247-253 + libfft.py:263-311; shape = the truncated size
    'pad_c2c_8_p4_3half': (4, dict(shape=(8, 8, 8), dtype='D', padding=[1.5, 1.5, 1.5]), True),
    'pad_r2c_8_12_10_p4_3half': (4, dict(shape=(8, 12, 10), dtype='d', padding=[1.5, 1.5, 1.5]), True),
    'pad_c2c_9_7_p2_mixed': (2, dict(shape=(9, 7), dtype='D', padding=[2, 1.5]), True),
    'pad_r2c_10_9_8_p1': (1, dict(shape=(10, 9, 8), dtype='d', padding=[1.5, 1.0, 1.5]), True),
}

# serial padded stages (libfft.FFT with padding): shape (already padded), axis, dtype, factor
SERIAL_PAD = {
    'spad_c_even': ((5, 12, 3), 1, 'D', 1.5),
    'spad_c_odd': ((4, 15), 1, 'D', 1.5),
    'spad_c_first_axis': ((18, 4), 0, 'D', 2.0),
    'spad_r_even_half': ((3, 12), 1, 'd', 1.5),     # kept 8 -> half spectrum 5 (odd)
    'spad_r_odd_half': ((3, 18), 1, 'd', 1.5),      # kept 12 -> half spectrum 7
    'spad_r_evenhalf2': ((2, 9, 4), 1, 'd', 1.5),   # kept 6 -> half spectrum 4 (even: Nyquist rule fires)
    'spad_r_first_axis': ((15, 4), 0, 'd', 1.5),    # kept 10 -> half spectrum 6 (even)
}


# r2r stages cannot be generated from the reference here: its numpy backend
# casts every stage output to complex (libfft.py:94) and the FFTW backend is
# not buildable.  r2r values are pinned against scipy instead, exactly as the
# reference's own tests do (tests/test_fftw.py:106-117).


def rank_body(kw, want_values, seed):
    comm = MPI.COMM_WORLD
    kw = dict(kw)
    if 'padding' in kw:
        kw['padding'] = list(kw['padding'])     # the reference rewrites the list in place (mpifft.py:253)
    fft = PFFT(comm, backend='numpy', **kw)
    info = dict(
        subcomm_sizes=[c.Get_size() for c in fft.subcomm],
        subcomm_ranks=[c.Get_rank() for c in fft.subcomm],
        axes=[list(a) for a in fft.axes],
        input_shape=list(fft.global_shape(False)), output_shape=list(fft.global_shape(True)),
        local_shape_in=list(fft.shape(False)), local_shape_out=list(fft.shape(True)),
        local_slice_in=[[s.start, s.stop] for s in fft.local_slice(False)],
        local_slice_out=[[s.start, s.stop] for s in fft.local_slice(True)],
        dtype_in=fft.dtype(False).char, dtype_out=fft.dtype(True).char,
        stages=[dict(axes=list(x.axes),
                     in_shape=list(x.forward.input_array.shape), in_dtype=x.forward.input_array.dtype.char,
                     out_shape=list(x.forward.output_array.shape), out_dtype=x.forward.output_array.dtype.char)
                for x in fft.xfftn],
        transfers=[dict(axisA=t.axisA, axisB=t.axisB, subshapeA=list(t.subshapeA), subshapeB=list(t.subshapeB),
                        shape=list(t.shape), group_size=t.comm.Get_size(), group_rank=t.comm.Get_rank(),
                        dtype=t.dtype.char) for t in fft.transfer],
        pencil_in=dict(subshape=list(fft.pencil[0].subshape), substart=list(fft.pencil[0].substart), axis=fft.pencil[0].axis),
        pencil_out=dict(subshape=list(fft.pencil[1].subshape), substart=list(fft.pencil[1].substart), axis=fft.pencil[1].axis),
    )
    vals = None
    if want_values:
        rng = np.random.default_rng(seed)
        g = rng.random(fft.global_shape(False))
        if fft.dtype(False).char in 'FD':
            g = g + 1j * rng.random(fft.global_shape(False))
        g = g.astype(fft.dtype(False))
        u = newDistArray(fft, False)
        u[:] = g[fft.local_slice(False)]
        uh = fft.forward(u).copy()
        ub = fft.backward(uh.copy()).copy()
        vals = (g, fft.local_slice(True), np.array(uh), fft.local_slice(False), np.array(ub))
    return info, vals


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    layouts, values = {}, {}
    for seed, (name, (nranks, kw, want)) in enumerate(sorted(CASES.items())):
        res = MPI.run_ranks(nranks, rank_body, kw, want, seed)
        meta = dict(nranks=nranks, kwargs={k: (v if not isinstance(v, dict) else {str(a): list(b) for a, b in v.items()})
                                           for k, v in kw.items()}, seed=seed)
        layouts[name] = dict(meta=meta, ranks=[r[0] for r in res])
        if want:
            g = res[0][1][0]
            out_shape = res[0][0]['output_shape']
            fwd = np.zeros(out_shape, dtype=res[0][1][2].dtype)
            bwd = np.zeros(g.shape, dtype=res[0][1][4].dtype)
            for info, (gg, sl_out, uh, sl_in, ub) in res:
                fwd[sl_out] = uh
                bwd[sl_in] = ub
            values[name + '__input'] = g
            values[name + '__forward'] = fwd
            values[name + '__backward'] = bwd
        print('%-36s ranks=%d grid=%s out=%s' % (name, nranks, res[0][0]['subcomm_sizes'], res[0][0]['output_shape']))

    # serial padded stages through the reference's own libfft.FFT (numpy backend)
    from mpi4py_fft.libfft import FFT as RefFFT
    for seed, (name, (shape, axis, dt, pf)) in enumerate(sorted(SERIAL_PAD.items())):
        f = RefFFT(shape, axes=(axis,), dtype=dt, padding=pf, backend='numpy')
        rng = np.random.default_rng(100 + seed)
        x = rng.random(shape)
        if dt in 'FD':
            x = x + 1j * rng.random(shape)
        x = x.astype(dt)
        f.forward.input_array[...] = x
        y = np.array(f.forward()).copy()
        f.backward.input_array[...] = y
        z = np.array(f.backward()).copy()
        values[name + '__input'] = x
        values[name + '__forward'] = y
        values[name + '__backward'] = z
        layouts['_' + name] = dict(shape=list(shape), axis=axis, dtype=dt, padding=pf,
                                   trunc_shape=list(y.shape), trunc_dtype=y.dtype.char)
        print('%-36s %s -> %s' % (name, shape, y.shape))

    # stand-alone layout goldens quoted in the reference's docstrings
    def pencil_doc():
        s = Subcomm(MPI.COMM_WORLD, [0, 0, 1, 0])
        p0 = Pencil(s, (8, 8, 8, 8), 2)
        p1 = p0.pencil(0)
        return list(p0.subshape), list(p1.subshape)

    def subcomm_doc():
        return [c.Get_size() for c in Subcomm(MPI.COMM_WORLD, [0, 0, 1])]

    def distarray_doc():
        z = DistArray((16, 14, 12), dtype=float, alignment=0)
        return [[s.start, s.stop] for s in z.local_slice()]

    layouts['_doc_pencil_8x4_p4'] = MPI.run_ranks(4, pencil_doc)
    layouts['_doc_subcomm_p4'] = MPI.run_ranks(4, subcomm_doc)
    layouts['_doc_subcomm_p6'] = MPI.run_ranks(6, subcomm_doc)
    layouts['_doc_distarray_local_slice_p4'] = MPI.run_ranks(4, distarray_doc)
    layouts['_compute_dims'] = {str(n) + ':' + str(d): MPI.Compute_dims(n, d)
                                for n in (1, 2, 3, 4, 6, 8, 12, 16, 24) for d in (1, 2, 3)}

    with open(os.path.join(GOLDEN, 'layouts.json'), 'w') as f:
        json.dump(layouts, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(GOLDEN, 'values.npz'), **values)
    print('wrote', GOLDEN)


if __name__ == '__main__':
    main()
