"""The rotating schedule of a 3-axis c2c stage (csrc/fft_rot.cuh, rot_plan.h) on the
device through the C ABI, against numpy's pocketfft: every variant of
B2F_ROT_TABLE, batches, ragged tiles, forward / backward, the merged PFFT path
(mpifft.Transform._merge) and its equality with the stage-by-stage chain.
Replaces fftw_execute_dft of a multi-axis guru plan
(/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:52-56).
Tolerances: 1e-12 (fp64) / 1e-5 (fp32) relative to max|reference|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {'D': 1e-12, 'F': 1e-5}


def rand(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.random(shape) + 1j * rng.random(shape)).astype(dtype)


def relerr(a, ref):
    return np.abs(np.asarray(a) - ref).max() / np.abs(ref).max()


@pytest.fixture(scope='module')
def B():
    """the rotating schedule is opt-in (option "rotate" / B2F_ROTATE=1): on 1024^3 complex128 it ties
    with the per-axis kernels in short runs and loses under the power cap (profiles/r2_rot_*.txt)"""
    import torch
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200 import _lib
    torch.cuda.set_device(0)
    _lib.set_option('rotate', 1)
    yield B
    _lib.set_option('rotate', 0)


@pytest.mark.parametrize('shape', [(64, 64, 64), (128, 64, 256), (2, 64, 128, 64), (72, 64, 64)])
@pytest.mark.parametrize('dtype', ['D', 'F'])
def test_rotating_schedule_values(B, shape, dtype):
    from mpi4py_fft_b200 import _lib
    x = rand(shape, dtype)
    nd = len(shape)
    axes = tuple(range(nd - 3, nd))
    a = B.fftw.aligned(shape, dtype=dtype)
    b = B.fftw.aligned(shape, dtype=dtype)
    a[...] = x
    fwd = B.fftw.fftn(a, axes=axes, output_array=b)
    rotates = 'rotating' in fwd.plan().describe()
    assert rotates == all(n in (64, 128, 256, 512, 1024, 2048) for n in shape[-3:])
    x64 = x.astype('D')
    n0 = _lib.launch_count()
    y = np.asarray(fwd(a, b))
    if rotates:
        assert _lib.launch_count() - n0 == 3
    assert relerr(y, np.fft.fftn(x64, axes=axes)) < TOL[dtype]
    assert np.array_equal(np.asarray(a), x)                       # the input survives
    bwd = B.fftw.ifftn(a, axes=axes, output_array=b)
    y = np.asarray(bwd(a, b, normalize=True))
    assert relerr(y, np.fft.ifftn(x64, axes=axes)) < TOL[dtype]
    # in place: the classic per-axis schedule takes over, same values
    y = np.asarray(fwd(a, a))
    assert relerr(y, np.fft.fftn(x64, axes=axes)) < TOL[dtype]
    fwd.destroy()
    bwd.destroy()


@pytest.mark.parametrize('n', [64, 128, 256, 512, 1024])
def test_rotating_variants(B, n):
    """every row of B2F_ROT_TABLE for one length, other extents small and ragged (I = 20 is no tile multiple)"""
    from mpi4py_fft_b200 import _lib
    for dtype in ('D', 'F'):
        shape = (64, 64, n)
        x = rand(shape, dtype, seed=n)
        ref = np.fft.fftn(x.astype('D'))
        a = B.fftw.aligned(shape, dtype=dtype)
        b = B.fftw.aligned(shape, dtype=dtype)
        a[...] = x
        plan = B.fftw.fftn(a, axes=(0, 1, 2), output_array=b)
        try:
            for var in range(8):
                _lib.set_option('variant_rot', var)
                b[...] = 0
                y = np.asarray(plan(a, b))
                assert relerr(y, ref) < TOL[dtype], (n, dtype, var)
        finally:
            _lib.set_option('variant_rot', -1)
        plan.destroy()


def test_pfft_merged_equals_staged(B, monkeypatch):
    """single rank: the merged 3-axis plan gives what the reference's stage-by-stage chain gives"""
    shape = (64, 128, 64)
    x = rand(shape, 'D', seed=3)
    fft = B.PFFT(B.COMM_WORLD, shape, dtype='D')
    assert fft.forward._merged is not None and len(fft.xfftn) == 3
    u = B.newDistArray(fft, False)
    u[...] = x
    uh = np.asarray(fft.forward(u)).copy()
    assert relerr(uh, np.fft.fftn(x) / x.size) < 1e-12
    ub = np.asarray(fft.backward(fft.forward(u)))
    assert np.abs(ub - x).max() < 1e-12
    fft.destroy()
    monkeypatch.setenv('B2F_MERGE', '0')
    fft2 = B.PFFT(B.COMM_WORLD, shape, dtype='D')
    assert fft2.forward._merged is None
    u2 = B.newDistArray(fft2, False)
    u2[...] = x
    uh2 = np.asarray(fft2.forward(u2))
    assert np.abs(uh2 - uh).max() < 1e-14
    fft2.destroy()


def test_full_array_512_cubed_vs_numpy(B):
    """BASELINE.json config 2 (3D c2c 512^3 complex128 on one B200): every point of the forward
    transform against numpy's pocketfft on the host, and the round trip."""
    import scipy.fft as sfft
    shape = (512, 512, 512)
    rng = np.random.default_rng(512)
    x = rng.random(shape) + 1j * rng.random(shape)
    fft = B.PFFT(B.COMM_WORLD, shape, dtype='D')
    u = B.newDistArray(fft, False)
    u[...] = x
    uh = fft.forward(u)
    ref = sfft.fftn(x, workers=-1)
    ref *= 1.0 / x.size
    got = np.asarray(uh)
    err = np.abs(got - ref).max()
    assert err < 1e-12, err
    del ref
    ub = np.asarray(fft.backward(uh))
    assert np.abs(ub - x).max() < 1e-12
    fft.destroy()
