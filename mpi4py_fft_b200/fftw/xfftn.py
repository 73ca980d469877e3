"""Planner front end for serial transforms on the device.

Same public surface as /root/reference/mpi4py_fft/fftw/xfftn.py -- ``fftn, ifftn,
rfftn, irfftn, dctn, idctn, dstn, idstn`` return a *planned transform object*
(:class:`FFT`) bound to an input and an output array; calling it executes --
but the object drives ``b2f_planxfftn`` / ``b2f_execute`` of ``libb200fft.so``
instead of FFTW.  The shape/dtype/normalisation rules are the reference's
(xfftn.py:228-239 r2c halves ``axes[-1]``; :306-326 c2r; :763-816 normalisation);
the r2r type->kind tables repeat FFTW's numbering (:14-36).

Arrays may be device arrays or unallocated :class:`ArraySpec` s; in the latter
case HBM is allocated on first access (planning itself is host arithmetic).
``hfftn``/``ihfftn`` and long double are outside the B200 scope and raise.
"""
from __future__ import annotations

import numpy as np

from ..devarray import copy_out as _copy_out, ArraySpec, DeviceArray, as_tensor, device_ptr, empty, np_dtype_of
from .utilities import (FFTW_FORWARD, FFTW_BACKWARD, FFTW_REDFT00, FFTW_REDFT01, FFTW_REDFT10,
                        FFTW_REDFT11, FFTW_RODFT00, FFTW_RODFT01, FFTW_RODFT10, FFTW_RODFT11,
                        FFTW_MEASURE, FFTW_DESTROY_INPUT, FFTW_UNALIGNED, FFTW_CONSERVE_MEMORY,
                        FFTW_EXHAUSTIVE, FFTW_PRESERVE_INPUT, FFTW_PATIENT, FFTW_ESTIMATE,
                        FFTW_WISDOM_ONLY, C2C_FORWARD, C2C_BACKWARD, R2C, C2R, FFTW_R2HC, FFTW_HC2R,
                        FFTW_DHT, get_alignment, aligned, aligned_like)

flag_dict = {key: val for key, val in list(globals().items()) if key.startswith('FFTW_')}

# scipy/numpy "type" -> FFTW kind, forward and inverse (DCT-II <-> DCT-III etc.)
dct_type = {1: FFTW_REDFT00, 2: FFTW_REDFT10, 3: FFTW_REDFT01, 4: FFTW_REDFT11}
idct_type = {1: FFTW_REDFT00, 2: FFTW_REDFT01, 3: FFTW_REDFT10, 4: FFTW_REDFT11}
dst_type = {1: FFTW_RODFT00, 2: FFTW_RODFT10, 3: FFTW_RODFT01, 4: FFTW_RODFT11}
idst_type = {1: FFTW_RODFT00, 2: FFTW_RODFT01, 3: FFTW_RODFT10, 4: FFTW_RODFT11}

_R2R_LOGICAL = {FFTW_REDFT00: lambda n: 2 * (n - 1), FFTW_RODFT00: lambda n: 2 * (n + 1)}


def get_normalization(kind, shape, axes):
    """1 / (product of the logical DFT sizes of the transformed axes): N for
    Fourier kinds, 2N for the DCT/DST kinds except 2(N-1) for REDFT00 and
    2(N+1) for RODFT00 (reference xfftn.py:763-816)."""
    kinds = [kind] * len(axes) if isinstance(kind, (int, np.integer)) else list(kind)
    assert len(kinds) == len(axes)
    total = 1
    for knd, axis in zip(kinds, axes):
        n = int(shape[axis])
        if knd in _R2R_LOGICAL:
            total *= _R2R_LOGICAL[knd](n)
        elif FFTW_REDFT00 <= knd <= FFTW_RODFT11:
            total *= 2 * n
        else:
            total *= n
    return 1. / total


class FFT(object):
    """Planned batched transform over ``axes`` of a C-contiguous block.

    Counterpart of the reference's Cython ``FFT``
    (/root/reference/mpi4py_fft/fftw/fftw_xfftn.pyx:50-296): holds the plan and a
    pair of arrays, ``__call__`` runs the plan on the held arrays or -- the
    "implicit" new-array execute -- directly on compatible arrays handed in.
    ``normalize=True`` multiplies by the stored factor, fused into the last
    kernel pass.
    """

    def __init__(self, input_array, output_array, axes=(-1,), kind=FFTW_FORWARD, threads=1,
                 flags=FFTW_MEASURE, normalization=1.0):
        self._in = input_array
        self._out = output_array
        self.input_shape = tuple(input_array.shape)
        self.output_shape = tuple(output_array.shape)
        self.input_dtype = np.dtype(input_array.dtype)
        self.output_dtype = np.dtype(output_array.dtype)
        nd = len(self.input_shape)
        self.axes = tuple(int(a) % nd for a in axes)
        self.kinds = [int(kind)] if isinstance(kind, (int, np.integer)) else [int(k) for k in kind]
        self.kind = self.kinds[0]
        self._M = float(normalization)
        if self.input_dtype.char.lower() == 'g':
            raise RuntimeError("Failure creating B200 plan: long double has no device type")
        self.precision = 4 if self.input_dtype.char.lower() == 'f' else 8
        self._plan = None

    # -- arrays (allocated on first touch) ------------------------------------
    @property
    def input_array(self):
        if isinstance(self._in, ArraySpec):
            self._in = self._in.allocate()
        return self._in

    @property
    def output_array(self):
        if isinstance(self._out, ArraySpec):
            self._out = self._out.allocate()
        return self._out

    def update_arrays(self, input_array, output_array):
        assert self.input_shape == tuple(input_array.shape)
        assert self.input_dtype == np_dtype_of(input_array)
        assert self.output_shape == tuple(output_array.shape)
        assert self.output_dtype == np_dtype_of(output_array)
        self._in = input_array
        self._out = output_array

    def get_normalization(self):
        """The factor applied when called with ``normalize=True``."""
        return self._M

    # -- plan --------------------------------------------------------------------
    def plan(self):
        if self._plan is None:
            from .._lib import Plan
            self._plan = Plan(self.input_shape, self.output_shape, self.axes, self.kinds,
                              self.precision)
        return self._plan

    def print_plan(self):
        print(self.plan().describe())

    def destroy(self):
        if self._plan is not None:
            self._plan.destroy()
            self._plan = None

    def execute(self, src, dst, scale=1.0):
        """Enqueue the transform ``src -> dst`` (device arrays of the planned
        shapes/dtypes) on the current CUDA stream, times ``scale``."""
        assert tuple(src.shape) == self.input_shape and np_dtype_of(src) == self.input_dtype
        assert tuple(dst.shape) == self.output_shape and np_dtype_of(dst) == self.output_dtype
        self.plan().execute(device_ptr(src), device_ptr(dst), scale)
        return dst

    def execute_scatter(self, src, work, scale, transfer_handle, direction, peer_ptrs, sync=True):
        """The transform with its last pass storing into the owners' windows of the
        following redistribution (b2f_execute_scatter); ``work`` (may be None for a
        one-axis stage) holds the intermediate of a multi-axis stage."""
        assert tuple(src.shape) == self.input_shape and np_dtype_of(src) == self.input_dtype
        self.plan().execute_scatter(device_ptr(src), device_ptr(work) if work is not None else 0, scale,
                                    transfer_handle, direction, peer_ptrs, sync)

    def execute_chunk(self, src, dst, scale, mode, begin, count, view_outer=0, view_ostride=0, grid_cap=0):
        """part of a one-axis stage (b2f_execute_chunk)"""
        self.plan().execute_chunk(device_ptr(src), device_ptr(dst), scale, mode, begin, count, view_outer,
                                  view_ostride, grid_cap)

    def execute_scatter_chunk(self, src, scale, transfer_handle, direction, peer_ptrs, sync_flags, mode, begin, count,
                              view_outer=0, view_ostride=0, grid_cap=0):
        self.plan().execute_scatter_chunk(device_ptr(src), scale, transfer_handle, direction, peer_ptrs, sync_flags,
                                          mode, begin, count, view_outer, view_ostride, grid_cap)

    def __call__(self, input_array=None, output_array=None, implicit=True, normalize=False, **kw):
        """Compute the transform; returns the output array.

        Arrays given here are used in place of the held ones when they are
        device arrays of the planned shape and dtype (and ``implicit`` is
        true); anything else (host arrays, other dtypes) is staged through the
        held arrays.
        """
        src = self._usable(input_array, self.input_shape, self.input_dtype) if implicit else None
        if src is None:
            src = self.input_array
            if input_array is not None:
                src[...] = input_array
        dst = self._usable(output_array, self.output_shape, self.output_dtype) if implicit else None
        direct = dst is not None
        if dst is None:
            dst = self.output_array
        self.execute(src, dst, self._M if normalize else 1.0)
        if output_array is not None and not direct:
            _copy_out(dst, output_array)
            return output_array
        return dst

    @staticmethod
    def _usable(a, shape, dtype):
        if a is None or isinstance(a, np.ndarray):
            return None
        try:
            t = as_tensor(a)
        except TypeError:
            return None
        if tuple(t.shape) != tuple(shape) or np_dtype_of(a) != dtype or not t.is_cuda or not t.is_contiguous():
            return None
        return a


def _like(arr, shape, dtype):
    """Array (or spec) of ``shape``/``dtype`` in the same state as ``arr``:
    planning from specs stays allocation free."""
    if isinstance(arr, ArraySpec):
        return ArraySpec(shape, dtype)
    return empty(shape, dtype)


def _norm_axes(axes, ndim):
    axes = (axes,) if isinstance(axes, (int, np.integer)) else tuple(axes)
    return tuple(int(a) % ndim for a in axes)


def fftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    """Planned complex-to-complex forward transform (sign -1, unnormalised)."""
    assert input_array.dtype.char in 'FD'
    axes = _norm_axes(axes, len(input_array.shape))
    if output_array is None:
        output_array = _like(input_array, input_array.shape, input_array.dtype)
    else:
        assert tuple(input_array.shape) == tuple(output_array.shape)
        assert output_array.dtype.char == input_array.dtype.char.upper()
    size = int(np.prod(np.take(input_array.shape, axes)))
    return FFT(input_array, output_array, axes, FFTW_FORWARD, threads, flags, 1.0 / size)


def ifftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    """Planned complex-to-complex backward transform (sign +1, unnormalised)."""
    assert input_array.dtype.char in 'FD'
    axes = _norm_axes(axes, len(input_array.shape))
    if output_array is None:
        output_array = _like(input_array, input_array.shape, input_array.dtype)
    else:
        assert tuple(input_array.shape) == tuple(output_array.shape)
    size = int(np.prod(np.take(input_array.shape, axes)))
    return FFT(input_array, output_array, axes, FFTW_BACKWARD, threads, flags, 1.0 / size)


def rfftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    """Planned real-to-complex transform: the last listed axis keeps the
    non-redundant half, n//2+1 points."""
    assert input_array.dtype.char in 'fd'
    axes = _norm_axes(axes, len(input_array.shape))
    half = input_array.shape[axes[-1]] // 2 + 1
    if output_array is None:
        shp = list(input_array.shape)
        shp[axes[-1]] = half
        output_array = _like(input_array, shp, np.dtype(input_array.dtype.char.upper()))
    else:
        assert output_array.shape[axes[-1]] == half
    size = int(np.prod(np.take(input_array.shape, axes)))
    return FFT(input_array, output_array, axes, R2C, threads, flags, 1.0 / size)


def irfftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    """Planned complex-to-real transform.  The real length of the last axis is
    ``s[-1]`` when ``s`` is given, else the even length 2(n-1)."""
    assert input_array.dtype.char in 'FD'
    flags = (flags,) if isinstance(flags, (int, np.integer)) else tuple(flags)
    assert FFTW_PRESERVE_INPUT not in flags
    axes = _norm_axes(axes, len(input_array.shape))
    shp = list(input_array.shape)
    if s is not None:
        assert len(axes) == len(s)
        for n, axis in zip(s, axes):
            shp[axis] = int(n)
    else:
        shp[axes[-1]] = 2 * shp[axes[-1]] - 2
    if output_array is None:
        output_array = _like(input_array, shp, np.dtype(input_array.dtype.char.lower()))
    else:
        assert list(output_array.shape) == shp
    assert shp[axes[-1]] // 2 + 1 == input_array.shape[axes[-1]]
    size = int(np.prod(np.take(output_array.shape, axes)))
    return FFT(input_array, output_array, axes, C2R, threads, flags, 1.0 / size)


def _r2r(table, input_array, axes, type, threads, flags, output_array):
    assert input_array.dtype.char in 'fd'
    axes = _norm_axes(axes, len(input_array.shape))
    if output_array is None:
        output_array = _like(input_array, input_array.shape, input_array.dtype)
    else:
        assert tuple(input_array.shape) == tuple(output_array.shape)
    kinds = [table[type]] * len(axes)
    return FFT(input_array, output_array, axes, kinds, threads, flags,
               get_normalization(kinds, input_array.shape, axes))


def dctn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,),
         output_array=None):
    """Planned discrete cosine transform of ``type`` 1-4 (unnormalised, FFTW
    REDFT conventions == ``scipy.fft.dctn(norm=None)``)."""
    return _r2r(dct_type, input_array, axes, type, threads, flags, output_array)


def idctn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,),
          output_array=None):
    """Planned inverse (up to normalisation) of :func:`dctn` of the same type."""
    return _r2r(idct_type, input_array, axes, type, threads, flags, output_array)


def dstn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,),
         output_array=None):
    """Planned discrete sine transform of ``type`` 1-4 (FFTW RODFT conventions)."""
    return _r2r(dst_type, input_array, axes, type, threads, flags, output_array)


def idstn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,),
          output_array=None):
    """Planned inverse (up to normalisation) of :func:`dstn` of the same type."""
    return _r2r(idst_type, input_array, axes, type, threads, flags, output_array)


def hfftn(*args, **kw):
    raise NotImplementedError("hfftn is outside the B200 hot-path scope (SURVEY.md section 2)")


def ihfftn(*args, **kw):
    raise NotImplementedError("ihfftn is outside the B200 hot-path scope (SURVEY.md section 2)")


inverse = {
    FFTW_RODFT11: FFTW_RODFT11, FFTW_REDFT11: FFTW_REDFT11,
    FFTW_RODFT01: FFTW_RODFT10, FFTW_RODFT10: FFTW_RODFT01,
    FFTW_REDFT01: FFTW_REDFT10, FFTW_REDFT10: FFTW_REDFT01,
    FFTW_RODFT00: FFTW_RODFT00, FFTW_REDFT00: FFTW_REDFT00,
    rfftn: irfftn, irfftn: rfftn, fftn: ifftn, ifftn: fftn,
    dctn: idctn, idctn: dctn, dstn: idstn, idstn: dstn,
}
