"""Serial (one rank) transform stage: the backend plug point of the reference.

Interface of /root/reference/mpi4py_fft/libfft.py (``FFT(shape, axes, dtype,
padding, backend, transforms, **kw)`` with ``.forward`` / ``.backward`` callables
exposing ``.input_array`` / ``.output_array``), backed by a single device
planner instead of the reference's table of CPU backends (libfft.py:379-385).

Differences that follow from living in HBM:
  * arrays are allocated on first touch, planning is host arithmetic;
  * ``forward(u)`` runs straight from ``u`` (and into ``output_array`` when one
    is given) whenever those are device arrays of the planned shape -- the two
    staging copies of libfft.py:213-216 disappear; host (numpy) arrays are
    staged through the plan-owned device arrays;
  * normalisation is a scale fused into the last kernel pass, not a second
    sweep (libfft.py:412-413).
"""
from __future__ import annotations

import numpy as np

from . import fftw
from .devarray import copy_out as _copy_out, ArraySpec, DeviceArray, np_dtype_of


def _Xfftn_plan_b200(shape, axes, dtype, transforms, options):
    """(forward, backward) planned transform objects for one stage; same role
    as ``_Xfftn_plan_fftw`` (reference libfft.py:48-79)."""
    options = dict(options)
    threads = options.pop('threads', 1)
    effort = options.pop('planner_effort', 'FFTW_MEASURE')
    options.pop('overwrite_input', None)
    flags = (fftw.flag_dict.get(effort, 0),)

    transforms = {} if transforms is None else transforms
    if tuple(axes) in transforms:
        plan_fwd, plan_bck = transforms[tuple(axes)]
    elif np.issubdtype(dtype, np.floating):
        plan_fwd, plan_bck = fftw.rfftn, fftw.irfftn
    else:
        plan_fwd, plan_bck = fftw.fftn, fftw.ifftn

    s = tuple(int(n) for n in np.take(shape, axes))
    U = ArraySpec(shape, dtype)
    fwd = plan_fwd(U, s=s, axes=axes, threads=threads, flags=flags)
    V = fwd._out
    bck = plan_bck(V, s=s, axes=axes, threads=threads, flags=flags, output_array=U)
    return fwd, bck


class _Stage(object):
    """One direction of a serial stage as a callable with arrays attached."""

    def __init__(self, owner, planned, default_normalize):
        self._owner = owner
        self._planned = planned
        self._default_normalize = default_normalize

    @property
    def input_array(self):
        return self._owner._array(self._planned, 'in')

    @property
    def output_array(self):
        return self._owner._array(self._planned, 'out')

    @property
    def input_shape(self):
        return self._planned.input_shape

    @property
    def output_shape(self):
        return self._planned.output_shape

    @property
    def input_dtype(self):
        return self._planned.input_dtype

    @property
    def output_dtype(self):
        return self._planned.output_dtype

    @property
    def destroys_input(self):
        """multi-axis complex-to-real overwrites what it reads (as FFTW's c2r)"""
        return self._planned.kind == fftw.C2R and len(self._planned.axes) > 1

    def run(self, src, dst, normalize=None):
        """Device arrays in, device arrays out, no staging (used by PFFT)."""
        if normalize is None:
            normalize = self._default_normalize
        self._planned.execute(src, dst, self._owner.M if normalize else 1.0)
        return dst

    def can_scatter(self, transfer_handle, direction):
        return self._planned.plan().can_scatter(transfer_handle, direction)

    def run_scatter(self, src, work, normalize, transfer_handle, direction, peer_ptrs):
        """:meth:`run` fused with the following redistribution: the last pass
        writes into the peers' windows (used by PFFT on a multi-GPU node)."""
        if normalize is None:
            normalize = self._default_normalize
        self._planned.execute_scatter(src, work, self._owner.M if normalize else 1.0, transfer_handle,
                                      direction, peer_ptrs, getattr(peer_ptrs, 'sync', True))

    # -- partial launches (pipelined redistribution, mpifft.Transform) ------------
    @property
    def chunkable(self):
        """a one-axis complex Stockham stage can be launched in pieces"""
        pl = self._planned
        if type(self) is not _Stage or len(pl.axes) != 1 or pl.kind not in (fftw.FFTW_FORWARD, fftw.FFTW_BACKWARD):
            return False
        try:
            return pl.plan().describe().startswith('stockham')
        except Exception:       # e.g. tables of a chirp-z stage cannot be built (no device): not chunkable
            return False

    def scale_for(self, normalize):
        if normalize is None:
            normalize = self._default_normalize
        return self._owner.M if normalize else 1.0

    def run_chunk(self, src, dst, normalize, spec, grid_cap=0):
        mode, begin, count, vo, vs = spec
        self._planned.execute_chunk(src, dst, self.scale_for(normalize), mode, begin, count, vo, vs, grid_cap)

    def run_scatter_chunk(self, src, normalize, transfer_handle, direction, peer_ptrs, sync_flags, spec, grid_cap=0):
        mode, begin, count, vo, vs = spec
        self._planned.execute_scatter_chunk(src, self.scale_for(normalize), transfer_handle, direction, peer_ptrs,
                                            sync_flags, mode, begin, count, vo, vs, grid_cap)

    def __call__(self, input_array=None, output_array=None, **kw):
        normalize = kw.pop('normalize', self._default_normalize)
        src = self._planned._usable(input_array, self.input_shape, self.input_dtype)
        if src is not None and self.destroys_input and src is not self.input_array:
            src = None      # stage through the owned array: the caller's input must survive
        if src is None:
            src = self.input_array
            if input_array is not None:
                src[...] = input_array
        dst = self._planned._usable(output_array, self.output_shape, self.output_dtype)
        direct = dst is not None
        if dst is None:
            dst = self.output_array
        self.run(src, dst, normalize)
        if output_array is not None and not direct:
            _copy_out(dst, output_array)
            return output_array
        return dst


class _PaddedStage(_Stage):
    """One direction of a padded (dealiased) serial stage: the transform runs at
    the padded length, the spectrum is truncated (forward) or zero padded
    (backward) by ``b2f_pad_truncate`` with the reference's Nyquist rule
    (libfft.py:263-311, 408-422).  Normalisation is by the padded length and is
    fused into the truncation pass (forward) or the last FFT pass (backward)."""

    def __init__(self, owner, planned, default_normalize, forward):
        _Stage.__init__(self, owner, planned, default_normalize)
        self._forward = forward
        self._fused = None          # None: not tried yet; False: two-pass; else the truncating plan

    def _fused_plan(self):
        """the transform with the dealiasing step folded into its last (forward) or first
        (backward) pass -- ``b2f_plan_set_truncation`` -- or False when the stage's kernels have
        no such flavour (then: transform + ``b2f_pad_truncate``).  B2F_FUSED_PAD=0 disables."""
        if self._fused is None:
            import os
            self._fused = False
            if os.environ.get('B2F_FUSED_PAD', '1') not in ('0', 'false', 'no', ''):
                from ._lib import Plan
                pl = self._planned
                own = self._owner
                try:
                    plan = Plan(pl.input_shape, pl.output_shape, pl.axes, pl.kinds, pl.precision)
                    if plan.set_truncation(own.trunc_shape[own.axes[-1]]):
                        self._fused = plan
                    else:
                        plan.destroy()
                except Exception:
                    self._fused = False
        return self._fused

    @property
    def input_shape(self):
        return self._planned.input_shape if self._forward else self._owner.trunc_shape

    @property
    def output_shape(self):
        return self._owner.trunc_shape if self._forward else self._planned.output_shape

    @property
    def input_dtype(self):
        return self._planned.input_dtype if self._forward else self._owner.trunc_dtype

    @property
    def output_dtype(self):
        return self._owner.trunc_dtype if self._forward else self._planned.output_dtype

    @property
    def destroys_input(self):
        return False

    def run(self, src, dst, normalize=None):
        from ._lib import pad_truncate
        from .devarray import device_ptr
        if normalize is None:
            normalize = self._default_normalize
        own = self._owner
        axis = own.axes[-1]
        spec_shape = own.fwd.output_shape                   # padded spectrum
        outer = int(np.prod(spec_shape[:axis])) if axis else 1
        inner = int(np.prod(spec_shape[axis + 1:])) if axis + 1 < len(spec_shape) else 1
        n_pad, n_keep = spec_shape[axis], own.trunc_shape[axis]
        scale = own.M if normalize else 1.0
        fused = self._fused_plan()
        if fused:
            # one launch: the padded spectrum never exists in memory
            assert tuple(dst.shape) == tuple(self.output_shape) and tuple(src.shape) == tuple(self.input_shape)
            fused.execute(device_ptr(src), device_ptr(dst), scale)
            return dst
        Vp = own._array(own.fwd, 'out')                    # the padded spectrum, plan owned
        if self._forward:
            assert tuple(dst.shape) == own.trunc_shape
            self._planned.execute(src, Vp, 1.0)
            pad_truncate(0, own.real_transform, own.fwd.precision, device_ptr(Vp), device_ptr(dst),
                         outer, n_pad, n_keep, inner, scale)
        else:
            assert tuple(src.shape) == own.trunc_shape
            pad_truncate(1, own.real_transform, own.fwd.precision, device_ptr(src), device_ptr(Vp),
                         outer, n_keep, n_pad, inner, 1.0)
            self._planned.execute(Vp, dst, scale)
        return dst

    def can_scatter(self, transfer_handle, direction):
        # forward: the truncating store comes last, it cannot scatter as well; backward: the
        # transform at the padded length is last and can store into the windows
        if self._forward:
            return False
        fused = self._fused_plan()
        if fused:
            return fused.can_scatter(transfer_handle, direction)
        return _Stage.can_scatter(self, transfer_handle, direction)

    def run_scatter(self, src, work, normalize, transfer_handle, direction, peer_ptrs):
        from ._lib import pad_truncate
        from .devarray import device_ptr
        assert not self._forward
        if normalize is None:
            normalize = self._default_normalize
        own = self._owner
        fused = self._fused_plan()
        if fused:
            fused.execute_scatter(device_ptr(src), device_ptr(work) if work is not None else 0,
                                  own.M if normalize else 1.0, transfer_handle, direction, peer_ptrs,
                                  getattr(peer_ptrs, 'sync', True))
            return
        axis = own.axes[-1]
        spec_shape = own.fwd.output_shape
        outer = int(np.prod(spec_shape[:axis])) if axis else 1
        inner = int(np.prod(spec_shape[axis + 1:])) if axis + 1 < len(spec_shape) else 1
        Vp = own._array(own.fwd, 'out')
        pad_truncate(1, own.real_transform, own.fwd.precision, device_ptr(src), device_ptr(Vp),
                     outer, own.trunc_shape[axis], spec_shape[axis], inner, 1.0)
        self._planned.execute_scatter(Vp, work, own.M if normalize else 1.0, transfer_handle, direction, peer_ptrs,
                                      getattr(peer_ptrs, 'sync', True))

    def __call__(self, input_array=None, output_array=None, **kw):
        normalize = kw.pop('normalize', self._default_normalize)
        src = self._planned._usable(input_array, self.input_shape, self.input_dtype)
        if src is None:
            src = self.input_array
            if input_array is not None:
                src[...] = input_array
        dst = self._planned._usable(output_array, self.output_shape, self.output_dtype)
        direct = dst is not None
        if dst is None:
            dst = self.output_array
        self.run(src, dst, normalize)
        if output_array is not None and not direct:
            _copy_out(dst, output_array)
            return output_array
        return dst

    @property
    def input_array(self):
        return self._owner._array(self._owner.fwd, 'in') if self._forward else self._owner._trunc()

    @property
    def output_array(self):
        return self._owner._trunc() if self._forward else self._owner._array(self._owner.fwd, 'in')


class FFTBase(object):
    """Argument normalisation shared by serial transforms (reference
    libfft.py:221-261)."""

    def __init__(self, shape, axes=None, dtype=float, padding=False):
        shape = [int(n) for n in shape] if np.ndim(shape) else [int(shape)]
        assert len(shape) > 0
        assert min(shape) > 0
        if axes is None:
            axes = list(range(len(shape)))
        else:
            axes = [int(a) for a in axes] if np.ndim(axes) else [int(axes)]
            axes = [a + len(shape) if a < 0 else a for a in axes]
        assert min(axes) >= 0
        assert max(axes) < len(shape)
        assert 0 < len(axes) <= len(shape)
        assert sorted(axes) == sorted(set(axes))
        dtype = np.dtype(dtype)
        assert dtype.char in 'fdgFDG'
        self.shape = shape
        self.axes = axes
        self.dtype = dtype
        self.padding = padding
        self.real_transform = np.issubdtype(dtype, np.floating)
        self.padding_factor = 1


class FFT(FFTBase):
    """Serial transform over ``axes`` of a block of ``shape`` on the device.

    ``forward`` is normalised by default and ``backward`` is not;
    ``normalize=`` at call time overrides either (reference libfft.py:408-422).
    ``backend`` is accepted for source compatibility; every value runs the
    B200 kernels (there are no CPU backends in this package).
    """

    def __init__(self, shape, axes=None, dtype=float, padding=False, backend='b200',
                 transforms=None, **kw):
        FFTBase.__init__(self, shape, axes, dtype, padding)
        if self.dtype.char in 'gG':
            raise RuntimeError("long double transforms are not available on the device")
        pf = 1.0
        if padding is not False:
            pf = padding[self.axes[-1]] if np.ndim(padding) else padding
        self.padding_factor = float(pf)
        self.backend = backend
        self.fwd, self.bck = _Xfftn_plan_b200(self.shape, self.axes, self.dtype, transforms, kw)
        self.M = self.fwd.get_normalization()
        # the two plan-owned arrays: physical side U, spectral side V
        self._U = None
        self._V = None
        self._T = None
        if abs(pf - 1.0) > 1e-8:
            # padded stage: `shape` is the padded physical shape, the spectrum keeps
            # round(n / factor) modes (libfft.py:401-406, 424-434)
            assert len(self.axes) == 1
            assert self.fwd.kind in (fftw.FFTW_FORWARD, fftw.R2C), "padding needs a Fourier stage"
            axis = self.axes[-1]
            tshape = list(self.fwd.output_shape)
            keep = int(np.round(self.shape[axis] / pf))
            tshape[axis] = keep // 2 + 1 if self.real_transform else keep
            self.trunc_shape = tuple(tshape)
            self.trunc_dtype = self.fwd.output_dtype
            self.forward = _PaddedStage(self, self.fwd, True, True)
            self.backward = _PaddedStage(self, self.bck, False, False)
        else:
            self.forward = _Stage(self, self.fwd, True)
            self.backward = _Stage(self, self.bck, False)

    def _trunc(self):
        if self._T is None:
            self._T = ArraySpec(self.trunc_shape, self.trunc_dtype).allocate()
        return self._T

    def _array(self, planned, side):
        physical = (planned is self.fwd) == (side == 'in')
        if physical:
            if self._U is None:
                self._U = ArraySpec(self.fwd.input_shape, self.fwd.input_dtype).allocate()
            return self._U
        if self._V is None:
            self._V = ArraySpec(self.fwd.output_shape, self.fwd.output_dtype).allocate()
        return self._V

    def adopt(self, U=None, V=None):
        """Let the stage use caller-provided device arrays as its owned pair
        (PFFT shares one pair of work arrays between stages)."""
        if U is not None:
            self._U = U
        if V is not None:
            self._V = V

    def destroy(self):
        self.fwd.destroy()
        self.bck.destroy()
        for st in (self.forward, self.backward):
            pl = getattr(st, '_fused', None)
            if pl:
                pl.destroy()
                st._fused = None
