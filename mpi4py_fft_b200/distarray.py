"""Distributed array: a device-resident local block plus its :class:`Pencil`.

Interface of /root/reference/mpi4py_fft/distarray.py:10-363,442-485.  The
reference subclasses ``numpy.ndarray``; a block in HBM cannot, so ``DistArray``
subclasses :class:`DeviceArray` (numpy-like handle on a CUDA tensor) and keeps
the same constructor, properties (``alignment, global_shape, substart, subcomm,
commsizes, pencil, rank, dimensions, v``), ``local_slice``, ``redistribute``
and ``__getitem__`` behaviour for tensor components.  Host staging is explicit
(``np.asarray(a)`` / ``a[...] = host_array``).  Parallel file IO
(``write/read/get``, distarray.py:182-241,365-439) is outside the hot path.
"""
from __future__ import annotations

from numbers import Number, Integral

import numpy as np

from .comm import Comm, COMM_SELF, COMM_WORLD
from .devarray import DeviceArray, as_tensor, empty, np_dtype_of, torch_dtype
from .pencil import Pencil, Subcomm

comm = COMM_WORLD


class DistArray(DeviceArray):
    """Local block of a global array distributed over a process grid.

    Parameters as the reference's: ``global_shape``, ``subcomm`` (None,
    :class:`Subcomm`, or a sequence of ints understood by :class:`Subcomm`, or a
    sequence of communicators), ``val``, ``dtype``, ``buffer`` (device array or
    CUDA tensor of the local shape to wrap), ``alignment``, ``rank`` (number of
    leading tensor-component axes, never distributed).
    """

    def __init__(self, global_shape, subcomm=None, val=None, dtype=float, buffer=None,
                 strides=None, alignment=None, rank=0):
        global_shape = tuple(int(n) for n in global_shape)
        dtype = np.dtype(dtype)
        if len(global_shape[rank:]) < 2:          # nothing to distribute: plain local array
            local_shape, p0 = global_shape, None
        else:
            if isinstance(subcomm, Subcomm):
                pass
            elif isinstance(subcomm, (tuple, list)):
                assert len(subcomm) == len(global_shape[rank:])
                if not all(isinstance(s, Comm) for s in subcomm):
                    subcomm = Subcomm(comm, subcomm)
            else:
                assert subcomm is None
                dims = [0] * len(global_shape[rank:])
                if alignment is not None:
                    dims[alignment] = 1
                else:
                    dims[-1] = 1
                    alignment = len(dims) - 1
                subcomm = Subcomm(comm, dims)
            sizes = [s.Get_size() for s in subcomm]
            if alignment is not None:
                assert isinstance(alignment, (int, np.integer))
                assert sizes[alignment] == 1
            else:
                alignment = int(np.flatnonzero(np.array(sizes) == 1)[-1])
            p0 = Pencil(subcomm, global_shape[rank:], axis=alignment)
            local_shape = global_shape[:rank] + p0.subshape
        if buffer is not None:
            t = as_tensor(buffer)
            assert tuple(t.shape) == tuple(local_shape), (tuple(t.shape), local_shape)
            assert np_dtype_of(buffer) == dtype
        else:
            t = empty(local_shape, dtype, fill=val if isinstance(val, Number) else None)._t
        DeviceArray.__init__(self, t)
        self._p0 = p0
        self._rank = rank

    # -- views keep the distribution ------------------------------------------------
    def _wrap(self, t):
        """elementwise results of the same local shape stay DistArrays"""
        if tuple(t.shape) == self.shape:
            return self._view(t, self._p0, self._rank)
        return DeviceArray(t)

    @classmethod
    def _view(cls, t, p0, rank):
        obj = cls.__new__(cls)
        DeviceArray.__init__(obj, t)
        obj._p0 = p0
        obj._rank = rank
        return obj

    def __getitem__(self, i):
        # a tensor component is still a DistArray; anything else is a plain view
        if self.ndim == 1 or self._rank == 0:
            return DeviceArray.__getitem__(self, i)
        if isinstance(i, (Integral, slice)):
            t = self._t[i]
            return self._view(t, self._p0, self._rank - (self._t.dim() - t.dim()))
        if isinstance(i, tuple) and len(i) <= self._rank and \
                all(isinstance(j, (Integral, slice)) for j in i):
            t = self._t[i]
            return self._view(t, self._p0, self._rank - (self._t.dim() - t.dim()))
        return DeviceArray.__getitem__(self, i)

    # -- distribution metadata ---------------------------------------------------------
    @property
    def alignment(self):
        """Axis (not counting tensor axes) along which the block is undivided."""
        return self._p0.axis

    @property
    def global_shape(self):
        return self.shape[:self.rank] + self._p0.shape

    @property
    def substart(self):
        return (0,) * self.rank + self._p0.substart

    @property
    def subcomm(self):
        return (COMM_SELF,) * self.rank + self._p0.subcomm

    @property
    def commsizes(self):
        return [s.Get_size() for s in self.subcomm]

    @property
    def pencil(self):
        return self._p0

    @property
    def rank(self):
        return self._rank

    @property
    def dimensions(self):
        return len(self._p0.shape)

    @property
    def v(self):
        """The local block as a plain :class:`DeviceArray` (shares memory)."""
        return DeviceArray(self._t)

    def local_slice(self):
        """Slices of the global array covered by the local block."""
        v = [slice(s, s + n) for s, n in zip(self._p0.substart, self._p0.subshape)]
        return tuple([slice(0, n) for n in self.shape[:self.rank]] + v)

    # -- global redistribution ----------------------------------------------------------
    def get_pencil_and_transfer(self, axis):
        p1 = self._p0.pencil(axis)
        return p1, self._p0.transfer(p1, self.dtype)

    def redistribute(self, axis=None, out=None):
        """Realign the array along ``axis`` (or into ``out``) with one global
        transpose on the device; semantics of reference distarray.py:298-363."""
        if axis == self.alignment:
            return self
        if axis is not None and isinstance(out, DistArray):
            assert axis == out.alignment
        if axis is not None and self.commsizes[self.rank + axis] == 1:
            # already undivided along axis: only the bookkeeping changes
            self._p0.axis = axis
            return self
        if out is not None:
            assert isinstance(out, DistArray)
            assert self.global_shape == out.global_shape
            axis = out.alignment
            if self.commsizes == out.commsizes:
                out[...] = self
                return out
            for i in range(len(self._p0.shape)):
                if i not in (self.alignment, out.alignment):
                    assert self.pencil.subcomm[i] == out.pencil.subcomm[i]
                    assert self.pencil.subshape[i] == out.pencil.subshape[i]

        p1, transfer = self.get_pencil_and_transfer(axis)
        if out is None:
            out = DistArray(self.global_shape, subcomm=p1.subcomm, dtype=self.dtype,
                            alignment=axis, rank=self.rank)
        if self.rank == 0:
            transfer.forward(self, out)
        elif self.rank == 1:
            for i in range(self.shape[0]):
                transfer.forward(self[i], out[i])
        elif self.rank == 2:
            for i in range(self.shape[0]):
                for j in range(self.shape[1]):
                    transfer.forward(self[i, j], out[i, j])
        transfer.destroy()
        return out

    # -- IO: outside the hot path --------------------------------------------------------
    def get(self, gslice):
        raise NotImplementedError("DistArray.get needs the parallel-IO side of the reference, "
                                  "which is outside the B200 hot-path scope")

    def write(self, *args, **kw):
        raise NotImplementedError("parallel HDF5/NetCDF output is outside the B200 hot-path scope; "
                                  "stage with np.asarray(a) and write on the host")

    def read(self, *args, **kw):
        raise NotImplementedError("parallel HDF5/NetCDF input is outside the B200 hot-path scope")


def newDistArray(pfft, forward_output=True, val=0, rank=0, view=False):
    """New :class:`DistArray` shaped and typed for the input (``forward_output``
    False) or output (True) of ``pfft.forward`` (reference distarray.py:442-485).
    ``rank`` prefixes the shape with that many tensor axes of length ndim."""
    global_shape = pfft.global_shape(forward_output)
    p0 = pfft.pencil[forward_output]
    dtype = pfft.dtype(forward_output is True)
    global_shape = (len(global_shape),) * rank + tuple(global_shape)
    z = DistArray(global_shape, subcomm=p0.subcomm, val=val, dtype=dtype,
                  alignment=p0.axis, rank=rank)
    return z.v if view else z


def Function(*args, **kwargs):  # pragma: no cover
    import warnings
    warnings.warn("Function() is deprecated; use newDistArray().", FutureWarning)
    if 'tensor' in kwargs:
        kwargs['rank'] = 1
        del kwargs['tensor']
    return newDistArray(*args, **kwargs)
