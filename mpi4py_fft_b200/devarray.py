"""Device-resident arrays with a numpy-flavoured face.

The reference keeps every block in host ``numpy`` arrays allocated with 32-byte
alignment for FFTW's SIMD codelets (``fftw.aligned``,
/root/reference/mpi4py_fft/fftw/utilities.pyx:54-104).  Here blocks live in HBM:
a :class:`DeviceArray` wraps a CUDA ``torch.Tensor`` (torch is used for memory,
streams and elementwise convenience only -- transforms and transposes go
through ``libb200fft.so``) and exposes ``shape / dtype / ndim / size``,
indexing, in-place assignment from numpy, arithmetic, and ``__array__`` for
host read-back, which is what scripts written against the reference touch.
"""
from __future__ import annotations

from numbers import Number

import numpy as np

_TORCH_OF = None
_NP_OF = None


def _maps():
    global _TORCH_OF, _NP_OF
    if _TORCH_OF is None:
        import torch
        _TORCH_OF = {np.dtype('f4'): torch.float32, np.dtype('f8'): torch.float64,
                     np.dtype('c8'): torch.complex64, np.dtype('c16'): torch.complex128,
                     np.dtype('i4'): torch.int32, np.dtype('i8'): torch.int64,
                     np.dtype('u1'): torch.uint8, np.dtype('bool'): torch.bool}
        _NP_OF = {v: k for k, v in _TORCH_OF.items()}
    return _TORCH_OF, _NP_OF


def torch_dtype(dtype):
    return _maps()[0][np.dtype(dtype)]


def np_dtype_of(a):
    """numpy dtype of a DeviceArray, torch tensor or numpy array."""
    if isinstance(a, DeviceArray):
        return a.dtype
    if isinstance(a, np.ndarray):
        return a.dtype
    return _maps()[1][a.dtype]


def device():
    """The CUDA device of this rank; raises when there is none -- the B200 path
    has no CPU fallback."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("mpi4py_fft_b200 needs a CUDA device (B200, sm_100a); "
                           "no CPU fallback exists for transforms or transposes")
    return torch.device('cuda', torch.cuda.current_device())


def as_tensor(a):
    """CUDA torch tensor behind ``a`` (DeviceArray or tensor)."""
    if isinstance(a, DeviceArray):
        return a._t
    import torch
    if isinstance(a, torch.Tensor):
        return a
    raise TypeError("expected a device array, got %r" % type(a))


def device_ptr(a):
    t = as_tensor(a)
    if not t.is_cuda:
        raise RuntimeError("array is not on a CUDA device")
    if not t.is_contiguous():
        raise RuntimeError("device arrays handed to libb200fft must be C-contiguous")
    return t.data_ptr()


class ArraySpec(object):
    """Shape and dtype of an array that has not been allocated (planning is
    pure host arithmetic; HBM is touched on first use)."""
    __slots__ = ('shape', 'dtype')

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)

    @property
    def ndim(self):
        return len(self.shape)

    def allocate(self, fill=0):
        return empty(self.shape, self.dtype, fill=fill)


def pinned_empty(shape, dtype=np.float64):
    """Page-locked host array (numpy view of pinned torch memory) for fast
    staging in and out of device arrays."""
    import torch
    shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
    return torch.empty(shape, dtype=torch_dtype(dtype), pin_memory=True).numpy()


def empty(shape, dtype=np.float64, fill=None):
    import torch
    shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
    t = torch.empty(shape, dtype=torch_dtype(dtype), device=device())
    if fill is not None:
        t.fill_(fill)
    return DeviceArray(t)


def _unwrap(x):
    """operand for a torch expression"""
    if isinstance(x, DeviceArray):
        return x._t
    if isinstance(x, np.ndarray):
        import torch
        return torch.from_numpy(np.ascontiguousarray(x))
    return x


class DeviceArray(object):
    """numpy-like handle on a CUDA tensor (see module docstring)."""

    __array_priority__ = 1000

    def __init__(self, tensor):
        self._t = tensor

    # -- numpy-style metadata ---------------------------------------------------
    @property
    def tensor(self):
        """the underlying ``torch.Tensor`` (shares memory)"""
        return self._t

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return _maps()[1][self._t.dtype]

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def size(self):
        return self._t.numel()

    @property
    def itemsize(self):
        return self._t.element_size()

    @property
    def nbytes(self):
        return self._t.numel() * self._t.element_size()

    @property
    def real(self):
        return self._wrap(self._t.real if self._t.is_complex() else self._t)

    @property
    def imag(self):
        import torch
        return self._wrap(self._t.imag if self._t.is_complex() else torch.zeros_like(self._t))

    def __len__(self):
        return self._t.shape[0]

    def _wrap(self, t):
        return DeviceArray(t)

    # -- host staging -------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        a = self._t.detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def asnumpy(self):
        """device -> host copy as a plain ``numpy.ndarray``"""
        return self.__array__()

    def copy_to_host(self, host):
        """device -> host copy straight into ``host`` (numpy array of the same
        shape and dtype; pinned memory gives full PCIe rate)"""
        import torch
        assert isinstance(host, np.ndarray) and host.dtype == self.dtype and tuple(host.shape) == self.shape
        assert host.flags.c_contiguous
        torch.from_numpy(host).copy_(self._t)
        return host

    def set(self, host):
        """host -> device copy (``host`` broadcastable to ``self.shape``)"""
        self[...] = host
        return self

    # -- indexing -------------------------------------------------------------------
    def _index_result(self, t):
        return self._wrap(t)

    def __getitem__(self, idx):
        idx = _unwrap_index(idx)
        return self._index_result(self._t[idx])

    def __setitem__(self, idx, value):
        import torch
        idx = _unwrap_index(idx)
        if isinstance(value, DeviceArray):
            value = value._t
        elif isinstance(value, np.ndarray):
            dst = self._t[idx]
            if value.dtype == self.dtype and tuple(value.shape) == tuple(dst.shape) and value.flags.c_contiguous:
                dst.copy_(torch.from_numpy(value))       # one host->device copy, no staging tensor
                return
            value = torch.from_numpy(np.ascontiguousarray(value)).to(self._t.device, non_blocking=False)
        elif not isinstance(value, (Number, torch.Tensor)):
            value = torch.as_tensor(np.asarray(value)).to(self._t.device)
        if isinstance(value, torch.Tensor) and value.dtype != self._t.dtype:
            if value.is_complex() and not self._t.is_complex():
                value = value.real
            value = value.to(self._t.dtype)
        self._t[idx] = value

    def fill(self, value):
        self._t.fill_(value)

    def copy(self):
        return self._wrap(self._t.clone())

    def astype(self, dtype):
        return self._wrap(self._t.to(torch_dtype(dtype)))

    def reshape(self, *shape):
        return DeviceArray(self._t.reshape(*shape))

    # -- arithmetic (elementwise, on device) ---------------------------------------
    def _bin(self, other, op, reverse=False):
        import torch
        o = _unwrap(other)
        if isinstance(o, torch.Tensor) and not o.is_cuda:
            o = o.to(self._t.device)
        r = op(o, self._t) if reverse else op(self._t, o)
        return self._wrap(r)

    def _ibin(self, other, op):
        import torch
        o = _unwrap(other)
        if isinstance(o, torch.Tensor) and not o.is_cuda:
            o = o.to(self._t.device)
        op(self._t, o)
        return self

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return self._wrap(-self._t)
    def __abs__(self): return self._wrap(self._t.abs())
    def __iadd__(self, o): return self._ibin(o, lambda a, b: a.add_(b))
    def __isub__(self, o): return self._ibin(o, lambda a, b: a.sub_(b))
    def __imul__(self, o): return self._ibin(o, lambda a, b: a.mul_(b))
    def __itruediv__(self, o): return self._ibin(o, lambda a, b: a.div_(b))

    def conj(self):
        return self._wrap(self._t.conj().resolve_conj())

    def sum(self):
        return self._t.sum().item()

    def __repr__(self):
        return "%s(shape=%r, dtype=%s, device=%s)" % (type(self).__name__, self.shape, self.dtype,
                                                      self._t.device)


def _unwrap_index(idx):
    if isinstance(idx, tuple):
        return tuple(_unwrap_index(i) for i in idx)
    if isinstance(idx, DeviceArray):
        return idx._t
    if isinstance(idx, np.integer):
        return int(idx)
    return idx


def copy_out(dev, target):
    """``target[...] = dev`` for a host (numpy) or device target"""
    if isinstance(target, np.ndarray):
        if target.dtype == dev.dtype and tuple(target.shape) == dev.shape and target.flags.c_contiguous:
            dev.copy_to_host(target)
        else:
            target[...] = np.asarray(dev)
    else:
        target[...] = dev
