"""examples/taylor_green_dns.py on the CPU: the script's own numerics (spectral
right-hand side, RK4, energy) run against a numpy stand-in for the package --
same call surface (PFFT.forward/backward with arrays that expose ``.tensor``,
newDistArray(rank=1), DeviceArray) -- and must reproduce the reference's known
answer, energy 0.124953117517 at 64^3 after ten steps (reference
examples/spectral_dns_solver.py:129).  The device run of the same script is
tests/test_gpu_dns.py."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT


class _Arr(object):
    def __init__(self, t):
        self.tensor = t

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def __getitem__(self, i):
        return _Arr(self.tensor[i])

    def __setitem__(self, i, v):
        self.tensor[i] = torch.as_tensor(np.asarray(v)) if not isinstance(v, (int, float)) else v


class _Comm(object):
    def Get_rank(self):
        return 0

    def allreduce(self, x):
        return x


class _PFFT(object):
    """single rank, r2c over all axes, forward normalised (reference libfft.py:408-422)"""

    def __init__(self, comm, shape, collapse=False, padding=False):
        assert padding is False
        self.n = tuple(shape)

    def local_slice(self, forward_output=True):
        out = [slice(0, m) for m in self.n]
        if forward_output:
            out[-1] = slice(0, self.n[-1] // 2 + 1)
        return tuple(out)

    def forward(self, a, out):
        out.tensor.copy_(torch.fft.rfftn(a.tensor, norm='forward'))
        return out

    def backward(self, a, out):
        out.tensor.copy_(torch.fft.irfftn(a.tensor, s=self.n, norm='forward'))
        return out

    def destroy(self):
        pass


def _new(fft, forward_output=True, rank=0):
    shape = [m for m in fft.n]
    if forward_output:
        shape[-1] = shape[-1] // 2 + 1
    dt = torch.complex128 if forward_output else torch.float64
    return _Arr(torch.zeros([3] * rank + shape, dtype=dt))


def test_taylor_green_numerics_on_numpy_stand_in(monkeypatch):
    fake = types.ModuleType('mpi4py_fft_b200')
    fake.PFFT, fake.newDistArray, fake.DeviceArray = _PFFT, _new, _Arr
    fake.init = lambda: _Comm()
    monkeypatch.setitem(sys.modules, 'mpi4py_fft_b200', fake)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    sys.modules.pop('taylor_green_dns', None)
    import taylor_green_dns as dns
    try:
        e = dns.solve(6)
    finally:
        sys.modules.pop('taylor_green_dns', None)
    assert round(e - dns.KNOWN_ENERGY_64, 7) == 0, e
