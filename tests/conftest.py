"""Shared test plumbing.

`-m "not gpu"`: oracle vs the reference's golden fixtures, host-side index maps
of the product (virtual ranks), kernel emulation on the CPU, C-ABI symbol check,
world_size-2 gloo tests.  `-m gpu`: the parity tests proper, CUDA path (through
the C ABI) vs the oracle / fixtures.  Nothing here reads /root/reference.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, 'oracle') not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def layouts():
    with open(os.path.join(GOLDEN, 'layouts.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def values():
    return np.load(os.path.join(GOLDEN, 'values.npz'))


def case_kwargs(meta):
    """PFFT keyword arguments of a golden case (JSON lists -> tuples)"""
    kw = dict(meta['kwargs'])
    out = {}
    for k, v in kw.items():
        if k == 'axes':
            out[k] = tuple(tuple(a) if isinstance(a, list) else a for a in v)
        elif isinstance(v, list):
            out[k] = tuple(v)
        else:
            out[k] = v
    return out


@pytest.fixture(scope='session')
def dft_ref():
    """the plain-C restatement of the FFTW definitions (oracle/dft_ref.c)"""
    so = os.path.join(ROOT, 'oracle', 'libdft_ref.so')
    src = os.path.join(ROOT, 'oracle', 'dft_ref.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-o', so, src, '-lm'])
    lib = ctypes.CDLL(so)
    lib.dft_ref.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]

    def run(kind, n, x):
        x = np.ascontiguousarray(x, dtype=complex if kind in (-1, 1, 2) else float)
        if kind in (-1, 1):
            out = np.zeros(n, dtype=complex)
        elif kind == -2:
            out = np.zeros(n // 2 + 1, dtype=complex)
        else:
            out = np.zeros(n, dtype=float)
        assert lib.dft_ref(kind, n, x.ctypes.data, out.ctypes.data) == 0
        return out
    return run


@pytest.fixture(scope='session')
def emu():
    """CPU stepping of the pow2 kernels (tests/emu/emu_fft.cpp)"""
    so = os.path.join(ROOT, 'tests', 'emu', 'libemu_fft.so')
    src = os.path.join(ROOT, 'tests', 'emu', 'emu_fft.cpp')
    deps = [src] + [os.path.join(ROOT, 'mpi4py_fft_b200', 'csrc', f)
                    for f in ('fft_core.cuh', 'fft_pow2.cuh', 'fft_tma.cuh', 'fft_configs.h', 'transfer_put.h', 'chirpz.cuh', 'chirpz_host.h', 'fft_rot.cuh', 'rot_plan.h')]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(['g++', '-std=c++17', '-O1', '-shared', '-fPIC', '-o', so, src])
    lib = ctypes.CDLL(so)
    lib.emu_fft_pow2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                 ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_double, ctypes.c_int]
    lib.emu_fft_tma.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
    lib.emu_fft_scatter.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.c_double, ctypes.c_int]
    lib.emu_fft_real.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
    lib.emu_chirpz.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
    lib.emu_fft_scatter_chunk.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.POINTER(ctypes.c_void_p), ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong]
    lib.emu_put.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                            ctypes.POINTER(ctypes.c_int)]
    return lib
