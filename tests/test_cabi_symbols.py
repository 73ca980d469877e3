"""The C-ABI library loads and exports every symbol include/b200fft.h declares
(no compute calls: there is no GPU in the build container)."""
import os
import re

from conftest import ROOT


def test_library_exports_header_symbols():
    from mpi4py_fft_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'b200fft.h')).read()
    declared = set(re.findall(r'\b(b2f_[a-z_0-9]+)\s*\(', header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b2f_version() == 100
    assert _lib.launch_count() == 0


def test_bad_arguments_fail_loudly():
    import ctypes as C
    import pytest
    from mpi4py_fft_b200 import _lib
    with pytest.raises(_lib.B200FFTError):
        _lib.Plan((8, 8), (8, 8), (0,), -1, 16)           # precision 16: long double
    with pytest.raises(_lib.B200FFTError):
        _lib.Plan((8, 8), (8, 9), (0,), -1, 8)            # c2c with different shapes


def test_no_cpu_fallback_without_device():
    """allocating a device array without CUDA raises instead of computing on the host"""
    import pytest
    import torch
    from mpi4py_fft_b200.devarray import empty
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(RuntimeError):
        empty((4, 4), 'd')
