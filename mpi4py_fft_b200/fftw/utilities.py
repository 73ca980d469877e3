"""Enums and array allocation for the serial-transform front end.

Same names and integer values as the reference's Cython module
(/root/reference/mpi4py_fft/fftw/utilities.pyx:7-37): the kind integers are
passed straight through the C ABI (``include/b200fft.h``).  ``aligned`` returns
a device array: cudaMalloc'd blocks are 256-byte aligned, which subsumes the
32-byte SIMD alignment the reference arranges by hand (utilities.pyx:54-84).
"""
import numpy as np

from ..devarray import DeviceArray, empty

FFTW_FORWARD = -1
FFTW_R2HC = 0
FFTW_BACKWARD = 1
FFTW_HC2R = 1
FFTW_DHT = 2
FFTW_REDFT00 = 3
FFTW_REDFT01 = 4
FFTW_REDFT10 = 5
FFTW_REDFT11 = 6
FFTW_RODFT00 = 7
FFTW_RODFT01 = 8
FFTW_RODFT10 = 9
FFTW_RODFT11 = 10

C2C_FORWARD = -1
C2C_BACKWARD = 1
R2C = -2
C2R = 2

# planner flags are accepted for source compatibility; plans here carry no
# measured state, so they have no effect
FFTW_MEASURE = 0
FFTW_DESTROY_INPUT = 1
FFTW_UNALIGNED = 2
FFTW_CONSERVE_MEMORY = 4
FFTW_EXHAUSTIVE = 8
FFTW_PRESERVE_INPUT = 16
FFTW_PATIENT = 32
FFTW_ESTIMATE = 64
FFTW_WISDOM_ONLY = 2097152


def get_alignment(array):
    """Largest power of two <= 32 dividing the address of ``array``
    (utilities.pyx:39-52); device allocations always give 32."""
    if isinstance(array, DeviceArray):
        addr = array.tensor.data_ptr()
    elif isinstance(array, np.ndarray):
        addr = array.ctypes.data
    else:
        return 32
    for i in range(5, -1, -1):
        if addr % (1 << i) == 0:
            return 1 << i
    return 1


def aligned(shape, n=32, dtype=np.dtype('d'), fill=None):
    """New device array of ``shape``/``dtype`` (optionally filled)."""
    return empty(shape, np.dtype(dtype), fill=fill)


def aligned_like(z, fill=None):
    """New device array with the shape and dtype of ``z``."""
    return empty(tuple(z.shape), np.dtype(z.dtype), fill=fill)
