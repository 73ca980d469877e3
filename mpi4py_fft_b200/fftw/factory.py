"""Plan factory (mirror of /root/reference/mpi4py_fft/fftw/factory.py:44-107).

``fftlib`` maps the precision letter to the executor: the reference holds up to
three compiled FFTW wrappers ('F', 'D', 'G'); here 'F' and 'D' are served by
the same CUDA library and long double ('G') does not exist on the device.
FFTW wisdom / time limits (factory.py:109-182) have no analogue -- there is no
measured planner state -- and are kept as no-ops so calling code still runs.
"""
from .utilities import FFTW_FORWARD, FFTW_MEASURE


def get_fftw_lib(dtype):
    """Executor class for precision ``dtype`` ('f'/'d'), None for 'g'."""
    from . import xfftn
    return xfftn if str(dtype).lower() in ('f', 'd') else None


class _Lib(dict):
    def __missing__(self, key):
        raise KeyError("no device executor for precision %r (long double is not available on B200)" % key)


fftlib = _Lib()
for _t in 'fd':
    fftlib[_t.upper()] = 'b200fft'


def get_planned_FFT(input_array, output_array, axes=(-1,), kind=FFTW_FORWARD,
                    threads=1, flags=(FFTW_MEASURE,), normalization=1.0):
    """Planned transform object for the given arrays (or array specs)."""
    from .xfftn import FFT
    dtype = input_array.dtype.char
    assert dtype.upper() in fftlib
    return FFT(input_array, output_array, axes, kind, threads, flags, normalization)


def export_wisdom(filename):
    """No-op: B200 plans have no wisdom."""


def import_wisdom(filename):
    """No-op: B200 plans have no wisdom."""


def forget_wisdom():
    """No-op: B200 plans have no wisdom."""


def set_timelimit(limit):
    """No-op: planning is closed-form."""


def cleanup():
    """No-op: device tables are cached for the life of the process."""
