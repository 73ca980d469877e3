"""Serial transforms on the device behind the reference's ``mpi4py_fft.fftw`` names."""
from .xfftn import *          # noqa: F401,F403
from .xfftn import FFT, flag_dict, get_normalization, inverse, dct_type, idct_type, dst_type, idst_type
from .utilities import *      # noqa: F401,F403
from .factory import (get_planned_FFT, export_wisdom, import_wisdom, forget_wisdom, cleanup,
                      set_timelimit, get_fftw_lib, fftlib)
