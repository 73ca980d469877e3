"""mpi4py_fft_b200 -- Blackwell-native distributed FFTs behind the mpi4py-fft API.

Drop-in for the ``PFFT.forward/backward`` path of mpi4py-fft: the names exported
here mirror /root/reference/mpi4py_fft/__init__.py:22-26 (``PFFT``, ``DistArray``,
``newDistArray``, ``fftw``); ``MPI`` is the communicator facade that replaces
``mpi4py.MPI`` on a one-process-per-GPU B200 box.  Transforms and transposes run
in ``libb200fft.so`` (hand-written sm_100a kernels + NCCL); there is no CPU path.
"""
__version__ = '0.1.0'

from . import comm as MPI
from .comm import COMM_WORLD, COMM_SELF, Comm, Compute_dims
from .devarray import DeviceArray
from .distarray import DistArray, newDistArray, Function
from .mpifft import PFFT
from .pencil import Pencil, Subcomm, Transfer
from .libfft import FFT
from . import fftw


def init(backend=None):
    """Join the job started by ``torchrun`` (one process per GPU): selects
    ``cuda:LOCAL_RANK`` and initialises ``torch.distributed`` (NCCL on GPUs, gloo
    otherwise).  A no-op for a single process without torchrun's environment."""
    import os
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            kw['device_id'] = torch.device('cuda', torch.cuda.current_device())
        dist.init_process_group(backend=backend, **kw)
    return COMM_WORLD
