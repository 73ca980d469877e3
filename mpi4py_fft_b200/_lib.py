"""ctypes binding of libb200fft.so (C ABI in include/b200fft.h).

This is the only door from Python to the device code.  There is no fallback:
if the shared library is missing or a call fails, an exception is raised --
the product never computes a transform or a transpose on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libb200fft.so')

# every symbol include/b200fft.h declares: name -> (restype, argtypes)
_i64p = C.POINTER(C.c_int64)
_intp = C.POINTER(C.c_int)
SYMBOLS = {
    'b2f_version': (C.c_int, []),
    'b2f_last_error': (C.c_char_p, []),
    'b2f_launch_count': (C.c_int64, []),
    'b2f_set_option': (C.c_int, [C.c_char_p, C.c_int64]),
    'b2f_get_option': (C.c_int64, [C.c_char_p]),
    'b2f_planxfftn': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, _i64p, _i64p, C.c_int, _intp, _intp,
                                C.c_int, C.c_uint]),
    'b2f_execute': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    'b2f_destroy_plan': (C.c_int, [C.c_void_p]),
    'b2f_plan_describe': (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    'b2f_plan_set_truncation': (C.c_int, [C.c_void_p, C.c_int64]),
    'b2f_pad_truncate': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_double, C.c_void_p]),
    'b2f_comm_unique_id': (C.c_int, [C.c_void_p]),
    'b2f_comm_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]),
    'b2f_comm_destroy': (C.c_int, [C.c_void_p]),
    'b2f_transfer_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, _i64p,
                                      C.c_int, _i64p, C.c_int, _i64p, C.c_int]),
    'b2f_transfer_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b2f_transfer_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b2f_transfer_destroy': (C.c_int, [C.c_void_p]),
    'b2f_transfer_geometry': (C.c_int, [C.c_void_p, _i64p, _i64p, _i64p, _i64p]),
    'b2f_transfer_pack': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b2f_transfer_unpack': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'b2f_malloc': (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    'b2f_free': (C.c_int, [C.c_void_p]),
    'b2f_ipc_export': (C.c_int, [C.c_void_p, C.c_void_p]),
    'b2f_ipc_open': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    'b2f_ipc_close': (C.c_int, [C.c_void_p]),
    'b2f_transfer_put': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    'b2f_transfer_exchange_p2p': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    'b2f_transfer_set_flags': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    'b2f_transfer_barrier': (C.c_int, [C.c_void_p, C.c_void_p]),
    'b2f_execute_chunk': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int64, C.c_int64,
                                    C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    'b2f_execute_scatter_chunk': (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                            C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                            C.c_int64, C.c_int, C.c_void_p]),
    'b2f_plan_can_scatter': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'b2f_execute_scatter': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                      C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


class B200FFTError(RuntimeError):
    """A libb200fft.so call returned a non-zero status."""


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200FFTError(
                "%s not found: build it with `python -m mpi4py_fft_b200.build` "
                "(there is no CPU fallback for the B200 path)" % LIB_PATH)
        # torch first: it loads the NCCL copy that transfer.cu then binds with dlopen
        try:
            import torch  # noqa: F401
        except Exception:  # pragma: no cover
            pass
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().b2f_last_error()
        raise B200FFTError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ''))


def _i64(seq):
    return (C.c_int64 * len(seq))(*[int(s) for s in seq])


def _ints(seq):
    return (C.c_int * len(seq))(*[int(s) for s in seq])


def launch_count():
    return int(lib().b2f_launch_count())


def set_option(key, value):
    check(lib().b2f_set_option(key.encode(), int(value)), 'b2f_set_option')


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------
# serial transform plan
# ---------------------------------------------------------------------------
class Plan(object):
    """b2f_plan: the counterpart of the fftw_plan held by the reference's Cython
    FFT object (/root/reference/mpi4py_fft/fftw/fftw_xfftn.pyx:109-163)."""

    def __init__(self, sizes_in, sizes_out, axes, kind, precision, flags=0):
        self._h = C.c_void_p()
        kind = [kind] if isinstance(kind, (int, np.integer)) else list(kind)
        if len(kind) < len(axes):
            kind = kind + [kind[-1]] * (len(axes) - len(kind))
        check(lib().b2f_planxfftn(C.byref(self._h), len(sizes_in), _i64(sizes_in), _i64(sizes_out),
                                  len(axes), _ints(axes), _ints(kind), int(precision), int(flags)),
              'b2f_planxfftn')
        if not self._h:
            raise RuntimeError("Failure creating B200 FFT plan")

    def execute(self, in_ptr, out_ptr, scale=1.0, stream=None):
        check(lib().b2f_execute(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), float(scale),
                                stream if stream is not None else current_stream_ptr()),
              'b2f_execute')

    def set_truncation(self, n_keep):
        """fold the dealiasing step into the transform (b2f_plan_set_truncation); False when this
        plan's kernels have no such flavour"""
        return lib().b2f_plan_set_truncation(self._h, int(n_keep)) == 0

    def can_scatter(self, transfer_handle, direction):
        """True when the stage's last pass can store into the owners' windows of
        ``transfer_handle`` (b2f_plan_can_scatter)."""
        return bool(lib().b2f_plan_can_scatter(self._h, transfer_handle._h, int(direction)))

    def execute_scatter(self, in_ptr, work_ptr, scale, transfer_handle, direction, peer_ptrs, sync=True, stream=None):
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        check(lib().b2f_execute_scatter(self._h, C.c_void_p(in_ptr), C.c_void_p(work_ptr) if work_ptr else None,
                                        float(scale), transfer_handle._h, int(direction), arr, 1 if sync else 0,
                                        stream if stream is not None else current_stream_ptr()),
              'b2f_execute_scatter')

    def execute_chunk(self, in_ptr, out_ptr, scale, mode, begin, count, view_outer=0, view_ostride=0, grid_cap=0,
                      stream=None):
        """partial launch of a one-axis stage (b2f_execute_chunk)"""
        check(lib().b2f_execute_chunk(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), float(scale), int(mode),
                                      int(begin), int(count), int(view_outer), int(view_ostride), int(grid_cap),
                                      stream if stream is not None else current_stream_ptr()), 'b2f_execute_chunk')

    def execute_scatter_chunk(self, in_ptr, scale, transfer_handle, direction, peer_ptrs, sync_flags, mode, begin,
                              count, view_outer=0, view_ostride=0, grid_cap=0, stream=None):
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        check(lib().b2f_execute_scatter_chunk(self._h, C.c_void_p(in_ptr), float(scale), transfer_handle._h,
                                              int(direction), arr, int(sync_flags), int(mode), int(begin), int(count),
                                              int(view_outer), int(view_ostride), int(grid_cap),
                                              stream if stream is not None else current_stream_ptr()),
              'b2f_execute_scatter_chunk')

    def describe(self):
        buf = C.create_string_buffer(4096)
        check(lib().b2f_plan_describe(self._h, buf, 4096), 'b2f_plan_describe')
        return buf.value.decode()

    def destroy(self):
        if self._h:
            lib().b2f_destroy_plan(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def pad_truncate(mode, half_spectrum, precision, src_ptr, dst_ptr, outer, n_src, n_dst, inner, scale=1.0):
    """b2f_pad_truncate on the current stream (mode 0 truncate, 1 pad)"""
    check(lib().b2f_pad_truncate(int(mode), 1 if half_spectrum else 0, int(precision), C.c_void_p(src_ptr),
                                 C.c_void_p(dst_ptr), int(outer), int(n_src), int(n_dst), int(inner), float(scale),
                                 current_stream_ptr()), 'b2f_pad_truncate')


# ---------------------------------------------------------------------------
# NCCL communicators, one per distinct group of ranks, built on demand
# ---------------------------------------------------------------------------
_nccl_comms = {}


def nccl_comm_for(comm):
    """b2f_comm for the ranks of ``comm`` (a :class:`..comm.Comm`); the unique
    id travels through torch.distributed's host-side object broadcast."""
    ranks = tuple(comm.ranks)
    if len(ranks) == 1:
        return None
    key = ranks
    if key in _nccl_comms:
        return _nccl_comms[key]
    import torch
    idbuf = C.create_string_buffer(128)
    if comm.Get_rank() == 0:
        check(lib().b2f_comm_unique_id(idbuf), 'b2f_comm_unique_id')
    raw = comm.bcast(bytes(idbuf.raw), root=0)
    idbuf = C.create_string_buffer(raw, 128)
    h = C.c_void_p()
    torch.cuda.synchronize()
    check(lib().b2f_comm_create(C.byref(h), idbuf, len(ranks), comm.Get_rank()), 'b2f_comm_create')
    _nccl_comms[key] = h
    return h


class TransferHandle(object):
    """b2f_transfer: device pack -> NCCL all-to-all(v) -> device unpack."""

    def __init__(self, comm, shape, itemsize, subshapeA, axisA, subshapeB, axisB, exchange=True):
        """``exchange=False`` gives a handle without an NCCL communicator: it
        serves geometry / pack / unpack only (host-side parity tests, and the
        single-GPU tests that play every rank of a group in turn)."""
        self.nranks = comm.Get_size()
        self.rank = comm.Get_rank()
        self._h = C.c_void_p()
        ncomm = nccl_comm_for(comm) if (exchange and self.nranks > 1) else None
        check(lib().b2f_transfer_create(C.byref(self._h), ncomm, self.nranks, self.rank, len(shape),
                                        _i64(shape), int(itemsize), _i64(subshapeA), int(axisA),
                                        _i64(subshapeB), int(axisB)), 'b2f_transfer_create')

    def geometry(self):
        n = self.nranks
        arrs = [(C.c_int64 * n)() for _ in range(4)]
        check(lib().b2f_transfer_geometry(self._h, *arrs), 'b2f_transfer_geometry')
        return dict(send_counts=list(arrs[0]), send_offsets=list(arrs[1]),
                    recv_counts=list(arrs[2]), recv_offsets=list(arrs[3]))

    def forward(self, arrayA, arrayB):
        from .devarray import device_ptr
        check(lib().b2f_transfer_forward(self._h, C.c_void_p(device_ptr(arrayA)),
                                         C.c_void_p(device_ptr(arrayB)), current_stream_ptr()),
              'b2f_transfer_forward')

    def backward(self, arrayB, arrayA):
        from .devarray import device_ptr
        check(lib().b2f_transfer_backward(self._h, C.c_void_p(device_ptr(arrayB)),
                                          C.c_void_p(device_ptr(arrayA)), current_stream_ptr()),
              'b2f_transfer_backward')

    def put(self, direction, src, peer_ptrs):
        """the peer-memory kernel alone (no ordering against the peers)"""
        from .devarray import device_ptr
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        check(lib().b2f_transfer_put(self._h, int(direction), C.c_void_p(device_ptr(src)), arr,
                                     current_stream_ptr()), 'b2f_transfer_put')

    def exchange_p2p(self, direction, src, peer_ptrs):
        """group barrier -> put kernel -> group barrier on the current stream"""
        from .devarray import device_ptr
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        check(lib().b2f_transfer_exchange_p2p(self._h, int(direction), C.c_void_p(device_ptr(src)), arr,
                                              current_stream_ptr()), 'b2f_transfer_exchange_p2p')

    def set_flags(self, peer_flag_ptrs):
        """arrival counters of every group rank as mapped here: the group barrier of the
        peer-memory path becomes one small kernel instead of an NCCL all-reduce"""
        if peer_flag_ptrs is None:
            check(lib().b2f_transfer_set_flags(self._h, None), 'b2f_transfer_set_flags')
            return
        arr = (C.c_void_p * len(peer_flag_ptrs))(*[int(p) for p in peer_flag_ptrs])
        check(lib().b2f_transfer_set_flags(self._h, arr), 'b2f_transfer_set_flags')

    def barrier(self):
        check(lib().b2f_transfer_barrier(self._h, current_stream_ptr()), 'b2f_transfer_barrier')

    def pack(self, direction, src, packed):
        from .devarray import device_ptr
        check(lib().b2f_transfer_pack(self._h, int(direction), C.c_void_p(device_ptr(src)),
                                      C.c_void_p(device_ptr(packed)), current_stream_ptr()),
              'b2f_transfer_pack')

    def unpack(self, direction, packed, dst):
        from .devarray import device_ptr
        check(lib().b2f_transfer_unpack(self._h, int(direction), C.c_void_p(device_ptr(packed)),
                                        C.c_void_p(device_ptr(dst)), current_stream_ptr()),
              'b2f_transfer_unpack')

    def destroy(self):
        if self._h:
            lib().b2f_transfer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# peer-memory windows: cudaMalloc'ed by the library, exported / mapped with CUDA IPC
# ---------------------------------------------------------------------------
class _RawCuda(object):
    """__cuda_array_interface__ carrier so that torch can view library memory"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


_live_windows = None


def _close_windows_at_exit():
    """interpreter exit without PFFT.destroy(): unmap the peers' windows while the CUDA
    context is still alive; the own windows are left to process teardown (see Window.release)"""
    for w in list(_live_windows or ()):
        try:
            w.release()
        except Exception:
            pass


class Window(object):
    """A byte buffer in HBM that peers of this node can map (b2f_malloc +
    b2f_ipc_export); ``tensor`` is a uint8 torch view of it."""

    def __init__(self, nbytes):
        import torch
        global _live_windows
        if _live_windows is None:
            import atexit
            import weakref
            _live_windows = weakref.WeakSet()
            atexit.register(_close_windows_at_exit)
        _live_windows.add(self)
        self.nbytes = int(max(nbytes, 16))
        p = C.c_void_p()
        check(lib().b2f_malloc(C.byref(p), self.nbytes), 'b2f_malloc')
        self.ptr = int(p.value)
        self._carrier = _RawCuda(self.ptr, self.nbytes)
        self.tensor = torch.as_tensor(self._carrier, device=torch.device('cuda', torch.cuda.current_device()))
        assert self.tensor.data_ptr() == self.ptr
        self._peers = []
        self._exported = False

    def zero(self):
        """clear the window and wait: peers may map (and poll) it right after"""
        import torch
        self.tensor.zero_()
        torch.cuda.synchronize()

    def handle(self):
        self._exported = True
        buf = C.create_string_buffer(64)
        check(lib().b2f_ipc_export(C.c_void_p(self.ptr), buf), 'b2f_ipc_export')
        return bytes(buf.raw)

    def open_peer(self, handle_bytes):
        buf = C.create_string_buffer(handle_bytes, 64)
        p = C.c_void_p()
        check(lib().b2f_ipc_open(buf, C.byref(p)), 'b2f_ipc_open')
        self._peers.append(int(p.value))
        return int(p.value)

    def close_peers(self):
        """unmap the peers' windows (do this on every rank before any rank frees)"""
        if self._peers:
            import torch
            torch.cuda.synchronize()
            for q in self._peers:
                lib().b2f_ipc_close(C.c_void_p(q))
            self._peers = []

    def free(self):
        """Release the window.  COLLECTIVE use only (PFFT.destroy): the caller has made sure
        that no peer still maps or writes this memory (group barriers in _Buffers.free)."""
        if self.ptr:
            self.close_peers()
            self.tensor = None
            lib().b2f_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def release(self):
        """Uncoordinated teardown (garbage collection, interpreter exit): a window that was
        exported may still be mapped -- or being written -- by a peer, and freeing IPC-exported
        memory under a peer is undefined behaviour.  So only the peers' windows are unmapped
        here; the own allocation stays until the process (and with it every mapping) ends."""
        if self.ptr:
            self.close_peers()
            if not self._exported:
                self.tensor = None
                lib().b2f_free(C.c_void_p(self.ptr))
            else:
                import warnings
                warnings.warn("mpi4py_fft_b200: a peer-memory window was dropped without PFFT.destroy(); its %d MiB "
                              "stay allocated until the process exits (destroy() is collective and frees them safely)"
                              % (self.nbytes >> 20), ResourceWarning)
            self.ptr = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
