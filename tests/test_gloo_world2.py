"""world_size-2 (and 4) runs over gloo on the CPU: the communicator facade
(collectives, Cartesian sub-groups) and the exchange plan of Transfer -- the
product's per-peer counts/offsets drive a real all-to-all between processes,
with the pack/unpack that the CUDA kernels do on the device restated in numpy
by the test.  Checks the result against the oracle's exchange."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(a, axis, p):
    from mpi4py_fft_b200.pencil import _blockdist
    segs = []
    for i in range(p):
        n, s = _blockdist(a.shape[axis], p, i)
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(s, s + n)
        segs.append(np.ascontiguousarray(a[tuple(sl)]).ravel())
    return np.concatenate(segs)


def _unpack(buf, shape, axis, p, dtype):
    from mpi4py_fft_b200.pencil import _blockdist
    out = np.zeros(shape, dtype=dtype)
    off = 0
    for i in range(p):
        n, s = _blockdist(shape[axis], p, i)
        sl = [slice(None)] * len(shape)
        sl[axis] = slice(s, s + n)
        blk_shape = list(shape)
        blk_shape[axis] = n
        cnt = int(np.prod(blk_shape))
        out[tuple(sl)] = buf[off:off + cnt].reshape(blk_shape)
        off += cnt
    return out


def _worker(rank, world, port, shape, dtype, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import mpi4py_fft_b200 as B
        from mpi4py_fft_b200 import PFFT, COMM_WORLD, MPI
        import pfft_oracle as O

        comm = COMM_WORLD
        assert comm.Get_size() == world and comm.Get_rank() == rank
        assert comm.allreduce(rank + 1) == world * (world + 1) // 2
        assert comm.bcast('hello' if rank == 0 else None) == 'hello'
        assert comm.gather(rank) == (list(range(world)) if rank == 0 else None)
        assert comm.reduce(2.0, op=MPI.MAX) == (2.0 if rank == 0 else None)

        fft = PFFT(comm, shape, dtype=dtype)
        orc = O.OraclePFFT(world, shape, dtype=dtype)
        g = np.random.default_rng(7).random(shape).astype(dtype)
        # stage-0 outputs of every rank, according to the oracle
        blocks = [np.ascontiguousarray(g[orc.local_slice(r, False)]) for r in range(world)]
        st0 = orc.stages[0]
        cur = [(O.serial_transform(b, st0['axes'], st0['kinds_f']) * st0['M']).astype(st0['out_dtype']) for b in blocks]
        for ti, tr in enumerate(fft.transfer):
            expect = orc._exchange(orc.transfers[ti], cur)
            geo = tr.geometry
            sub = tr.comm
            p = sub.Get_size()
            assert tuple(cur[rank].shape) == tr.subshapeA
            if p == 1:
                got = cur[rank].copy()
            else:
                send = _pack(cur[rank], tr.axisA, p)
                assert [int(c) for c in geo['send_counts']] == [len(x) for x in np.split(send, np.cumsum(geo['send_counts'])[:-1])]
                recv = np.zeros(int(sum(geo['recv_counts'])), dtype=send.dtype)
                group = MPI.world().group_for(sub.ranks)
                assert group is not None
                ts, tr_ = torch.from_numpy(send.view(np.float64)), torch.from_numpy(recv.view(np.float64))
                k = send.dtype.itemsize // 8
                dist.all_to_all_single(tr_, ts, [int(c) * k for c in geo['recv_counts']],
                                       [int(c) * k for c in geo['send_counts']], group=group)
                got = _unpack(recv, tr.subshapeB, tr.axisB, p, send.dtype)
            assert np.array_equal(got, expect[rank]), "transfer %d differs on rank %d" % (ti, rank)
            # next stage on every rank (oracle) so that the following transfer has inputs
            st = orc.stages[ti + 1]
            cur = [(O.serial_transform(b, st['axes'], st['kinds_f']) * st['M']).astype(st['out_dtype']) for b in expect]
        # sub-communicator collectives
        for c in fft.subcomm:
            assert c.allreduce(1) == c.Get_size()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize('world,shape,dtype', [(2, (8, 6, 4), 'D'), (2, (9, 7, 10), 'd'), (4, (9, 8, 6), 'D')])
def test_transfer_exchange_over_gloo(world, shape, dtype):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, dtype, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, msg in results:
        assert msg == 'ok', "rank %d:\n%s" % (rank, msg)


# ---------------------------------------------------------------------------
# negotiation of the peer-memory windows (mpifft._Buffers.setup_windows): the
# host-side protocol on two CPU processes with a stand-in for the CUDA-IPC window
# ---------------------------------------------------------------------------
def _window_worker(rank, world, port, fail_rank, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group('gloo', rank=rank, world_size=world)
        from mpi4py_fft_b200 import PFFT, COMM_WORLD, _lib

        class FakeWindow(object):
            """stands in for _lib.Window: no device memory, handles are plain bytes"""
            made = []

            def __init__(self, nbytes):
                if rank == fail_rank:
                    raise RuntimeError("cudaIpcGetMemHandle: not permitted (simulated)")
                self.nbytes = int(nbytes)
                self.ptr = 0x1000 * (rank + 1) + len(FakeWindow.made)
                self.opened = []
                FakeWindow.made.append(self)

            def handle(self):
                return ('rank%d:%d' % (rank, self.ptr)).encode().ljust(64, b'.')

            def open_peer(self, h):
                assert len(h) == 64 and not h.startswith(('rank%d:' % rank).encode())
                self.opened.append(h)
                return 0x900000 + int(h.split(b':')[1].rstrip(b'.'))

            def close_peers(self):
                self.opened = []

            def free(self):
                self.ptr = 0

            def zero(self):
                pass

        _lib.Window = FakeWindow
        os.environ['B2F_FLAG_BARRIER'] = '0'      # arming the flag kernels needs the device library
        fft = PFFT(COMM_WORLD, (8, 6, 4), dtype='D')
        assert fft.forward._plan['windowed'] and fft.forward._plan['a'][-1].startswith('W')
        buf = fft._buffers
        buf.setup_windows(fft.transfer)
        nontrivial = [t for t in fft.transfer if t.comm.Get_size() > 1]
        assert len(nontrivial) == 1
        table = buf.peers[id(nontrivial[0])]
        if fail_rank is None:
            # both ranks mapped each other's windows: per label one pointer per group rank, own one in place
            assert set(table) == set(buf.need) | {'__sync__'}
            for label, ptrs in table.items():
                assert len(ptrs) == world and ptrs[rank] == buf.windows[label].ptr
                assert all(p >= 0x900000 for j, p in enumerate(ptrs) if j != rank)
        else:
            # one rank could not export: BOTH fall back to the NCCL path for this group, nobody hangs
            assert table is None
        buf.free([t.comm for t in fft.transfer])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize('fail_rank', [None, 1, 0])
def test_window_negotiation_over_gloo(fail_rank):
    """the windows protocol is collective and all-or-nothing per group: handles are
    exchanged through the host, every rank maps its peers, and a rank that cannot
    export (IPC not permitted in its container) makes the whole group keep the NCCL
    exchange instead of deadlocking or diverging"""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_window_worker, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, msg in results:
        assert msg == 'ok', "rank %d:\n%s" % (rank, msg)
