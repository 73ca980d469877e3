"""Parallel multidimensional transforms: ``PFFT`` on B200.

Same constructor, attributes and call semantics as the reference's ``PFFT`` /
``Transform`` (/root/reference/mpi4py_fft/mpifft.py:8-79,202-419): a chain of
serial stages (one or more undivided axes each) separated by global
redistributions, forward normalised, backward not.  What changes is where the
data lives and how it moves:

  * every array is device resident; stages are ``b2f_execute`` launches and
    redistributions are ``b2f_transfer_*`` (pack -> NCCL all-to-all -> unpack),
    all enqueued on the current CUDA stream -- a whole forward or backward is
    one stream with no host round trip;
  * the reference allocates an input and an output array per stage and copies
    between them (mpifft.py:66,76; libfft.py:72-78).  Here the chain is mapped
    onto at most two work buffers plus the two end-point arrays: c2c/r2r stages
    run in place, a redistribution inside a group of one rank is an alias (no
    bytes move), the first stage reads the caller's array directly and the last
    stage writes the caller's output directly.  See ``Transform._layout``.
"""
from __future__ import annotations

import numpy as np

from .devarray import copy_out as _copy_out, ArraySpec, DeviceArray, as_tensor, np_dtype_of, torch_dtype, device
from .libfft import FFT
from .pencil import Pencil, Subcomm


_SYNC = '__sync__'       # window holding the arrival counters of the flag barriers
_SYNC_SLOT = 128         # bytes per transfer: 16 ranks x 8-byte counter
_MAX_PEERS = 16          # B2F_MAX_PEERS / B2F_PUT_MAX_PEERS of the kernels


class _Buffers(object):
    """End-point arrays X (physical) / Y (spectral) and two byte work buffers,
    shared by the forward and backward transforms of one PFFT; all lazy."""

    def __init__(self, x_spec, y_spec):
        self.spec = {'X': x_spec, 'Y': y_spec}
        self.arr = {}
        self.work = {}
        self.need = {}          # label -> bytes (max over every use, both directions)
        self.windows = None     # label -> _lib.Window once the peer-memory path is set up
        self.peers = {}         # id(Transfer) -> {label: [mapped base pointer per group rank]} | None

    def reserve(self, label, nbytes):
        self.need[label] = max(self.need.get(label, 0), int(nbytes))

    def setup_windows(self, transfers):
        """Collective over every transfer group (each rank walks its transfers in
        plan order): allocate the work buffers as IPC-exportable windows, exchange
        the handles through the host side and map the peers' windows.  A group in
        which any rank fails keeps the NCCL path for that transfer."""
        from . import _lib
        if self.windows is not None:
            return
        self.windows = {}
        handles = {}
        ok = True
        try:
            for label, nbytes in sorted(self.need.items()):
                w = _lib.Window(nbytes)
                self.windows[label] = w
                handles[label] = w.handle()
            # arrival counters of the flag barrier: 16 x 8 bytes per transfer, zeroed before anybody maps them
            w = _lib.Window(_SYNC_SLOT * max(1, len(transfers)))
            w.zero()
            self.windows[_SYNC] = w
            handles[_SYNC] = w.handle()
        except Exception as exc:   # e.g. IPC not permitted in this container
            ok = False
            handles = {'error': repr(exc)[:200]}
        opened = {}
        for ti, t in enumerate(transfers):
            comm = t.comm
            if comm.Get_size() == 1:
                continue
            if comm.Get_size() > _MAX_PEERS:
                # the peer-store and put kernels address at most 16 owners: larger groups keep pack + NCCL + unpack
                self.peers[id(t)] = None
                continue
            everyone = comm.allgather((ok, handles))
            me = comm.Get_rank()
            good = all(e[0] for e in everyone)
            table = None
            if good:
                try:
                    table = {}
                    for label, w in self.windows.items():
                        ptrs = []
                        for j, (_, hs) in enumerate(everyone):
                            if j == me:
                                ptrs.append(w.ptr)
                            else:
                                key = hs[label]
                                if key not in opened:
                                    opened[key] = w.open_peer(key)
                                ptrs.append(opened[key])
                        table[label] = ptrs
                except Exception:
                    table = None
            # the decision must be the same on every rank of the group
            agreed = all(comm.allgather(table is not None))
            self.peers[id(t)] = table if agreed else None
            if agreed and flag_barrier_enabled():
                # group barrier = one small kernel on counters in peer memory instead of an NCCL all-reduce;
                # all ranks of the group use it or none does
                handle = t._plan()
                try:
                    handle.set_flags([ptr + ti * _SYNC_SLOT for ptr in table[_SYNC]])
                    mine = True
                except Exception:
                    mine = False
                if not all(comm.allgather(mine)) and mine:
                    handle.set_flags(None)
            if not agreed and me == 0:
                import warnings
                reasons = [e[1].get('error') for e in everyone if isinstance(e[1], dict) and e[1].get('error')]
                warnings.warn("mpi4py_fft_b200: peer-memory windows are not available in this group (%s); "
                              "its redistributions use pack + NCCL + unpack" % (reasons[0] if reasons else
                                                                                "a peer window could not be mapped"),
                              RuntimeWarning)

    def window_view(self, label, shape, dtype):
        import torch
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        w = self.windows[label]
        assert nbytes <= w.nbytes, "work buffer %s smaller than a planned use" % label
        return DeviceArray(w.tensor[:nbytes].view(torch_dtype(dtype)).view(tuple(shape)))

    def free(self, groups=()):
        """``groups``: the transfer communicators -- when given (PFFT.destroy is
        collective, as the reference's) every rank unmaps its peers' windows before
        any rank releases its own."""
        if self.windows is not None:      # set up (or attempted) on every rank alike: the barriers below must match
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            for comm in groups:               # nobody is still storing into a window
                if comm.Get_size() > 1:
                    comm.Barrier()
            for w in self.windows.values():
                w.close_peers()
            for comm in groups:               # every mapping is gone before any owner frees
                if comm.Get_size() > 1:
                    comm.Barrier()
            for w in self.windows.values():
                w.free()
        self.windows = None
        self.peers = {}
        self.work = {}

    def endpoint(self, name):
        if name not in self.arr:
            self.arr[name] = self.spec[name].allocate(fill=0)
        return self.arr[name]

    def view(self, name, shape, dtype):
        """Typed view of work buffer ``name`` (grown on demand)."""
        import torch
        if self.windows and name in self.windows:
            return self.window_view(name, shape, dtype)
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        buf = self.work.get(name)
        if buf is None or buf.numel() < nbytes:
            self.work[name] = None
            buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device())
            self.work[name] = buf
        return DeviceArray(buf[:nbytes].view(torch_dtype(dtype)).view(tuple(shape)))


class Transform(object):
    """One direction of a parallel transform (forward or backward).

    ``xfftn`` are the per-stage callables, ``transfer`` the redistributions
    between them (bound methods of :class:`Transfer`), ``pencil`` the input and
    output pencils; as in the reference ``len(xfftn) == len(transfer) + 1``.
    """

    def __init__(self, xfftn, transfer, pencil, buffers=None, ends=('X', 'Y')):
        assert len(xfftn) == len(transfer) + 1 and len(pencil) == 2
        self._xfftn = tuple(xfftn)
        self._transfer = tuple(transfer)
        self._pencil = tuple(pencil)
        if buffers is None:
            buffers = _Buffers(ArraySpec(xfftn[0].input_shape, xfftn[0].input_dtype),
                               ArraySpec(xfftn[-1].output_shape, xfftn[-1].output_dtype))
            ends = ('X', 'Y')
        self._buffers = buffers
        self._ends = ends
        self._all_transfers = self._transfer   # PFFT replaces this with the forward-order list (collective order)
        self._side = None                      # second stream of the pipelined redistribution
        self._plan = self._layout()
        self._merged = self._merge()
        self._marks = None                     # profiling: a list collects (stage indices, CUDA event) per stage

    # -- arrays -------------------------------------------------------------------
    @property
    def input_array(self):
        return self._buffers.endpoint(self._ends[0])

    @property
    def output_array(self):
        return self._buffers.endpoint(self._ends[1])

    @property
    def input_pencil(self):
        return self._pencil[0]

    @property
    def output_pencil(self):
        return self._pencil[1]

    # -- static buffer assignment ---------------------------------------------------
    def _layout(self):
        """Where every intermediate of the chain lives.

        Logical arrays: a_i (input of stage i), b_i (its output), with
        a_{i+1} = T_i(b_i).  Labels: 'IN' (the caller's / end-point input, never
        written), 'OUT' (the final destination), 'W0'/'W1' (work buffers).
          * a_{i+1} shares b_i's buffer when T_i acts inside a group of one;
          * b_i shares a_i's buffer when the stage can run in place;
          * 'OUT' is propagated backwards from the last stage as far as those
            two rules allow, so the tail of the chain needs no extra copy.
        """
        m = len(self._xfftn)
        inplace = []
        for st in self._xfftn:
            inplace.append(st.input_shape == st.output_shape and st.input_dtype == st.output_dtype)
        trivial = [getattr(t, '__self__', t).comm.Get_size() == 1 for t in self._transfer]
        # peer-memory transfers store into plan-owned windows only: the caller's
        # output array is then written by the last stage (out of place), never by
        # a transfer, so a non-trivial transfer stops the 'OUT' propagation
        windowed = p2p_enabled() and not all(trivial)
        a = [None] * m
        b = [None] * m
        a[0] = 'IN'
        b[m - 1] = 'OUT'
        i = m - 1
        while i > 0 and inplace[i]:
            if windowed and not trivial[i - 1]:
                break
            a[i] = 'OUT'
            if not trivial[i - 1]:
                break
            b[i - 1] = 'OUT'
            i -= 1
        toggle = 0
        for i in range(m):
            if i > 0 and a[i] is None:
                if trivial[i - 1]:
                    a[i] = b[i - 1]
                else:
                    toggle ^= 1
                    a[i] = 'W%d' % toggle
            if b[i] is None:
                if inplace[i] and a[i] not in ('IN',):
                    b[i] = a[i]
                else:
                    toggle ^= 1
                    b[i] = 'W%d' % toggle
                    if b[i] == a[i]:   # cannot happen with two buffers, but stay safe
                        toggle ^= 1
                        b[i] = 'W%d' % toggle
        for i, st in enumerate(self._xfftn):
            for label, shp, dt in ((a[i], st.input_shape, st.input_dtype), (b[i], st.output_shape, st.output_dtype)):
                if label not in ('IN', 'OUT'):
                    self._buffers.reserve(label, int(np.prod(shp)) * np.dtype(dt).itemsize)
        return dict(a=a, b=b, trivial=trivial, windowed=windowed)

    def _merge(self):
        """One planned transform over the axes of ALL stages, or None.

        When no redistribution moves data (every transfer acts inside a group of one
        rank: a single GPU, or axes that are not divided) and every stage is a plain
        c2c stage on the same block, the chain is one multi-axis transform.  Handing it
        to the library as such lets it run the rotating schedule (csrc/fft_rot.cuh:
        whole-pencil bulk loads, page-local stores) instead of one strided pass per
        axis; results are those of the stage-by-stage chain (c2c axes commute).  The
        reference offers the same merge to the user as ``collapse=True``
        (mpifft.py:255-272); here it is an execution detail, ``xfftn`` / ``transfer`` /
        ``pencil`` keep the reference's structure."""
        import os
        from .fftw.utilities import FFTW_FORWARD, FFTW_BACKWARD
        from .fftw.xfftn import FFT as Planned
        from .libfft import _Stage
        if os.environ.get('B2F_MERGE', '1') in ('0', 'false', 'no', ''):
            return None
        if len(self._xfftn) < 2 or not all(self._plan['trivial']):
            return None
        first = self._xfftn[0]
        axes, kind = [], None
        for st in self._xfftn:
            if type(st) is not _Stage:
                return None
            pl = st._planned
            if pl.kind not in (FFTW_FORWARD, FFTW_BACKWARD) or any(k != pl.kind for k in pl.kinds):
                return None
            if kind is not None and pl.kind != kind:
                return None
            kind = pl.kind
            if tuple(st.input_shape) != tuple(first.input_shape) or tuple(st.output_shape) != tuple(first.input_shape) \
                    or st.input_dtype != first.input_dtype or st.output_dtype != first.input_dtype:
                return None
            axes += [int(a) for a in pl.axes]
        if len(set(axes)) != len(axes):
            return None
        spec = ArraySpec(first.input_shape, first.input_dtype)
        return Planned(spec, ArraySpec(first.input_shape, first.input_dtype), axes, kind)

    # -- execution -------------------------------------------------------------------
    def _fused(self, i, st, tr, direction):
        """can stage i store straight into the windows of transfer i? (decided once;
        it depends on geometry only, so every rank of the group agrees)"""
        key = ('fused', i)
        if key not in self._plan:
            ok = fused_enabled() and hasattr(st, 'can_scatter')
            if ok:
                handle = tr._plan()       # collective (creates the group's communicator): never inside a try
                try:
                    ok = st.can_scatter(handle, direction)
                except Exception:
                    ok = False
            self._plan[key] = ok
        return self._plan[key]

    def _pipeline(self, i):
        """Chunk plan for overlapping stage i (+ its redistribution) with stage i+1, or
        None.  The producing stage stores chunk c into the peers' windows while the
        consuming stage already transforms chunk c-1 on a second stream; chunks are
        ranges of an array axis that neither stage transforms and the transfer does not
        touch: the last axis (inner ranges, re-viewed rows where other axes lie
        between) or the first one (outer ranges).  Geometry only: every rank of the
        group takes the same decision."""
        key = ('pipe', i)
        if key in self._plan:
            return self._plan[key]
        plan = None
        m = len(self._xfftn)
        prod, cons = self._xfftn[i], self._xfftn[i + 1]
        k = pipeline_chunks(int(np.prod(prod.output_shape)) * np.dtype(prod.output_dtype).itemsize)
        follows_trivially = (i + 2 >= m) or self._plan['trivial'][i + 1]
        if k > 1 and follows_trivially and getattr(prod, 'chunkable', False) and getattr(cons, 'chunkable', False):
            ps, cs = tuple(prod.output_shape), tuple(cons.input_shape)
            nd = len(ps)
            ap, ac = prod._planned.axes[0], cons._planned.axes[0]
            pr = lambda t: int(np.prod(t)) if len(t) else 1
            if nd >= 3 and ap != nd - 1 and ac != nd - 1 and ps[-1] == cs[-1] and ps[-1] >= k:
                # ranges of the last axis
                def view(shape, ax):
                    mid = pr(shape[ax + 1:-1])
                    if mid == 1:
                        return (0, 0)                       # plain inner range
                    if pr(shape[:ax]) != 1:
                        return None                         # rows before AND between: not expressible
                    return (mid, shape[-1])
                vp, vc = view(ps, ap), view(cs, ac)
                if vp is not None and vc is not None:
                    # cut at multiples of 16 elements (tile rows stay aligned) when the axis is long enough
                    g = 16 if ps[-1] >= 16 * k else 1
                    cuts = sorted(set([0, ps[-1]] + [(ps[-1] // g * j // k) * g for j in range(1, k)]))
                    plan = [((1, lo, hi - lo) + vp, (1, lo, hi - lo) + vc) for lo, hi in zip(cuts[:-1], cuts[1:])]
            if plan is None and ap != 0 and ac != 0 and ps[0] == cs[0] and ps[0] >= k:
                # ranges of the first axis = ranges of the outer index of both stages
                rp, rc = pr(ps[1:ap]), pr(cs[1:ac])
                cuts = [ps[0] * j // k for j in range(k + 1)]
                plan = [((2, lo * rp, (hi - lo) * rp, 0, 0), (2, lo * rc, (hi - lo) * rc, 0, 0))
                        for lo, hi in zip(cuts[:-1], cuts[1:])]
        self._plan[key] = plan
        return plan

    def _run_pipelined(self, i, chunks, cur, recv, dst2, tr, direction, peers, normalize):
        """stage i in chunks on the current stream (each followed by the group
        barrier), stage i+1 chunk by chunk on a side stream"""
        import torch
        prod, cons = self._xfftn[i], self._xfftn[i + 1]
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        p = tr.comm.Get_size()
        nsm = sm_count()
        sms = pipeline_producer_sms(p, nsm)
        handle = tr._plan()
        # events are reused from call to call (one per chunk plus the join): no allocation on the hot path
        evs = self._plan.get(('events', i))
        if evs is None or len(evs) != len(chunks) + 1:
            evs = self._plan[('events', i)] = [torch.cuda.Event() for _ in range(len(chunks) + 1)]
        for c, (pspec, cspec) in enumerate(chunks):
            flags = (1 if c == 0 else 0) | 2          # enter once, leave after every chunk
            prod.run_scatter_chunk(cur, normalize, handle, direction, peers, flags, pspec, grid_cap=sms)
            evs[c].record(main)
            side.wait_event(evs[c])
            with torch.cuda.stream(side):
                # the last chunk has the GPU to itself; the others share it with the producer
                last = c + 1 == len(chunks)
                cons.run_chunk(recv, dst2, normalize, cspec, grid_cap=0 if last else max(8, nsm - sms))
        evs[-1].record(side)
        main.wait_event(evs[-1])

    def _resolve(self, label, shape, dtype, src, out):
        if label == 'IN':
            return src
        if label == 'OUT':
            return out
        return self._buffers.view(label, shape, dtype)

    def __call__(self, input_array=None, output_array=None, **kw):
        """Compute the transform.

        ``input_array`` / ``output_array`` may be device arrays of the planned
        shape and dtype (used directly), host arrays (staged through the
        plan-owned end-point arrays) or omitted (the end-point arrays are used
        and the output end point is returned, as in the reference).
        ``normalize=True/False`` is forwarded to every stage.
        """
        first, last = self._xfftn[0], self._xfftn[-1]
        src = _usable(input_array, first.input_shape, first.input_dtype)
        if src is None:
            src = self.input_array
            if input_array is not None:
                src[...] = input_array
        out = _usable(output_array, last.output_shape, last.output_dtype)
        direct_out = out is not None
        if out is None:
            out = self.output_array

        plan = self._plan
        m = len(self._xfftn)
        if plan['windowed'] and self._buffers.windows is None:
            self._buffers.setup_windows([getattr(t, '__self__', t) for t in self._all_transfers])
        if getattr(first, 'destroys_input', False) and plan['a'][0] == 'IN' and input_array is not None \
                and src is not self.input_array:
            # a multi-axis c2r stage overwrites what it reads: keep the caller's array intact
            self.input_array[...] = src
            src = self.input_array
        marks = self._marks
        if marks is not None:
            marks.append(((), _mark()))
        if self._merged is not None:
            scale = 1.0
            for st in self._xfftn:
                scale *= st.scale_for(kw.get('normalize'))
            self._merged.execute(src, out, scale)
            if marks is not None:
                marks.append((tuple(range(m)), _mark()))
            if output_array is not None and not direct_out:
                _copy_out(out, output_array)
                return output_array
            return out
        cur = src
        skip = -1
        for i in range(m):
            if i == skip:
                # stage i already ran inside the pipelined redistribution before it; what follows
                # it is a trivial transfer (an alias) or nothing
                continue
            st = self._xfftn[i]
            dst = self._resolve(plan['b'][i], st.output_shape, st.output_dtype, src, out)
            if i + 1 < m and not plan['trivial'][i]:
                nxt = self._xfftn[i + 1]
                recv = self._resolve(plan['a'][i + 1], nxt.input_shape, nxt.input_dtype, src, out)
                tr = getattr(self._transfer[i], '__self__', None)
                table = self._buffers.peers.get(id(tr)) if tr is not None else None
                label = plan['a'][i + 1]
                direction = 0 if self._transfer[i].__name__ == 'forward' else 1
                if table is not None and label in table:
                    chunks = self._pipeline(i) if self._fused(i, st, tr, direction) else None
                    if chunks:
                        # stage i, the redistribution and stage i+1 overlapped chunk by chunk
                        dst2 = self._resolve(plan['b'][i + 1], nxt.output_shape, nxt.output_dtype, src, out)
                        self._run_pipelined(i, chunks, cur, recv, dst2, tr, direction, table[label], kw.get('normalize'))
                        cur = dst2
                        skip = i + 1
                        if marks is not None:
                            marks.append(((i, i + 1), _mark()))
                        continue
                    if self._fused(i, st, tr, direction):
                        # one launch: the stage's last pass stores into the owners' windows
                        st.run_scatter(cur, dst, kw.get('normalize'), tr._plan(), direction, table[label])
                    else:
                        st.run(cur, dst, kw.get('normalize'))
                        tr.exchange_p2p(direction, dst, recv, table[label])
                else:
                    st.run(cur, dst, kw.get('normalize'))
                    self._transfer[i](dst, recv)
                cur = recv
            else:
                st.run(cur, dst, kw.get('normalize'))
                cur = dst
            if marks is not None:
                marks.append(((i,), _mark()))

        if output_array is not None and not direct_out:
            _copy_out(out, output_array)
            return output_array
        return out


def _mark():
    """timing event on the current stream (Transform._marks)"""
    import torch
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(torch.cuda.current_stream())
    return ev


def p2p_enabled():
    """Peer-memory transfers (one put kernel over NVLink instead of pack -> NCCL ->
    unpack) are the default on a multi-GPU node; B2F_P2P=0 keeps the NCCL path."""
    import os
    return os.environ.get('B2F_P2P', '1') not in ('0', 'false', 'no', '')


def pipeline_chunks(block_bytes=None):
    """Chunks of the pipelined redistribution.  B2F_PIPELINE=K forces K (0 or 1: off);
    by default chunks of about 1 GiB (measured best on 2 and 4 B200s, profiles/
    r1e_pipeline.txt: 8 GiB blocks K=8, 4 GiB blocks K=4) and no pipelining below
    64 MiB, where the per-chunk group barrier costs more than the overlap gains."""
    import os
    env = os.environ.get('B2F_PIPELINE')
    if env is not None:
        try:
            return int(env)
        except ValueError:
            return 0
    if block_bytes is None or block_bytes < (64 << 20):
        return 0
    if flag_barrier_enabled():
        # a chunk barrier is one small kernel: at least 4 chunks even for small blocks (the tail of the
        # consuming stage that nothing overlaps is 1/K of it); more than 8 did not pay on 2 or 4 GPUs
        # (profiles/r2_pipeline.txt)
        return int(max(4, min(8, block_bytes >> 30)))
    return int(max(2, min(16, block_bytes >> 30)))


def sm_count():
    """SMs of the current device, from the library (cudaDevAttrMultiProcessorCount)"""
    from . import _lib
    n = int(_lib.lib().b2f_get_option(b'sm_count'))
    return n if n > 0 else 148


def pipeline_producer_sms(p, nsm=148):
    """SMs given to the producing (NVLink-bound) stage of a pipelined redistribution
    in a group of p ranks; the consuming stage gets the rest.  The remote share
    (p-1)/p of the stage output crosses NVLink at about 1/8 of the HBM rate, so the
    producer needs roughly 0.3 p/(p-1) of the GPU to keep the links busy."""
    import os
    env = os.environ.get('B2F_PIPE_SMS')
    if env:
        return int(env)
    return int(min(0.8 * nsm, max(0.2 * nsm, round(nsm * 0.30 * p / (p - 1)))))


def flag_barrier_enabled():
    """Group barriers of the peer-memory path as flag kernels (B2F_FLAG_BARRIER=0: NCCL all-reduce)."""
    import os
    return os.environ.get('B2F_FLAG_BARRIER', '1') not in ('0', 'false', 'no', '')


def fused_enabled():
    """Stage + redistribution in one kernel (B2F_FUSED=0: stage, then put kernel)."""
    import os
    return os.environ.get('B2F_FUSED', '1') not in ('0', 'false', 'no', '')


def _usable(a, shape, dtype):
    """``a`` as a device array that a kernel can read/write directly, or None."""
    if a is None or isinstance(a, np.ndarray):
        return None
    try:
        t = as_tensor(a)
    except TypeError:
        return None
    if tuple(t.shape) != tuple(shape) or np_dtype_of(a) != np.dtype(dtype):
        return None
    if not t.is_cuda or not t.is_contiguous():
        return None
    return a


class PFFT(object):
    """Parallel FFT (and r2r) over a block-distributed array on B200 GPUs.

    Parameters are the reference's (mpifft.py:82-204): ``comm`` (a communicator
    of this package, a :class:`Subcomm`, or a Cartesian communicator), ``shape``,
    ``axes`` (None | int | sequence of ints | sequence of sequences),
    ``dtype``, ``grid``, ``padding`` (False or one factor per axis), ``collapse``,
    ``backend`` (ignored: device kernels), ``transforms`` (axes tuple ->
    (forward planner, backward planner) from :mod:`mpi4py_fft_b200.fftw`),
    ``darray``, ``slab`` (deprecated).

    ``forward(input_array=None, output_array=None, **kw)`` and ``backward(...)``
    are :class:`Transform` instances.
    """

    def __init__(self, comm, shape=None, axes=None, dtype=float, grid=None, padding=False,
                 collapse=False, backend='b200', transforms=None, darray=None, **kw):
        if shape is None:
            assert darray is not None
            shape = darray.pencil.shape
        ndim = len(shape)

        # -- axes -> list of groups (tuples), one serial stage per group
        if axes is None:
            axes = list(range(ndim))
            if darray is not None:
                # the aligned axis of darray must be transformed first (= listed last)
                axes = list(np.roll(axes, ndim - 1 - darray.alignment))
        elif isinstance(axes, (int, np.integer)):
            axes = [axes]
        else:
            axes = list(axes)
        groups = []
        for ax in axes:
            if isinstance(ax, (int, np.integer)):
                grp = [int(ax) % ndim] if -ndim <= ax < ndim else [int(ax)]
            else:
                assert isinstance(ax, (tuple, list))
                grp = []
                for a in ax:
                    assert isinstance(a, (int, np.integer))
                    grp.append(int(a) + ndim if a < 0 else int(a))
            assert min(grp) >= 0
            assert max(grp) < ndim
            assert 0 < len(grp) <= ndim
            assert sorted(grp) == sorted(set(grp))
            groups.append(tuple(grp))
        self.axes = groups
        shape = [int(n) for n in shape]

        if darray is None:
            dtype = np.dtype(dtype)
            assert dtype.char in 'fdgFDG'
            if padding is not False:
                # the physical shape grows along every padded single-axis stage and the
                # factor becomes the exact ratio (reference mpifft.py:247-253)
                padding = list(padding)
                assert len(padding) == len(shape)
                for grp in groups:
                    if len(grp) == 1 and padding[grp[0]] > 1.0 + 1e-6:
                        old = float(shape[grp[0]])
                        shape[grp[0]] = int(np.floor(shape[grp[0]] * padding[grp[0]]))
                        padding[grp[0]] = shape[grp[0]] / old
            self._input_shape = tuple(shape)
            assert len(shape) > 0
            assert min(shape) > 0
            slab = kw.pop('slab', False)

            if grid is not None:
                assert not isinstance(comm, Subcomm)
                assert slab is False
                grid = tuple(grid)
                assert len(grid) <= ndim
                comm = Subcomm(comm, list(grid) + [1] * (ndim - len(grid)))

            if isinstance(comm, Subcomm):
                assert slab is False
                assert len(comm) == ndim
                assert all(comm[ax].Get_size() == 1 for ax in groups[-1])
                self.subcomm = comm
            else:
                if slab is False or slab is None:
                    dims = [0] * ndim
                    for ax in groups[-1]:
                        dims[ax] = 1
                else:  # deprecated slab keyword
                    if slab is True:
                        axis = (groups[-1][-1] + 1) % ndim
                    else:
                        axis = int(slab) % ndim
                    dims = [1] * ndim
                    dims[axis] = comm.Get_size()
                self.subcomm = Subcomm(comm, dims)
        else:
            dtype = darray.dtype
            self.subcomm = darray.subcomm
            self._input_shape = tuple(shape)
            sizes = darray.commsizes
            assert all(sizes[ax] == 1 for ax in groups[-1]), \
                "Set keyword axes such that axes to transform first are aligned"

        self.collapse = collapse
        if collapse is True:
            # merge, from the back, every group whose axes are all undivided into
            # the stage that runs first; divided groups stay separate stages
            merged = [[]]
            for grp in reversed(groups):
                if all(self.subcomm[a].Get_size() == 1 for a in grp):
                    merged[0] = list(grp) + merged[0]
                else:
                    merged.insert(0, list(grp))
            groups = merged

        self.axes = tuple(tuple(g) for g in groups)
        self.xfftn = []
        self.transfer = []
        self.pencil = [None, None]

        # -- first stage: the axes that are undivided on input
        grp = self.axes[-1]
        pencil = Pencil(self.subcomm, shape, grp[-1])
        stage = FFT(pencil.subshape, grp, dtype, padding, backend=backend, transforms=transforms, **kw)
        self.xfftn.append(stage)
        self.pencil[0] = pencilA = pencil
        if shape[grp[-1]] != stage.forward.output_shape[grp[-1]]:
            # real-to-complex: global shape shrinks along that axis, dtype is promoted
            dtype = stage.forward.output_dtype
            shape[grp[-1]] = stage.forward.output_shape[grp[-1]]
            pencilA = Pencil(self.subcomm, shape, grp[-1])

        # -- remaining stages, each behind a redistribution
        for grp in reversed(self.axes[:-1]):
            pencilB = pencilA.pencil(grp[-1])
            self.transfer.append(pencilA.transfer(pencilB, dtype))
            stage = FFT(pencilB.subshape, grp, dtype, padding, backend=backend, transforms=transforms, **kw)
            self.xfftn.append(stage)
            pencilA = pencilB
            if shape[grp[-1]] != stage.forward.output_shape[grp[-1]]:
                dtype = stage.forward.output_dtype
                shape[grp[-1]] = stage.forward.output_shape[grp[-1]]
                pencilA = Pencil(pencilB.subcomm, shape, grp[-1])

        self.pencil[1] = pencilA
        self._output_shape = tuple(shape)

        first, last = self.xfftn[0], self.xfftn[-1]
        self._buffers = _Buffers(ArraySpec(first.forward.input_shape, first.forward.input_dtype),
                                 ArraySpec(last.forward.output_shape, last.forward.output_dtype))
        self.forward = Transform([s.forward for s in self.xfftn],
                                 [t.forward for t in self.transfer],
                                 self.pencil, self._buffers, ('X', 'Y'))
        self.backward = Transform([s.backward for s in self.xfftn[::-1]],
                                  [t.backward for t in self.transfer[::-1]],
                                  self.pencil[::-1], self._buffers, ('Y', 'X'))
        # window set-up is collective: both directions walk the transfers in the same order
        self.forward._all_transfers = self.backward._all_transfers = [t.forward for t in self.transfer]

    def destroy(self):
        if isinstance(self.subcomm, Subcomm):
            self.subcomm.destroy()
        self._buffers.free([t.comm for t in self.transfer] if self._buffers.peers else ())
        for t in self.transfer:
            t.destroy()
        for s in self.xfftn:
            s.destroy()
        for tr in (self.forward, self.backward):
            if tr._merged is not None:
                tr._merged.destroy()

    def shape(self, forward_output=True):
        """Local shape: spectral space if ``forward_output`` else physical."""
        if forward_output is not True:
            return self.forward.input_pencil.subshape
        return tuple(self.xfftn[-1].forward.output_shape)

    def local_slice(self, forward_output=True):
        """Slices of the global array held by this rank."""
        p = self.backward.input_pencil if forward_output is True else self.forward.input_pencil
        return tuple(slice(s, s + n) for s, n in zip(p.substart, p.subshape))

    def global_shape(self, forward_output=False):
        """Global shape in spectral (``forward_output``) or physical space."""
        return self._output_shape if forward_output else self._input_shape

    @property
    def dimensions(self):
        return len(self.xfftn[0].forward.input_shape)

    def dtype(self, forward_output=False):
        """dtype of the spectral (``forward_output``) or physical arrays."""
        if forward_output:
            return self.xfftn[-1].forward.output_dtype
        return self.xfftn[0].forward.input_dtype
