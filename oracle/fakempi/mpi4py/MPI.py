"""Thread-per-rank stand-in for ``mpi4py.MPI`` -- TEST INFRASTRUCTURE ONLY.

mpi4py and an MPI library are not installable in this image, so the unmodified
reference (/root/reference/mpi4py_fft, imported from oracle/_ref) is run with
this module on ``sys.path`` as ``mpi4py.MPI``.  It implements exactly the calls
the reference's hot-path modules make: every "rank" is
a thread of one process (``run_ranks``), ``COMM_WORLD`` resolves the calling
thread's rank, ``Create_cart``/``Sub`` are row-major coordinate arithmetic,
``Compute_dims`` restates MPI_Dims_create (balanced, non-increasing), and
``Alltoallw`` is a rendezvous followed by numpy slice copies described by the
subarray datatypes.  Only tests/, bench.py's reference arm and
oracle/make_golden.py import it; the product never does.
"""
import threading

import numpy as np

CART = 2
UNDEFINED = -32766
SUM, MAX, MIN, PROD = 'sum', 'max', 'min', 'prod'

_tls = threading.local()
_registry_lock = threading.Lock()
_groups = {}


def _world_size():
    return getattr(_tls, 'size', 1)


def _world_rank():
    return getattr(_tls, 'rank', 0)


class _GroupState(object):
    """rendezvous state shared by the threads of one group of ranks"""

    def __init__(self, n):
        self.barrier = threading.Barrier(n)
        self.board = [None] * n


def _state(ranks):
    key = (getattr(_tls, 'job', 0), tuple(ranks))
    with _registry_lock:
        st = _groups.get(key)
        if st is None:
            st = _groups[key] = _GroupState(len(ranks))
        return st


def Compute_dims(nnodes, dims):
    """MPI_Dims_create: fill the zero entries of ``dims`` with a balanced,
    non-increasing factorisation of nnodes / prod(nonzero entries)."""
    if isinstance(dims, int):
        dims = [0] * dims
    dims = list(dims)
    fixed = int(np.prod([d for d in dims if d > 0])) if any(d > 0 for d in dims) else 1
    assert nnodes % fixed == 0
    rem = nnodes // fixed
    free = [i for i, d in enumerate(dims) if d == 0]
    k = len(free)
    best = [None]

    def rec(r, slots, cap, acc):
        if slots == 0:
            if r == 1 and (best[0] is None or acc < best[0]):
                best[0] = acc
            return
        for d in range(min(cap, r), 0, -1):
            if r % d == 0:
                rec(r // d, slots - 1, d, acc + (d,))

    rec(rem, k, rem, ())
    for i, f in zip(free, best[0] or ()):
        dims[i] = f
    return dims


class Datatype(object):
    def __init__(self, char):
        self.char = char
        self.sub = None

    def Create_subarray(self, sizes, subsizes, starts, order=None):
        t = Datatype(self.char)
        t.sub = (tuple(sizes), tuple(subsizes), tuple(starts))
        return t

    def Commit(self):
        return self

    def Free(self):
        self.sub = None

    def __bool__(self):
        return self.sub is not None or True

    def slices(self):
        sizes, subsizes, starts = self.sub
        return tuple(slice(s, s + n) for s, n in zip(starts, subsizes))


_typedict = {c: Datatype(c) for c in 'fdgFDGilqbBhHIQL?'}


class Comm(object):
    def __init__(self, ranks=None, dims=None, name=None):
        self._ranks = None if ranks is None else tuple(ranks)
        self._dims = None if dims is None else tuple(dims)
        self._name = name
        self._freed = False

    # identity
    def _members(self):
        if self._name == 'world':
            return tuple(range(_world_size()))
        if self._name == 'self':
            return (_world_rank(),)
        return self._ranks

    def Get_size(self):
        return len(self._members())

    def Get_rank(self):
        return self._members().index(_world_rank())

    size = property(Get_size)
    rank = property(Get_rank)

    def Is_inter(self):
        return False

    def Get_topology(self):
        return CART if self._dims is not None else UNDEFINED

    def Get_dim(self):
        return len(self._dims)

    def Create_cart(self, dims, periods=None, reorder=False):
        dims = list(dims)
        assert int(np.prod(dims)) == self.Get_size()
        return Comm(self._members(), dims)

    def Sub(self, remain_dims):
        members, dims = self._members(), self._dims
        me = np.unravel_index(members.index(_world_rank()), dims) if dims else ()
        keep = [bool(k) for k in remain_dims]
        kept = [d for d, k in zip(dims, keep) if k]
        out = []
        for idx in range(int(np.prod(kept)) if kept else 1):
            sub = list(np.unravel_index(idx, kept)) if kept else []
            full = list(me)
            it = iter(sub)
            for i, k in enumerate(keep):
                if k:
                    full[i] = next(it)
            out.append(members[int(np.ravel_multi_index(full, dims))])
        return Comm(out, kept)

    def Dup(self):
        return Comm(self._members(), self._dims)

    def Free(self):
        self._freed = True

    def __eq__(self, other):
        return isinstance(other, Comm) and self._members() == other._members()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._members())

    def __bool__(self):
        return not self._freed

    # collectives
    def Barrier(self):
        if self.Get_size() > 1:
            _state(self._members()).barrier.wait()

    barrier = Barrier

    def allgather(self, obj):
        n = self.Get_size()
        if n == 1:
            return [obj]
        st = _state(self._members())
        st.board[self.Get_rank()] = obj
        st.barrier.wait()
        out = list(st.board)
        st.barrier.wait()
        return out

    def gather(self, obj, root=0):
        out = self.allgather(obj)
        return out if self.Get_rank() == root else None

    def bcast(self, obj, root=0):
        return self.allgather(obj)[root]

    def allreduce(self, obj, op=SUM):
        vals = self.allgather(obj)
        if op == SUM:
            r = vals[0]
            for v in vals[1:]:
                r = r + v
            return r
        if op == MAX:
            return max(vals)
        if op == MIN:
            return min(vals)
        raise ValueError(op)

    def reduce(self, obj, op=SUM, root=0):
        r = self.allreduce(obj, op)
        return r if self.Get_rank() == root else None

    def Alltoallw(self, sendspec, recvspec):
        sendbuf, _, sendtypes = sendspec
        recvbuf, _, recvtypes = recvspec
        n = self.Get_size()
        me = self.Get_rank()
        if n == 1:
            recvbuf[recvtypes[0].slices()] = sendbuf[sendtypes[0].slices()]
            return
        st = _state(self._members())
        st.board[me] = (sendbuf, sendtypes)
        st.barrier.wait()
        for j in range(n):
            sbuf, stypes = st.board[j]
            recvbuf[recvtypes[j].slices()] = sbuf[stypes[me].slices()]
        st.barrier.wait()


COMM_WORLD = Comm(name='world')
COMM_SELF = Comm(name='self')
COMM_NULL = None

_job_counter = [0]


def run_ranks(nranks, fn, *args, **kw):
    """Run ``fn(*args, **kw)`` on ``nranks`` threads, each seeing its own rank
    through COMM_WORLD; returns the list of results (exceptions re-raised)."""
    results = [None] * nranks
    errors = [None] * nranks
    with _registry_lock:
        _job_counter[0] += 1
        job = _job_counter[0]

    def body(r):
        _tls.size, _tls.rank, _tls.job = nranks, r, job
        try:
            results[r] = fn(*args, **kw)
        except BaseException as e:  # noqa
            errors[r] = e
            # release peers stuck in a barrier
            with _registry_lock:
                for (j, _), st in list(_groups.items()):
                    if j == job:
                        st.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    with _registry_lock:
        for key in [k for k in _groups if k[0] == job]:
            del _groups[key]
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results
