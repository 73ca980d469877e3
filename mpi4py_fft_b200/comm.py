"""Communicator facade: the slice of ``mpi4py.MPI`` that the PFFT path touches.

The reference drives everything through mpi4py communicators
(/root/reference/mpi4py_fft/pencil.py:64-93 builds a Cartesian communicator and
one 1-D sub-communicator per axis; pencil.py:182-201 runs ``Alltoallw`` on
them).  There is no MPI on a B200 box: ranks are one process per GPU started by
``torchrun``; bootstrap and small host-side collectives go through
``torch.distributed`` and the bulk data path goes through NCCL communicators
owned by ``libb200fft.so`` (see transfer in pencil.py of this package).

A :class:`Comm` here is *value* state only -- the ordered tuple of world ranks
that form the group plus (optionally) the Cartesian dims laid over them -- so
``Create_cart`` / ``Sub`` are pure integer arithmetic and can be evaluated for
any "virtual" rank without any process group (that is what makes the index maps
testable bit-exactly on a CPU-only box).  Collective resources (a
``torch.distributed`` group for host collectives, an NCCL communicator for the
device all-to-all) are created lazily, keyed on the group's rank tuple.

Index conventions restated from MPI (and pinned by the reference's doc goldens,
pencil.py:55-62, docs/source/parallel.rst:236-244):
  * ``Compute_dims`` == MPI_Dims_create: balanced factors, non-increasing.
  * Cartesian rank order is row-major (last dim fastest).
"""
from __future__ import annotations

import pickle
import threading
from typing import Optional, Sequence, Tuple

# topology constants (values as in mpi.h of MPICH; only identity matters here)
GRAPH = 1
CART = 2
DIST_GRAPH = 3
UNDEFINED = -32766

SUM = 'sum'
MAX = 'max'
MIN = 'min'
PROD = 'prod'


# --------------------------------------------------------------------------
# MPI_Dims_create
# --------------------------------------------------------------------------
def _balanced_factors(n: int, k: int):
    """Return the k-tuple (d1 >= d2 >= ... >= dk) with product n that is
    lexicographically smallest, i.e. the most balanced factorisation."""
    if k == 0:
        if n != 1:
            raise ValueError("cannot factor %d over zero free dims" % n)
        return ()
    if k == 1:
        return (n,)
    best = None

    def rec(rem, slots, cap, acc):
        nonlocal best
        if slots == 1:
            if rem <= cap:
                cand = acc + (rem,)
                if best is None or cand < best:
                    best = cand
            return
        # largest factor first; factors are non-increasing (<= cap)
        for d in range(min(cap, rem), 0, -1):
            if rem % d:
                continue
            # d must be at least the slots-th root of rem, else the rest cannot
            # stay <= d
            if d ** slots < rem:
                break
            cand_prefix = acc + (d,)
            if best is not None and cand_prefix > best[:len(cand_prefix)]:
                continue
            rec(rem // d, slots - 1, d, cand_prefix)

    rec(n, k, n, ())
    return best


def Compute_dims(nnodes, dims):
    """MPI_Dims_create restated (used at pencil.py:79 of the reference).

    ``dims`` may be an int (number of dimensions, all free) or a sequence in
    which zero entries are free and positive entries are fixed.
    """
    if isinstance(dims, int):
        dims = [0] * dims
    dims = [int(d) for d in dims]
    fixed = 1
    for d in dims:
        if d < 0:
            raise ValueError("negative entry in dims")
        if d > 0:
            fixed *= d
    if nnodes % fixed:
        raise ValueError("nnodes %d is not a multiple of the fixed dims %r" % (nnodes, dims))
    free = [i for i, d in enumerate(dims) if d == 0]
    fac = _balanced_factors(nnodes // fixed, len(free))
    out = list(dims)
    for i, f in zip(free, fac):
        out[i] = f
    return out


def _unravel(r, dims):
    coords = [0] * len(dims)
    for i in range(len(dims) - 1, -1, -1):
        coords[i] = r % dims[i]
        r //= dims[i]
    return tuple(coords)


def _ravel(coords, dims):
    r = 0
    for c, d in zip(coords, dims):
        r = r * d + c
    return r


# --------------------------------------------------------------------------
# world context: who am I, how do I talk to the others
# --------------------------------------------------------------------------
class _World(object):
    """Process-wide state: world size/rank and lazily built collective groups."""

    def __init__(self):
        self._lock = threading.Lock()
        self._pg_cache = {}
        self._pg_ident = None   # id of the default process group the cache belongs to
        self._virtual = None   # (size, rank) override for index-map tests

    # -- identity ----------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        return dist if (dist.is_available() and dist.is_initialized()) else None

    def size(self):
        if self._virtual is not None:
            return self._virtual[0]
        d = self._dist()
        return d.get_world_size() if d else 1

    def rank(self):
        if self._virtual is not None:
            return self._virtual[1]
        d = self._dist()
        return d.get_rank() if d else 0

    # -- host-side collectives over torch.distributed ----------------------
    def group_for(self, ranks: Tuple[int, ...]):
        """torch.distributed group for an ordered tuple of world ranks.

        ``new_group`` is collective over the *world*, so sub-groups are made
        for every sibling partition at once, in a deterministic order; callers
        pass ``siblings`` via :meth:`Comm._siblings`.
        """
        d = self._dist()
        if d is None:
            return None
        if len(ranks) == d.get_world_size():
            return d.group.WORLD
        return self._cache(d).get(tuple(ranks))

    def _cache(self, d):
        """groups belong to ONE initialisation of the default process group: a cache left
        over from a destroyed and re-initialised world would hand out dead groups"""
        ident = id(d.group.WORLD)
        if self._pg_ident != ident:
            self._pg_cache = {}
            self._pg_ident = ident
        return self._pg_cache

    def make_groups(self, partitions):
        """Collectively create one group per partition (all world ranks call
        this with the same list)."""
        d = self._dist()
        if d is None:
            return
        cache = self._cache(d)
        covered = sorted(r for part in partitions for r in part)
        if covered != list(range(d.get_world_size())):
            # torch.distributed.new_group is collective over the WORLD, not over the parent
            # communicator (unlike MPI_Cart_sub): a grid on a strict subset of the ranks would
            # leave the others out of the call and hang.  Fail loudly instead.
            raise NotImplementedError(
                "process grids must span every rank of the torch.distributed world (%d ranks); got partitions "
                "covering ranks %s.  Build the Subcomm / PFFT on COMM_WORLD, or initialise torch.distributed "
                "with the subset as its world." % (d.get_world_size(), covered[:16]))
        for part in partitions:
            part = tuple(part)
            if part in cache or len(part) == d.get_world_size():
                continue
            cache[part] = d.new_group(ranks=list(part))


_world = _World()


class virtual_world(object):
    """Context manager: pretend to be ``rank`` of ``size`` (index maps only).

    Used by the CPU parity tests to evaluate the decomposition of every rank in
    one process; any attempt to communicate inside raises.
    """

    def __init__(self, size, rank):
        self._new = (int(size), int(rank))

    def __enter__(self):
        self._old = _world._virtual
        _world._virtual = self._new
        return self

    def __exit__(self, *exc):
        _world._virtual = self._old
        return False


# --------------------------------------------------------------------------
# Comm
# --------------------------------------------------------------------------
class Comm(object):
    """Group of world ranks, optionally with a Cartesian topology.

    Only the calls made by the hot-path modules of the reference are provided
    (SURVEY.md section 3.5); a few host collectives (bcast/gather/reduce/
    allreduce/barrier) exist for tests and examples.
    """

    __slots__ = ('_ranks', '_me', '_dims', '_freed', '_name')

    def __init__(self, ranks: Sequence[int], me: Optional[int], dims=None, name=None):
        self._ranks = tuple(int(r) for r in ranks)
        self._me = me            # index of this process inside _ranks (None for lazy world/self)
        self._dims = None if dims is None else tuple(int(d) for d in dims)
        self._freed = False
        self._name = name

    # -- lazily resolved identity (COMM_WORLD / COMM_SELF are created at
    #    import time, before torch.distributed is initialised) -------------
    def _resolve(self):
        if self._name == 'world':
            n = _world.size()
            return tuple(range(n)), _world.rank()
        if self._name == 'self':
            r = _world.rank()
            return (r,), 0
        return self._ranks, self._me

    @property
    def ranks(self):
        """Ordered tuple of world ranks in this communicator."""
        return self._resolve()[0]

    def Get_size(self):
        return len(self._resolve()[0])

    def Get_rank(self):
        return self._resolve()[1]

    size = property(Get_size)
    rank = property(Get_rank)

    def Is_inter(self):
        return False

    def Get_topology(self):
        return CART if self._dims is not None else UNDEFINED

    def Get_dim(self):
        if self._dims is None:
            raise ValueError("communicator has no Cartesian topology")
        return len(self._dims)

    @property
    def dims(self):
        return self._dims

    def Get_coords(self, rank=None):
        if rank is None:
            rank = self.Get_rank()
        return list(_unravel(rank, self._dims))

    @property
    def coords(self):
        return self.Get_coords()

    # -- constructors ---------------------------------------------------------
    def Create_cart(self, dims, periods=None, reorder=False):
        """Row-major Cartesian layout over the same ranks (no reordering --
        MPI implementations never reorder for this pattern either)."""
        ranks, me = self._resolve()
        dims = [int(d) for d in dims]
        n = 1
        for d in dims:
            n *= d
        if n != len(ranks):
            raise ValueError("cart dims %r do not cover %d ranks" % (dims, len(ranks)))
        return Comm(ranks, me, dims=dims)

    def Sub(self, remain_dims):
        """MPI_Cart_sub: keep the dims flagged True, fix my coords elsewhere."""
        if self._dims is None:
            raise ValueError("Sub needs a Cartesian communicator")
        ranks, me = self._resolve()
        dims = self._dims
        keep = [bool(k) for k in remain_dims]
        assert len(keep) == len(dims)
        mine = _unravel(me, dims)
        kept_dims = [d for d, k in zip(dims, keep) if k]
        members = []
        n_sub = 1
        for d in kept_dims:
            n_sub *= d
        for idx in range(n_sub):
            sub_coords = _unravel(idx, kept_dims) if kept_dims else ()
            full = list(mine)
            it = iter(sub_coords)
            for i, k in enumerate(keep):
                if k:
                    full[i] = next(it)
            members.append(ranks[_ravel(full, dims)])
        my_sub = _ravel([c for c, k in zip(mine, keep) if k], kept_dims) if kept_dims else 0
        return Comm(members, my_sub, dims=kept_dims)

    def partitions(self):
        """For a Cart-sub communicator made by :meth:`Sub`: not recoverable in
        general, so collective group creation is keyed on the *parent* (see
        Subcomm in pencil.py which records sibling partitions)."""
        return [self.ranks]

    def Dup(self):
        ranks, me = self._resolve()
        return Comm(ranks, me, dims=self._dims)

    def Free(self):
        self._freed = True

    # -- comparisons ----------------------------------------------------------
    def __eq__(self, other):
        if not isinstance(other, Comm):
            return NotImplemented
        return self.ranks == other.ranks and self._dims == other._dims \
            if (self._dims is not None and other._dims is not None) \
            else self.ranks == other.ranks

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash(self.ranks)

    def __bool__(self):
        return not self._freed

    def __repr__(self):
        return "Comm(ranks=%r, rank=%r, dims=%r)" % (self.ranks, self.Get_rank(), self._dims)

    # -- host collectives (tests/examples; python objects, not bulk data) ----
    def _group(self):
        import torch.distributed as dist
        ranks = self.ranks
        if len(ranks) == 1:
            return None, None
        if _world._virtual is not None or not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("communication requested but torch.distributed is not initialised "
                               "(world size %d)" % len(ranks))
        g = _world.group_for(ranks)
        if g is None:
            raise RuntimeError("no process group for ranks %r; sub-communicators must be created "
                               "through Subcomm (collective)" % (ranks,))
        return dist, g

    def Barrier(self):
        dist, g = self._group()
        if dist is not None:
            dist.barrier(group=g)

    barrier = Barrier

    def allgather(self, obj):
        dist, g = self._group()
        if dist is None:
            return [obj]
        out = [None] * self.Get_size()
        dist.all_gather_object(out, obj, group=g)
        return out

    def gather(self, obj, root=0):
        out = self.allgather(obj)
        return out if self.Get_rank() == root else None

    def bcast(self, obj, root=0):
        dist, g = self._group()
        if dist is None:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=self.ranks[root], group=g)
        return box[0]

    @staticmethod
    def _combine(vals, op):
        import functools
        import operator
        if op == SUM:
            return functools.reduce(operator.add, vals)
        if op == PROD:
            return functools.reduce(operator.mul, vals)
        if op == MAX:
            return max(vals)
        if op == MIN:
            return min(vals)
        raise ValueError("unknown reduction %r" % (op,))

    def allreduce(self, obj, op=SUM):
        return self._combine(self.allgather(obj), op)

    def reduce(self, obj, op=SUM, root=0):
        out = self.allreduce(obj, op)
        return out if self.Get_rank() == root else None


COMM_WORLD = Comm((), None, name='world')
COMM_SELF = Comm((), None, name='self')
COMM_NULL = None


def ensure_groups(partitions):
    """Collectively (over the world) create host groups for ``partitions``."""
    _world.make_groups(partitions)


def world():
    return _world
