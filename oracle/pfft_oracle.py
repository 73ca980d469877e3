"""CPU restatement of mpi4py-fft's PFFT.forward/backward path -- TEST INFRASTRUCTURE.

ORACLE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference arm may import this module.  The product path
(mpi4py_fft_b200/) never does and has no CPU fallback.

What is restated (numpy, all ranks of the job simulated in one process):
  blockdist         /root/reference/mpi4py_fft/pencil.py:5-9
  compute_dims      MPI_Dims_create as used at pencil.py:79 (balanced,
                    non-increasing; pinned by the doc goldens pencil.py:55-62)
  cart layout       row-major ranks, MPI_Cart_sub per axis, pencil.py:80-88
  Pencil            pencil.py:277-323 (subshape/substart, axis swap)
  exchange          pencil.py:12-29,182-183 (Alltoallw over subarray types ==
                    block i of axisA of sender -> block of axisB of receiver)
  plan              mpifft.py:213-337 (axes groups, grid, collapse, r2c shape
                    and dtype propagation)
  stage             libfft.py:408-422 (forward *= M, backward unnormalised)
  padding           mpifft.py:247-253 (padded physical shape), libfft.py:263-311
                    (spectral truncation / zero padding with the symmetric
                    Nyquist rule), libfft.py:424-434 (truncated extents)
  serial transforms the FFTW definitions (third party, un-vendored, no version
                    pinned by the reference: setup.py:64-81 probes library
                    names only) evaluated with numpy/scipy pocketfft: c2c
                    sign -1/+1 unnormalised, r2c/c2r half spectrum on the
                    last axis of the group, r2r kinds per xfftn.py:14-36.

Pinned against: the docstring vectors of fftw/xfftn.py, the layout goldens of
pencil.py/distarray.py/docs, and the unmodified reference run under
oracle/fakempi (fixtures in tests/golden/, made by oracle/make_golden.py).
"""
import itertools

import numpy as np
import scipy.fft as sfft

# FFTW kind integers (reference fftw/utilities.pyx:7-26)
FORWARD, BACKWARD, R2C, C2R = -1, 1, -2, 2
REDFT00, REDFT01, REDFT10, REDFT11, RODFT00, RODFT01, RODFT10, RODFT11 = 3, 4, 5, 6, 7, 8, 9, 10
_R2R_SCIPY = {REDFT00: ('dct', 1), REDFT10: ('dct', 2), REDFT01: ('dct', 3), REDFT11: ('dct', 4),
              RODFT00: ('dst', 1), RODFT10: ('dst', 2), RODFT01: ('dst', 3), RODFT11: ('dst', 4)}
# forward type -> (forward kind, backward kind), reference xfftn.py:14-36
DCT = {1: (REDFT00, REDFT00), 2: (REDFT10, REDFT01), 3: (REDFT01, REDFT10), 4: (REDFT11, REDFT11)}
DST = {1: (RODFT00, RODFT00), 2: (RODFT10, RODFT01), 3: (RODFT01, RODFT10), 4: (RODFT11, RODFT11)}


def blockdist(N, size, rank):
    q, r = divmod(N, size)
    n = q + (1 if r > rank else 0)
    s = rank * q + min(rank, r)
    return n, s


def compute_dims(nnodes, dims):
    dims = list(dims)
    fixed = 1
    for d in dims:
        if d > 0:
            fixed *= d
    assert nnodes % fixed == 0
    rem = nnodes // fixed
    free = [i for i, d in enumerate(dims) if d == 0]

    def factorisations(n, k, cap):
        if k == 0:
            if n == 1:
                yield ()
            return
        for d in range(min(cap, n), 0, -1):
            if n % d == 0:
                for rest in factorisations(n // d, k - 1, d):
                    yield (d,) + rest

    best = min(factorisations(rem, len(free), rem)) if free else ()
    for i, f in zip(free, best):
        dims[i] = f
    return dims


def subcomm_dims(nranks, dims):
    """process-grid sizes of Subcomm(comm, dims) (pencil.py:70-79)"""
    if dims is None:
        dims = [0]
    elif np.ndim(dims) > 0:
        dims = [max(0, d) for d in dims]
    else:
        dims = [0] * dims
    return compute_dims(nranks, dims)


class VPencil(object):
    """Pencil of one virtual rank: per-axis (group size, group rank)."""

    def __init__(self, sizes, ranks, shape, axis):
        self.sizes, self.ranks = tuple(sizes), tuple(ranks)
        self.shape, self.axis = tuple(shape), axis
        assert self.sizes[axis] == 1
        ns = [blockdist(n, p, r) for n, p, r in zip(shape, sizes, ranks)]
        for n, p in zip(shape, sizes):
            assert n >= p
        self.subshape = tuple(n for n, _ in ns)
        self.substart = tuple(s for _, s in ns)

    def pencil(self, axis):
        sizes, ranks = list(self.sizes), list(self.ranks)
        i, j = self.axis, axis
        sizes[i], sizes[j] = sizes[j], sizes[i]
        ranks[i], ranks[j] = ranks[j], ranks[i]
        return VPencil(sizes, ranks, self.shape, axis)

    def slices(self):
        return tuple(slice(s, s + n) for s, n in zip(self.substart, self.subshape))


def serial_transform(a, axes, kinds):
    """unnormalised FFTW-convention transform of ``a`` over ``axes``"""
    k0 = kinds[0]
    if k0 == FORWARD:
        return sfft.fftn(a, axes=axes)
    if k0 == BACKWARD:
        return sfft.ifftn(a, axes=axes, norm='forward')
    if k0 == R2C:
        return sfft.rfftn(a, axes=axes)
    if k0 == C2R:
        raise ValueError("c2r needs the real lengths: use serial_c2r")
    out = a
    for ax, k in zip(axes, kinds):
        name, typ = _R2R_SCIPY[k]
        out = getattr(sfft, name)(out, type=typ, axis=ax)
    return out


def serial_c2r(a, axes, s):
    return sfft.irfftn(a, s=s, axes=axes, norm='forward')


def truncate_forward(padded, axis, n_trunc, real):
    """libfft.py:263-286: keep n_trunc modes of a padded spectrum along ``axis``
    (n_trunc = extent of the truncated array: N for complex, N//2+1 for r2c)"""
    shape = list(padded.shape)
    shape[axis] = n_trunc
    trunc = np.zeros(shape, dtype=padded.dtype)
    N = n_trunc
    if real:
        s = [slice(None)] * trunc.ndim
        s[axis] = slice(0, N)
        trunc[:] = padded[tuple(s)]
        if N % 2 == 0:          # the reference tests the truncated array's own extent (libfft.py:267,274)
            s[axis] = N - 1
            s = tuple(s)
            trunc[s] = trunc[s].real
            trunc[s] *= 2
    else:
        su = [slice(None)] * trunc.ndim
        su[axis] = slice(0, N // 2 + 1)
        trunc[tuple(su)] = padded[tuple(su)]
        su[axis] = slice(-(N // 2), None)
        trunc[tuple(su)] += padded[tuple(su)]
    return trunc


def pad_backward(trunc, axis, n_padded, real):
    """libfft.py:288-311: zero-pad a truncated spectrum to n_padded along ``axis``"""
    shape = list(trunc.shape)
    shape[axis] = n_padded
    padded = np.zeros(shape, dtype=trunc.dtype)
    N = trunc.shape[axis]
    if real:
        s = [slice(0, n) for n in trunc.shape]
        padded[tuple(s)] = trunc[:]
        if N % 2 == 0:
            s[axis] = N - 1
            s = tuple(s)
            padded[s] = padded[s].real
            padded[s] *= 0.5
    else:
        su = [slice(None)] * trunc.ndim
        su[axis] = slice(0, N // 2 + 1)
        padded[tuple(su)] = trunc[tuple(su)]
        su[axis] = slice(-(N // 2), None)
        padded[tuple(su)] = trunc[tuple(su)]
        if N % 2 == 0:
            su[axis] = N // 2
            padded[tuple(su)] *= 0.5
            su[axis] = -(N // 2)
            padded[tuple(su)] *= 0.5
    return padded


def padded_stage_forward(x, axis, pf, normalize=True):
    """one serial padded stage, libfft.py:376-415 with 424-434: x has the padded
    extent along ``axis``"""
    real = not np.iscomplexobj(x)
    Np = x.shape[axis]
    V = sfft.rfft(x, axis=axis) if real else sfft.fft(x, axis=axis)
    N = int(np.round(Np / pf))
    out = truncate_forward(V, axis, N // 2 + 1 if real else N, real)
    return out / Np if normalize else out


def padded_stage_backward(y, axis, n_padded, real, normalize=False):
    V = pad_backward(y, axis, n_padded // 2 + 1 if real else n_padded, real)
    x = sfft.irfft(V, n=n_padded, axis=axis, norm='forward') if real else sfft.ifft(V, axis=axis, norm='forward')
    return x / n_padded if normalize else x


class OraclePFFT(object):
    """All ranks of a PFFT at once.

    ``transforms`` maps an axes tuple to ('dct'|'dst', type) -- the counterpart of
    the reference's {axes: (dctn, idctn)} dictionaries; default is rfftn/irfftn
    for real dtypes and fftn/ifftn for complex ones (libfft.py:60-70).
    """

    def __init__(self, nranks, shape, axes=None, dtype=float, grid=None, collapse=False,
                 transforms=None, subcomm_dims_arg=None, padding=False):
        self.nranks = nranks
        ndim = len(shape)
        # ---- axes normalisation (mpifft.py:213-240)
        if axes is not None:
            axes = list(axes) if not isinstance(axes, int) else [axes]
        else:
            axes = list(range(ndim))
        for i, ax in enumerate(axes):
            if isinstance(ax, (int, np.integer)):
                axes[i] = (ax + ndim if ax < 0 else ax,)
            else:
                axes[i] = tuple(a + ndim if a < 0 else a for a in ax)
        shape = list(shape)
        dtype = np.dtype(dtype)
        # ---- padding: the physical shape grows, the factor becomes the exact ratio (mpifft.py:247-253)
        if padding is not False:
            padding = list(padding)
            assert len(padding) == len(shape)
            for ax in axes:
                if len(ax) == 1 and padding[ax[0]] > 1.0 + 1e-6:
                    old = float(shape[ax[0]])
                    shape[ax[0]] = int(np.floor(shape[ax[0]] * padding[ax[0]]))
                    padding[ax[0]] = shape[ax[0]] / old
        self.padding = padding
        self.input_shape = tuple(shape)
        # ---- process grid (mpifft.py:259-290)
        if grid is not None:
            dims = list(grid) + [1] * (ndim - len(grid))
            dims = subcomm_dims(nranks, dims)
        elif subcomm_dims_arg is not None:
            dims = subcomm_dims(nranks, subcomm_dims_arg)
        else:
            dims = [0] * ndim
            for ax in axes[-1]:
                dims[ax] = 1
            dims = subcomm_dims(nranks, dims)
        self.dims = dims
        self.coords = [np.unravel_index(r, dims) for r in range(nranks)]   # row-major cart
        for ax in axes[-1]:
            assert dims[ax] == 1
        # ---- collapse (mpifft.py:298-306)
        if collapse is True:
            groups = [[]]
            for ax in reversed(axes):
                if all(dims[a] == 1 for a in ax):
                    groups[0] = list(ax) + groups[0]
                else:
                    groups.insert(0, list(ax))
            axes = groups
        self.axes = tuple(map(tuple, axes))
        transforms = {} if transforms is None else {tuple(k): v for k, v in transforms.items()}

        # ---- stage chain (mpifft.py:313-337); everything per rank
        self.stages = []       # dict(axes, kinds_f, kinds_b, M, in_shape, out_shape (global), in_dtype, out_dtype, pencils_in, pencils_out)
        self.transfers = []    # dict(axisA, axisB, pencilsA, pencilsB)

        def make_stage(grp, gshape, dt, pencils):
            real = np.issubdtype(dt, np.floating)
            oshape = list(gshape)
            odt = dt
            pf = 1.0 if self.padding is False else self.padding[grp[-1]]
            if abs(pf - 1.0) > 1e-8:
                # padded stage (libfft.py:401-406,424-434): one axis, spectrum truncated
                assert len(grp) == 1 and grp not in transforms
                N = int(np.round(gshape[grp[0]] / pf))
                oshape[grp[0]] = N // 2 + 1 if real else N
                odt = np.dtype(dt.char.upper())
                return dict(axes=grp, kinds_f=[R2C if real else FORWARD], kinds_b=[C2R if real else BACKWARD],
                            M=1.0 / gshape[grp[0]], in_gshape=tuple(gshape), out_gshape=tuple(oshape),
                            in_dtype=dt, out_dtype=odt, pencils_in=pencils, pad=pf)
            if grp in transforms:
                fam, typ = transforms[grp]
                kf, kb = (DCT if fam == 'dct' else DST)[typ]
                kinds_f, kinds_b = [kf] * len(grp), [kb] * len(grp)
                M = 1.0
                for a in grp:     # xfftn.py:763-816
                    n = gshape[a]
                    M *= 2 * (n + 1) if kf == RODFT00 else 2 * (n - 1) if kf == REDFT00 else 2 * n
                M = 1.0 / M
            elif real:
                kinds_f, kinds_b = [R2C], [C2R]
                oshape[grp[-1]] = gshape[grp[-1]] // 2 + 1
                odt = np.dtype(dt.char.upper())
                M = 1.0 / np.prod([gshape[a] for a in grp])
            else:
                kinds_f, kinds_b = [FORWARD], [BACKWARD]
                M = 1.0 / np.prod([gshape[a] for a in grp])
            return dict(axes=grp, kinds_f=kinds_f, kinds_b=kinds_b, M=M, in_gshape=tuple(gshape),
                        out_gshape=tuple(oshape), in_dtype=dt, out_dtype=odt, pencils_in=pencils)

        grp = self.axes[-1]
        pencils = [VPencil(dims, self.coords[r], shape, grp[-1]) for r in range(nranks)]
        self.input_pencils = pencils
        st = make_stage(grp, shape, dtype, pencils)
        self.stages.append(st)
        if st['out_gshape'] != tuple(shape):
            shape = list(st['out_gshape'])
            dtype = st['out_dtype']
            pencils = [VPencil(dims, self.coords[r], shape, grp[-1]) for r in range(nranks)]
        st['pencils_out'] = pencils
        pencilsA = pencils
        for grp in reversed(self.axes[:-1]):
            pencilsB = [p.pencil(grp[-1]) for p in pencilsA]
            self.transfers.append(dict(axisA=pencilsA[0].axis, axisB=pencilsB[0].axis,
                                       pencilsA=pencilsA, pencilsB=pencilsB, dtype=dtype))
            st = make_stage(grp, shape, dtype, pencilsB)
            self.stages.append(st)
            pencilsA = pencilsB
            if st['out_gshape'] != tuple(shape):
                shape = list(st['out_gshape'])
                dtype = st['out_dtype']
                pencilsA = [VPencil(p.sizes, p.ranks, shape, grp[-1]) for p in pencilsB]
            st['pencils_out'] = pencilsA
        self.output_pencils = pencilsA
        self.output_shape = tuple(shape)
        self.output_dtype = dtype
        self.input_dtype = self.stages[0]['in_dtype']

    # ---- introspection used by the parity tests ------------------------------------
    def local_slice(self, rank, forward_output=True):
        p = self.output_pencils[rank] if forward_output else self.input_pencils[rank]
        return p.slices()

    def layout(self):
        """JSON-able description of every rank's view: the bit-exact contract"""
        out = dict(dims=[int(d) for d in self.dims], axes=[list(a) for a in self.axes],
                   input_shape=list(self.input_shape), output_shape=list(self.output_shape),
                   output_dtype=self.output_dtype.char, ranks=[])
        for r in range(self.nranks):
            stages = []
            for st in self.stages:
                pi, po = st['pencils_in'][r], st['pencils_out'][r]
                stages.append(dict(axes=list(st['axes']), in_subshape=list(pi.subshape),
                                   in_substart=list(pi.substart), in_axis=pi.axis,
                                   out_subshape=list(po.subshape), out_substart=list(po.substart)))
            transfers = []
            for tr in self.transfers:
                pa, pb = tr['pencilsA'][r], tr['pencilsB'][r]
                p = pa.sizes[pb.axis]
                NA, NB = pa.shape[pa.axis], pa.shape[pb.axis]
                transfers.append(dict(axisA=pa.axis, axisB=pb.axis, group_size=int(p),
                                      group_rank=int(pa.ranks[pb.axis]),
                                      subshapeA=list(pa.subshape), subshapeB=list(pb.subshape),
                                      blocksA=[list(blockdist(NA, p, i)) for i in range(p)],
                                      blocksB=[list(blockdist(NB, p, i)) for i in range(p)]))
            out['ranks'].append(dict(coords=[int(c) for c in self.coords[r]], stages=stages,
                                     transfers=transfers))
        return out

    # ---- the exchange (pencil.py:12-29,182-183) ----------------------------------------
    def _exchange(self, tr, blocks, backward=False):
        pencilsA, pencilsB = tr['pencilsA'], tr['pencilsB']
        if backward:
            pencilsA, pencilsB = pencilsB, pencilsA
        axisA, axisB = pencilsA[0].axis, pencilsB[0].axis
        NA, NB = pencilsA[0].shape[axisA], pencilsA[0].shape[axisB]
        out = [np.zeros(pencilsB[r].subshape, dtype=blocks[r].dtype) for r in range(self.nranks)]
        for r in range(self.nranks):
            pa = pencilsA[r]
            p = pa.sizes[axisB]                 # size of the group doing this exchange
            # the members of r's group differ from r only in the rank along axisB
            for peer_rank in range(p):
                peer = self._find(pencilsA, pa, axisB, peer_rank)
                nA, sA = blockdist(NA, p, peer_rank)          # block of axisA I send to peer
                nB, sB = blockdist(NB, p, pa.ranks[axisB])    # where my data lands along axisB in peer's B
                src = [slice(None)] * len(pa.shape)
                src[axisA] = slice(sA, sA + nA)
                dst = [slice(None)] * len(pa.shape)
                dst[axisB] = slice(sB, sB + nB)
                out[peer][tuple(dst)] = blocks[r][tuple(src)]
        return out

    def _find(self, pencils, mine, axis, rank_along_axis):
        key = list(mine.ranks)
        key[axis] = rank_along_axis
        key = tuple(key)
        if not hasattr(self, '_index'):
            self._index = {}
        ident = id(pencils)
        if ident not in self._index:
            self._index[ident] = {tuple(p.ranks): r for r, p in enumerate(pencils)}
        return self._index[ident][key]

    # ---- the transforms -------------------------------------------------------------------
    def scatter(self, g, forward_output=False):
        return [np.ascontiguousarray(g[self.local_slice(r, forward_output)]) for r in range(self.nranks)]

    def gather(self, blocks, forward_output=True):
        shape = self.output_shape if forward_output else self.input_shape
        g = np.zeros(shape, dtype=blocks[0].dtype)
        for r in range(self.nranks):
            g[self.local_slice(r, forward_output)] = blocks[r]
        return g

    def forward(self, blocks, normalize=True):
        cur = [np.asarray(b) for b in blocks]
        for i, st in enumerate(self.stages):
            nxt = []
            for b in cur:
                if 'pad' in st:
                    v = padded_stage_forward(b.astype(st['in_dtype'], copy=False), st['axes'][0], st['pad'], normalize)
                    nxt.append(v.astype(st['out_dtype'], copy=False))
                    continue
                v = serial_transform(b.astype(st['in_dtype'], copy=False), st['axes'], st['kinds_f'])
                if normalize:
                    v = v * st['M']
                nxt.append(v.astype(st['out_dtype'], copy=False))
            cur = nxt
            if i < len(self.transfers):
                cur = self._exchange(self.transfers[i], cur)
        return cur

    def backward(self, blocks, normalize=False):
        cur = [np.asarray(b) for b in blocks]
        n = len(self.stages)
        for i in range(n - 1, -1, -1):
            st = self.stages[i]
            nxt = []
            for r, b in enumerate(cur):
                if 'pad' in st:
                    ax = st['axes'][0]
                    v = padded_stage_backward(b, ax, st['pencils_in'][r].subshape[ax],
                                              np.issubdtype(st['in_dtype'], np.floating), normalize)
                    nxt.append(v.astype(st['in_dtype'], copy=False))
                    continue
                if st['kinds_b'][0] == C2R:
                    s = [st['pencils_in'][r].subshape[a] for a in st['axes']]
                    v = serial_c2r(b, st['axes'], s)
                else:
                    v = serial_transform(b, st['axes'], st['kinds_b'])
                if normalize:
                    v = v * st['M']
                nxt.append(v.astype(st['in_dtype'], copy=False))
            cur = nxt
            if i > 0:
                cur = self._exchange(self.transfers[i - 1], cur, backward=True)
        return cur


def expected_forward(g, axes=None, transforms=None, dtype=None):
    """The distributed forward result computed on the undistributed array:
    normalised transform of the global array (what every decomposition must
    agree with; stronger than the reference's round-trip tests)."""
    a = np.asarray(g)
    ndim = a.ndim
    if axes is None:
        axes = list(range(ndim))
    axes = [axes] if isinstance(axes, int) else list(axes)
    groups = [(ax % ndim,) if isinstance(ax, (int, np.integer)) else tuple(x % ndim for x in ax) for ax in axes]
    transforms = {} if transforms is None else {tuple(k): v for k, v in transforms.items()}
    out = a
    for grp in reversed(groups):
        if grp in transforms:
            fam, typ = transforms[grp]
            kf = (DCT if fam == 'dct' else DST)[typ][0]
            M = 1.0
            for ax in grp:
                n = out.shape[ax]
                M *= 2 * (n + 1) if kf == RODFT00 else 2 * (n - 1) if kf == REDFT00 else 2 * n
            out = serial_transform(out, grp, [kf] * len(grp)) / M
        elif not np.iscomplexobj(out):      # real data meets a default stage: r2c (libfft.py:64-66)
            M = np.prod([out.shape[ax] for ax in grp])
            out = sfft.rfftn(out, axes=grp) / M
        else:
            M = np.prod([out.shape[ax] for ax in grp])
            out = sfft.fftn(out, axes=grp) / M
    return out
