"""Worker for the multi-GPU parity tests: launched by torchrun, one rank per GPU,
NCCL all-to-all inside Transfer.  Runs every golden case whose rank count equals
the world size and compares each rank's block with the reference's fixture."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def main():
    import torch
    import mpi4py_fft_b200 as B
    from conftest import case_kwargs
    import pfft_oracle as O
    comm = B.init()
    world, rank = comm.Get_size(), comm.Get_rank()
    layouts = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'layouts.json')))
    values = np.load(os.path.join(ROOT, 'tests', 'golden', 'values.npz'))
    ran = 0
    only = os.environ.get('MP_ONLY')
    for name, case in sorted(layouts.items()):
        if name.startswith('_') or case['meta']['nranks'] != world or name + '__input' not in values:
            continue
        if only and name != only:
            continue
        kw = case_kwargs(case['meta'])
        g = values[name + '__input']
        tol = 1e-5 if g.dtype.char in 'fF' else 1e-12
        fft = B.PFFT(comm, **kw)
        ref_rank = case['ranks'][rank]
        assert [c.Get_size() for c in fft.subcomm] == ref_rank['subcomm_sizes']
        assert [[s.start, s.stop] for s in fft.local_slice(True)] == ref_rank['local_slice_out']
        u = B.newDistArray(fft, False)
        u[...] = np.ascontiguousarray(g[fft.local_slice(False)])
        uh = fft.forward(u)
        ref = values[name + '__forward']
        err = np.abs(np.asarray(uh) - ref[fft.local_slice(True)]).max()
        assert err <= tol * max(1.0, np.abs(ref).max()), (name, rank, err)
        ub = B.newDistArray(fft, False)
        fft.backward(uh, ub)
        refb = values[name + '__backward']        # == g except for padded (lossy) transforms
        err = np.abs(np.asarray(ub) - refb[fft.local_slice(False)]).max()
        assert err <= 10 * tol, (name, rank, err)
        # DistArray.redistribute == a bare Transfer (reference distarray.py:298-363)
        z = B.DistArray(g.shape, dtype=g.dtype, alignment=len(g.shape) - 1)
        z[...] = np.ascontiguousarray(g[z.local_slice()])
        z0 = z.redistribute(0)
        assert np.array_equal(np.asarray(z0), g[z0.local_slice()]), name
        # same elements, other partition: sum in float64 so that only the order of additions differs
        s0 = comm.allreduce(float((np.abs(np.asarray(z).astype(np.result_type(g.dtype, np.float64))) ** 2).sum()))
        s1 = comm.allreduce(float((np.abs(np.asarray(z0).astype(np.result_type(g.dtype, np.float64))) ** 2).sum()))
        assert abs(s0 - s1) <= 1e-9 * s0
        fft.destroy()
        ran += 1
        if rank == 0:
            print('ok', name, flush=True)
    # a bigger power-of-two case against the oracle (every rank computes the global reference)
    shape = (64, 64, 64)
    g = np.random.default_rng(3).random(shape) + 1j * np.random.default_rng(4).random(shape)
    fft = B.PFFT(comm, shape, dtype='D')
    u = B.newDistArray(fft, False)
    u[...] = np.ascontiguousarray(g[fft.local_slice(False)])
    uh = fft.forward(u)
    ref = O.expected_forward(g)
    err = np.abs(np.asarray(uh) - ref[fft.local_slice(True)]).max()
    assert err < 1e-12, err
    ub = fft.backward(uh)
    assert np.abs(np.asarray(ub) - g[fft.local_slice(False)]).max() < 1e-12
    torch.cuda.synchronize()
    comm.Barrier()
    modes = sorted(set('p2p' if v is not None else 'nccl' for v in fft._buffers.peers.values())) or ['nccl']
    if rank == 0:
        print('MULTI_OK cases=%d world=%d transfers=%s' % (ran, world, '+'.join(modes)), flush=True)
    fft.destroy()
    torch.cuda.synchronize()
    comm.Barrier()
    import torch.distributed as dist
    dist.destroy_process_group()


if __name__ == '__main__':
    try:
        main()
    except BaseException:
        import traceback
        sys.stdout.write('RANK %s FAILED\n%s\n' % (os.environ.get('RANK'), traceback.format_exc()))
        sys.stdout.flush()
        raise
