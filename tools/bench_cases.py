"""Forward+backward timing of other BASELINE.json configurations on the GPUs at hand
(single rank, or under torchrun): per-stage kernel time and algorithmic GB/s.

    python tools/bench_cases.py --case c2|c4|c5|<a,b,c[,d]> [--dtype D|F|d|f] [--reps 5]

c2 = 512^3 complex128, c4 = 2048^3 float32 r2c/c2r (BASELINE configs[3] on one GPU),
c5 = 4-D c2c complex128 (256^4 needs 8 GPUs; one GPU runs 128x128x256x256).
Algorithmic bytes per stage = bytes read + bytes written of the local block.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='c2')
    ap.add_argument('--dtype', default=None)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--padding', type=float, default=0.0)
    args = ap.parse_args()
    import torch
    import mpi4py_fft_b200 as B
    comm = B.init()
    world = comm.Get_size()
    cases = {'c2': ((512, 512, 512), 'D', {}), 'c4': ((2048, 2048, 2048), 'f', dict(grid=(-1,))),
             'c5': ((256, 256, 256, 256) if world >= 8 else (128, 128, 256, 256), 'D', dict(grid=(4, 2)) if world >= 8 else {})}
    if args.case in cases:
        shape, dtype, kw = cases[args.case]
    else:
        shape, dtype, kw = tuple(int(x) for x in args.case.split(',')), 'D', {}
    if args.dtype:
        dtype = args.dtype
    if args.padding:
        kw['padding'] = [args.padding] * len(shape)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    fft = B.PFFT(comm, shape, dtype=dtype, **kw)
    # the plan's own end-point arrays serve as input and round-trip output (2048^3 leaves no room for more)
    u = fft.forward.input_array
    back = fft.backward.output_array
    real = np.dtype(dtype).kind == 'f'
    ut = B.devarray.as_tensor(u)
    n0 = ut.shape[0]
    for lo in range(0, n0, max(1, n0 // 8)):          # fill in slabs: torch.rand temporaries stay small
        sl = ut[lo:lo + max(1, n0 // 8)]
        if real:
            sl.copy_(torch.rand(tuple(sl.shape), dtype=sl.dtype, device='cuda'))
        else:
            sl.copy_(torch.view_as_complex(torch.rand(tuple(sl.shape) + (2,), dtype=sl.real.dtype, device='cuda')))
    probe = ut[:1].clone()
    stream = torch.cuda.current_stream()
    for _ in range(3):
        uh = fft.forward(u)
        fft.backward(uh, back)
    torch.cuda.synchronize()
    if not args.padding:
        err = float((B.devarray.as_tensor(back)[:1] - probe).abs().max().item())
        assert err < (1e-11 if np.dtype(dtype).itemsize in (8, 16) and dtype in 'dD' else 1e-3), err
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    tf = tb = 0.0
    for _ in range(args.reps):
        e0.record(stream)
        uh = fft.forward(u)
        e1.record(stream)
        fft.backward(uh, back)
        e2.record(stream)
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1)
        tb += e1.elapsed_time(e2)
    tf /= args.reps
    tb /= args.reps
    npts = float(np.prod(fft.global_shape(False)))
    if comm.Get_rank() == 0:
        print("case %s shape %s dtype %s ranks %d grid %s" % (args.case, fft.global_shape(False), dtype, world,
                                                              [c.Get_size() for c in fft.subcomm]))
        print("forward %.3f ms  backward %.3f ms  ->  %.2f GPoints/s (2*points / (fwd+bwd))" % (tf, tb, 2 * npts / (tf + tb) / 1e6))
    # per-stage kernels; the plan's buffers are released first (2048^3 needs the room)
    del u, back, uh, ut, probe, sl
    fft._buffers.arr.clear()
    fft._buffers.work.clear()
    torch.cuda.empty_cache()
    for i, st in enumerate(fft.xfftn):
        for name, d in (('fwd', st.forward), ('bwd', st.backward)):
            a = B.fftw.aligned(d.input_shape, dtype=d.input_dtype, fill=0)
            b = B.fftw.aligned(d.output_shape, dtype=d.output_dtype, fill=0)
            for _ in range(2):
                d.run(a, b)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(args.reps):
                d.run(a, b)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            nbytes = a.nbytes + b.nbytes
            if comm.Get_rank() == 0:
                print("  stage %d %s axes %s %s -> %s : %.3f ms  %.0f GB/s  %.3f of %.0f   [%s]" % (
                    i, name, list(st.axes), tuple(d.input_shape), tuple(d.output_shape), ms, nbytes / ms / 1e6,
                    nbytes / ms / 1e6 / peak, peak, d._planned.plan().describe().strip().replace('\n', ' | ')))
            del a, b
    fft.destroy()


if __name__ == '__main__':
    main()
