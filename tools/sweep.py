"""Per-kernel timing sweep on one GPU: every variant of fft_configs.h for the
three axes of an S^3 complex128 (or complex64) block; prints GB/s (algorithmic:
one read + one write of the block) and the fraction of the measured HBM peak.

    python tools/sweep.py [--size 512] [--dtype D] [--variants 0,1,2] [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--dtype', default='D')
    ap.add_argument('--variants', default='0,1,2')
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    import numpy as np
    import torch
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200 import _lib
    torch.cuda.set_device(0)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    S = args.size
    shape = (S, S, S)
    a = B.fftw.aligned(shape, dtype=args.dtype)
    b = B.fftw.aligned(shape, dtype=args.dtype)
    a.tensor.copy_(torch.view_as_complex(torch.rand(shape + (2,), dtype=a.tensor.real.dtype, device='cuda')))
    nbytes = 2.0 * a.nbytes
    stream = torch.cuda.current_stream()
    # plain copy for reference
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        b.tensor.copy_(a.tensor)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.reps):
        b.tensor.copy_(a.tensor)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print("copy            %8.3f ms  %7.1f GB/s  %.3f of measured peak %.0f" % (ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak))
    rows = []
    for axis in (2, 1, 0):
        for var in [int(v) for v in args.variants.split(',')]:
            _lib.set_option('variant', var)
            for inplace in (False, True):
                plan = B.fftw.fftn(a, axes=(axis,), output_array=(a if inplace else b))
                for _ in range(3):
                    plan()
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(args.reps):
                    plan()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                gbs = nbytes / ms / 1e6
                rows.append((axis, var, inplace, ms, gbs))
                print("axis %d var %d %s %8.3f ms  %7.1f GB/s  %.3f" % (axis, var, 'inplace ' if inplace else 'outplace', ms, gbs, gbs / peak), flush=True)
    _lib.set_option('variant', 0)


if __name__ == '__main__':
    main()
