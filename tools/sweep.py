"""Per-kernel timing sweep on one GPU: every variant of fft_configs.h for the
three axes of an S^3 complex128 (or complex64) block; prints GB/s (algorithmic:
one read + one write of the block) and the fraction of the measured HBM peak.

    python tools/sweep.py [--size 512] [--dtype D] [--variants 0-7] [--reps 10] [--shape a,b,c]
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
    ap.add_argument('--shape', default=None, help='a,b,c (overrides --size)')
    ap.add_argument('--dtype', default='D')
    ap.add_argument('--variants', default='0-7')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--axes', default='2,1,0')
    ap.add_argument('--inplace', action='store_true')
    ap.add_argument('--prealloc-gib', type=int, default=0, help='allocate (and touch) this much memory first')
    ap.add_argument('--engine', default='reg', choices=['reg', 'tma'], help='strided axes: register path or TMA-staged')
    args = ap.parse_args()
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
    shape = (S, S, S) if args.shape is None else tuple(int(x) for x in args.shape.split(','))
    variants = []
    for part in args.variants.split(','):
        if '-' in part:
            lo, hi = part.split('-')
            variants += list(range(int(lo), int(hi) + 1))
        else:
            variants.append(int(part))
    ballast = [torch.zeros(1 << 30, dtype=torch.uint8, device='cuda') for _ in range(args.prealloc_gib)]
    a = B.fftw.aligned(shape, dtype=args.dtype)
    b = B.fftw.aligned(shape, dtype=args.dtype)
    a.tensor.copy_(torch.view_as_complex(torch.rand(shape + (2,), dtype=a.tensor.real.dtype, device='cuda')))
    nbytes = 2.0 * a.nbytes
    stream = torch.cuda.current_stream()
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
    print("shape %s dtype %s" % (shape, args.dtype))
    print("copy            %8.3f ms  %7.1f GB/s  %.3f of peak %.0f" % (ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak))
    _lib.set_option('variant_strict', 1)
    nd = len(shape)
    for axis in [int(x) for x in args.axes.split(',')]:
        key = 'variant_contig' if axis == nd - 1 else 'variant_strided'
        if axis != nd - 1:
            _lib.set_option('strided_engine', 2 if args.engine == 'tma' else 1)
            if args.engine == 'tma':
                key = 'variant_tma'
        for var in variants:
            _lib.set_option(key, var)
            for inplace in ((False, True) if args.inplace else (False,)):
                plan = B.fftw.fftn(a, axes=(axis,), output_array=(a if inplace else b))
                try:
                    for _ in range(3):
                        plan()
                except Exception:
                    continue    # variant not built for this length
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(args.reps):
                    plan()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                gbs = nbytes / ms / 1e6
                print("axis %d %s var %d %s %8.3f ms  %7.1f GB/s  %.3f" % (
                    axis, key[8:], var, 'inplace ' if inplace else 'outplace', ms, gbs, gbs / peak), flush=True)
        _lib.set_option(key, 0)
    _lib.set_option('variant_strict', 0)
    _lib.set_option('strided_engine', 0)


if __name__ == '__main__':
    main()
