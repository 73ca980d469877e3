"""Per-step timing of the rotating schedule (csrc/fft_rot.cuh) on one GPU: each of the
three out-of-place steps of a 3-axis c2c stage alone (option rot_step_mask), for
every variant of B2F_ROT_TABLE, next to the classic per-axis kernels and a copy.

    python tools/sweep_rot.py [--shape 1024,1024,1024] [--dtype D] [--variants 0-4] [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='1024,1024,1024')
    ap.add_argument('--dtype', default='D')
    ap.add_argument('--variants', default='0-4')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--backward', action='store_true')
    ap.add_argument('--steps', default='0', help='which single steps to time alone (0,1,2)')
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
    shape = tuple(int(x) for x in args.shape.split(','))
    variants = []
    for part in args.variants.split(','):
        if '-' in part:
            lo, hi = part.split('-')
            variants += list(range(int(lo), int(hi) + 1))
        else:
            variants.append(int(part))
    a = B.fftw.aligned(shape, dtype=args.dtype)
    b = B.fftw.aligned(shape, dtype=args.dtype)
    a.tensor.copy_(torch.view_as_complex(torch.rand(shape + (2,), dtype=a.tensor.real.dtype, device='cuda')))
    nbytes = 2.0 * a.nbytes
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps=args.reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timeit(lambda: b.tensor.copy_(a.tensor))
    print("shape %s dtype %s" % (shape, args.dtype))
    print("copy                      %8.3f ms  %7.1f GB/s  %.3f of peak %.0f" % (ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak))
    nd = len(shape)
    planner = B.fftw.ifftn if args.backward else B.fftw.fftn
    plan = planner(a, axes=tuple(range(nd - 3, nd)), output_array=b)
    print(plan.plan().describe().strip())
    # classic schedule, whole stage
    _lib.set_option('rotate', 0)
    ms = timeit(plan)
    print("classic 3 axes            %8.3f ms  %7.1f GB/s per axis avg  %.3f" % (ms, 3 * nbytes / ms / 1e6, 3 * nbytes / ms / 1e6 / peak))
    _lib.set_option('rotate', 1)
    names = ['z (in->out)', 'x (out->scratch)', 'y (scratch->out)']
    for var in variants:
        _lib.set_option('variant_rot', var)
        _lib.set_option('rot_step_mask', 7)
        try:
            ms = timeit(plan)
        except Exception as exc:
            print("var %d: %s" % (var, str(exc)[:100]))
            continue
        print("rot var %d  all 3 steps    %8.3f ms  %7.1f GB/s per axis avg  %.3f" % (var, ms, 3 * nbytes / ms / 1e6, 3 * nbytes / ms / 1e6 / peak), flush=True)
        for k in [int(x) for x in args.steps.split(',') if x != '']:
            _lib.set_option('rot_step_mask', 1 << k)
            ms = timeit(plan)
            gbs = nbytes / ms / 1e6
            print("rot var %d  step %-17s %8.3f ms  %7.1f GB/s  %.3f" % (var, names[k], ms, gbs, gbs / peak), flush=True)
    _lib.set_option('rot_step_mask', 7)
    _lib.set_option('variant_rot', -1)
    _lib.set_option('rotate', 0)
    # parity of the default against the classic schedule
    ref = B.fftw.aligned(shape, dtype=args.dtype)
    _lib.set_option('rotate', 0)
    plan(a, ref)
    _lib.set_option('rotate', 1)
    plan(a, b)
    torch.cuda.synchronize()
    err = float((b.tensor - ref.tensor).abs().max() / ref.tensor.abs().max())
    print("rotating vs classic: max rel diff %.3g" % err)


if __name__ == '__main__':
    main()
