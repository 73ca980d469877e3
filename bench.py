#!/usr/bin/env python
"""Benchmark of the PFFT.forward/backward hot path on B200 (contract: see the
task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4|c5] [--impl b200|reference]

One *step* = one `fft.forward(u)` + one `fft.backward(u_hat)`; GPoints/s =
2 * prod(shape) / t_step / 1e9.  The default workload is BASELINE.json's
headline (c3: 3-D c2c 1024^3 complex128); --config / B2F_BENCH_CASE select the
other BASELINE configurations (c2: 512^3 c128, c4: 2048^3 float32 r2c/c2r slab,
c5: 256^4 c128 on a (4, 2) grid).  For N > 1 the driver launches this file
under torchrun; the global array is block-distributed over the N ranks (strong
scaling), time is the max over ranks of CUDA-event time on the stream the
kernels run on.

`config.forward_max_err`  checked BEFORE the timed region at every N: the input is a
             sum of plane waves whose normalised spectrum is known in closed
             form (a_m at wavenumber k_m, zero elsewhere); every rank compares
             its whole block of the forward output with it (max over ranks).
`value`      device-resident arrays (inputs already in HBM).
`e2e`        the same pair through the public API with HOST (pinned) arrays:
             host->device copy of the input block and device->host copy of the
             round-trip result inside the timed region.
`roofline`   the slowest stage kernel OF THE TIMED STEP: algorithmic bytes (one
             read + one write of the local block) / its CUDA-event time, against
             the measured copy bandwidth in MEASURED_PEAKS.json.  N = 1: each
             kernel of the plan timed alone on the plan's stream; N > 1: the
             stages are timed inside a real forward (events between stages, max
             over ranks), i.e. the fused peer-store / pipelined variants.
`cpu_baseline` the reference's own PFFT (unmodified, staged under oracle/_ref,
             numpy/pocketfft serial backend because FFTW cannot be built here)
             on the host cores, thread-per-rank fake MPI, bounded sample.
`--impl reference` times only that CPU path on the SAME workload, honouring
             --steps / --warmup, and prints it in the same format.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default=os.environ.get('B2F_BENCH_CASE', 'c3'))
    ap.add_argument('--size', type=int, default=int(os.environ.get('B2F_BENCH_SIZE', 0)),
                    help='cube edge of the c2c workload (overrides --config; debugging)')
    ap.add_argument('--shape', default=os.environ.get('B2F_BENCH_SHAPE', ''),
                    help='a,b,c[,d]: global shape instead of the configuration\'s (debugging a configuration at reduced size)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-size', type=int, default=int(os.environ.get('B2F_BENCH_CPU_SIZE', 0)),
                    help='cube edge of the cpu_baseline sample (0: the workload itself when the host has the memory)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def workload(args, world):
    """(name, global shape, dtype char, PFFT keyword arguments)"""
    if args.size:
        S = args.size
        return "3D c2c %d^3 complex128 forward+backward" % S, (S, S, S), 'D', {}
    c = args.config
    if c == 'c3':
        return "3D c2c 1024^3 complex128 forward+backward", (1024,) * 3, 'D', {}
    if c == 'c2':
        return "3D c2c 512^3 complex128 forward+backward", (512,) * 3, 'D', {}
    if c == 'c4':
        return "3D r2c/c2r 2048^3 float32 slab forward+backward", (2048,) * 3, 'f', dict(grid=(world,))
    if c == 'c5':
        grid = {1: (1, 1), 2: (2, 1), 4: (2, 2)}.get(world, (world // 2, 2))
        return "4D c2c 256^4 complex128 grid %s forward+backward" % (tuple(grid),), (256,) * 4, 'D', dict(grid=grid)
    raise SystemExit("unknown --config %r" % c)


def workload_with_overrides(args, world):
    name, shape, dtype, kw = workload(args, world)
    if args.shape:
        shape = tuple(int(x) for x in args.shape.split(','))
        name += " [REDUCED to %s]" % 'x'.join(str(x) for x in shape)
    return name, shape, dtype, kw


def metric_name(args):
    if args.size or args.config in ('c2', 'c3'):
        return "3D c2c fp64 forward+backward throughput"
    return {"c4": "3D r2c/c2r fp32 forward+backward throughput", "c5": "4D c2c fp64 forward+backward throughput"}[args.config]


def reference_kwargs(args, nranks):
    """the same decomposition KIND on the host's thread-ranks: slab for c4, two distributed axes
    for c5, the reference's default (pencil) otherwise"""
    if args.size:
        return {}
    if args.config == 'c4':
        return dict(grid=(nranks,))
    if args.config == 'c5':
        a = 1
        while a * a < nranks:
            a *= 2
        return dict(grid=(a, nranks // a))
    return {}


# ---------------------------------------------------------------------------
# CPU arm: the reference's PFFT on host cores
# ---------------------------------------------------------------------------
def cpu_reference(shape, dtype, kw_of, steps, warmup, budget_s=None):
    """Times the reference's own forward+backward on the host with one thread-rank per
    core.  Returns (gpoints_per_s, ms_per_step, info dict)."""
    cores = len(os.sched_getaffinity(0))
    nranks = 1
    while nranks * 2 <= cores and all(s % (nranks * 2) == 0 for s in shape[:2]):
        nranks *= 2
    ref_dir = os.path.join(ROOT, 'oracle', '_ref')
    npts = float(np.prod(shape))
    steps_run = steps
    if os.path.isdir(os.path.join(ref_dir, 'mpi4py_fft')):
        kind = 'reference'
        for pth in (os.path.join(ROOT, 'oracle', 'fakempi'), ref_dir):
            if pth not in sys.path:
                sys.path.insert(0, pth)
        from mpi4py import MPI
        from mpi4py_fft import PFFT, newDistArray

        def body(shp, backend, nwarm, nsteps, budget):
            comm = MPI.COMM_WORLD
            fft = PFFT(comm, shp, dtype=dtype, backend=backend, **kw_of(nranks))
            u = newDistArray(fft, False)
            rng = np.random.default_rng(comm.Get_rank())
            u[:] = rng.random(u.shape)
            out = []
            done = 0
            t_begin = time.perf_counter()
            for it in range(nwarm + nsteps):
                comm.Barrier()
                t0 = time.perf_counter()
                uh = fft.forward(u)
                fft.backward(uh)
                comm.Barrier()
                t1 = time.perf_counter()
                out.append(t1 - t0)
                done += 1
                # every rank takes the same decision: rank 0's clock decides
                stop = comm.bcast(budget is not None and it + 1 >= nwarm + 1 and
                                  (t1 - t_begin) + (t1 - t0) > budget, root=0)
                if stop:
                    break
            return out[min(nwarm, done - 1):]

        # the faster of the reference's two serial backends that exist here (a short probe)
        backend = 'numpy'
        try:
            probe_shape = tuple(min(s, 128) for s in shape)
            tt = {}
            for bk in ('numpy', 'scipy'):
                tt[bk] = float(np.mean(np.max(np.array(MPI.run_ranks(nranks, body, probe_shape, bk, 1, 2, None)), axis=0)))
            backend = min(tt, key=tt.get)
        except Exception:
            pass
        res = MPI.run_ranks(nranks, body, tuple(shape), backend, warmup, steps, budget_s)
        per_step = np.max(np.array(res), axis=0)      # max over ranks per step
        steps_run = len(per_step)
        t = float(np.mean(per_step))
        what = ("unmodified reference PFFT (%s/pocketfft serial backend; FFTW not buildable here), "
                "%d thread-ranks over fake MPI" % (backend, nranks))
    else:
        kind = 'port'
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pfft_oracle as O
        orc = O.OraclePFFT(1, shape, dtype=dtype)
        g = np.random.default_rng(0).random(shape).astype(dtype)
        blocks = orc.scatter(g)
        ts = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            orc.backward(orc.forward(blocks))
            ts.append(time.perf_counter() - t0)
        t = float(np.mean(ts[warmup:]))
        nranks = 1
        what = "numpy restatement oracle/pfft_oracle.py, one thread"
    gps = 2.0 * npts / t / 1e9
    return gps, t * 1e3, dict(kind=kind, cores=nranks, host_cores=cores, steps=steps_run,
                              sample="%s %s fwd+bwd, %d steps; %s" % ('x'.join(str(s) for s in shape),
                                                                      np.dtype(dtype).name, steps_run, what))


def host_fits(shape, dtype, factor=8.5):
    """the reference path holds about 8 arrays of the global size (measured: 16.0 GiB resident for a
    2 GiB 512^3 complex128 problem on 8 thread-ranks)"""
    try:
        import psutil
        itemsize = np.dtype(dtype).itemsize * (2 if np.dtype(dtype).kind == 'f' else 1)
        need = float(np.prod(shape)) * itemsize * factor
        return psutil.virtual_memory().available > need * 1.1, need
    except Exception:
        return False, 0.0


def reference_shape(shape, dtype):
    """the workload itself when the host has the memory, else the largest halved cube that
    fits (flagged as a reduced sample)"""
    shp = tuple(shape)
    while not host_fits(shp, dtype)[0] and max(shp) > 64:
        shp = tuple(max(64, s // 2) for s in shp)
    return shp


def cpu_best_library(shape, dtype, steps=2):
    """The strongest CPU library line of the box for context: pocketfft's
    threaded fftn / ifftn on the undistributed array, every host core (not the reference's path)."""
    import scipy.fft as sfft
    cores = len(os.sched_getaffinity(0))
    real = np.dtype(dtype).kind == 'f'
    x = np.random.default_rng(0).random(shape).astype(dtype)
    fwd = (lambda a: sfft.rfftn(a, workers=cores)) if real else (lambda a: sfft.fftn(a, workers=cores))
    bwd = (lambda a: sfft.irfftn(a, s=shape, workers=cores)) if real else (lambda a: sfft.ifftn(a, workers=cores))
    bwd(fwd(x))
    t0 = time.perf_counter()
    for _ in range(steps):
        bwd(fwd(x))
    t = (time.perf_counter() - t0) / steps
    return {"value": 2.0 * float(np.prod(shape)) / t / 1e9, "unit": "GPoints/s", "cores": cores,
            "what": "scipy.fft %s(workers=%d) on %s %s, undistributed" % (
                'rfftn + irfftn' if real else 'fftn + ifftn', cores, 'x'.join(str(s) for s in shape), np.dtype(dtype).name)}


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', args.gpus))
    name, shape, dtype, kw = workload_with_overrides(args, max(world, args.gpus))
    shp = reference_shape(shape, dtype)
    reduced = tuple(shp) != tuple(shape)
    budget = float(os.environ.get('B2F_REF_BUDGET_S', 420))
    gps, ms, info = cpu_reference(shp, dtype, lambda n: reference_kwargs(args, n), max(1, args.steps), max(0, args.warmup),
                                   budget_s=budget)
    cfg = {"workload": name, "sample": info['sample'], "host_cores": info['host_cores'],
           "sample_shape": list(shp), "reduced_sample": reduced}
    if reduced:
        cfg["extrapolated"] = True
        cfg["note"] = "host memory too small for the workload through the reference path; value is the reduced cube's"
    if info['steps'] != args.steps:
        cfg["steps_reduced"] = "time budget %.0f s reached after %d timed steps" % (budget, info['steps'])
    line = {
        "impl": "reference", "metric": metric_name(args), "value": gps, "unit": "GPoints/s",
        "n_gpus": args.gpus, "steps": info['steps'], "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if dtype in 'dD' else "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": gps, "unit": "GPoints/s", "cores": info['cores'], "kind": info['kind'],
                         "sample": info['sample']},
        "e2e": {"value": gps, "unit": "GPoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------
class Clocks(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def gpu_local_cpus(torch, dev):
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), or None"""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        cpus = set()
        for part in txt.split(','):
            if '-' in part:
                lo, hi = part.split('-')
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus if cpus and cpus != allowed else None
    except Exception:
        return None


# ---------------------------------------------------------------------------
# closed-form input: a sum of plane waves
# ---------------------------------------------------------------------------
def plane_waves(shape, real, seed=7, count=5):
    """wavenumbers k_m and amplitudes a_m of the synthetic input
        u(r) = sum_m a_m exp(2 pi i k_m . r / shape)            (complex input)
        u(r) = sum_m a_m cos(2 pi k_m . r / shape + phi_m)      (real input)
    whose forward-normalised spectrum is a_m (resp. a_m exp(i phi_m) / 2) at k_m and
    zero at every other stored wavenumber."""
    rng = np.random.default_rng(seed)
    ks, amps = [], []
    while len(ks) < count:
        k = tuple(int(rng.integers(0, n)) for n in shape)
        if real:
            # keep +k inside the stored half spectrum and -k outside it
            k = k[:-1] + (int(rng.integers(1, max(2, shape[-1] // 2))),)
        if k in ks:
            continue
        ks.append(k)
        amps.append(complex(rng.uniform(0.5, 1.5) * np.exp(2j * np.pi * rng.uniform())))
    return ks, amps


def fill_plane_waves(torch, u_t, local_slice, shape, real, ks, amps, rows=None):
    """u_t (this rank's block, torch tensor) <- the plane-wave sum at its global indices, built
    slab by slab along the first axis so that temporaries stay small"""
    nd = len(shape)
    cdt = torch.complex128 if u_t.dtype in (torch.float64, torch.complex128) else torch.complex64
    starts = [s.start for s in local_slice]
    n0 = u_t.shape[0]
    rest = int(np.prod(u_t.shape[1:])) or 1
    rows = rows or max(1, min(n0, (1 << 26) // rest))          # ~1 GiB of complex128 per slab
    facs = []
    for k in ks:
        f = []
        for ax in range(nd):
            # exact phase reduction in integers: (k * x) mod n, so that cos / sin see an argument below 2 pi
            idx = torch.arange(starts[ax], starts[ax] + u_t.shape[ax], device=u_t.device, dtype=torch.int64)
            ph = ((idx * int(k[ax])) % int(shape[ax])).to(torch.float64) * (2.0 * np.pi / shape[ax])
            f.append(torch.complex(torch.cos(ph), torch.sin(ph)).to(cdt))
        facs.append(f)
    for lo in range(0, n0, rows):
        hi = min(n0, lo + rows)
        acc = None
        for f, a in zip(facs, amps):
            w = f[0][lo:hi] * a
            for ax in range(1, nd):
                w = w.reshape(w.shape + (1,)) * f[ax].reshape((1,) * ax + (-1,))
            acc = w if acc is None else acc.add_(w)
        u_t[lo:hi].copy_(acc.real if real else acc)
        del acc, w


def spectrum_error(torch, uh_t, local_slice, ks, amps, real):
    """max |uh - expected| over this rank's block of the forward output (uh_t is modified)"""
    for k, a in zip(ks, amps):
        inside = all(s.start <= kk < s.stop for kk, s in zip(k, local_slice))
        if inside:
            idx = tuple(kk - s.start for kk, s in zip(k, local_slice))
            uh_t[idx] -= (a / 2 if real else a)
    err = 0.0
    n0 = uh_t.shape[0]
    rest = int(np.prod(uh_t.shape[1:])) or 1
    rows = max(1, min(n0, (1 << 27) // rest))
    for lo in range(0, n0, rows):
        err = max(err, float(torch.view_as_real(uh_t[lo:lo + rows]).abs().max().item()))
    return err


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200 import _lib
    from mpi4py_fft_b200.devarray import pinned_empty, as_tensor

    comm = B.init()
    world, rank = comm.Get_size(), comm.Get_rank()
    if world != args.gpus and rank == 0:
        print("note: --gpus %d but world size %d (launch with torchrun for N>1)" % (args.gpus, world), file=sys.stderr)
    name, shape, dtype, kw = workload_with_overrides(args, world)
    real = dtype in 'fd'
    f64 = dtype in 'dD'
    npts = float(np.prod(shape))
    fft = B.PFFT(comm, shape, dtype=dtype, **kw)
    u = B.newDistArray(fft, False)
    back = B.newDistArray(fft, False)
    stream = torch.cuda.current_stream()
    local_points = int(np.prod(u.shape))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        tt = torch.tensor(list(vals), dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(x) for x in tt.tolist()]

    # ---- forward VALUES against the closed form, whole block of every rank -----------
    ks, amps = plane_waves(shape, real)
    fill_plane_waves(torch, as_tensor(u), fft.local_slice(False), shape, real, ks, amps)
    uh = fft.forward(u)
    ferr = spectrum_error(torch, as_tensor(uh), fft.local_slice(True), ks, amps, real)
    ferr, = allmax([ferr])
    tol_f = 1e-12 if f64 else 1e-5
    assert ferr <= tol_f, "forward transform differs from the closed-form spectrum: max err %g" % ferr

    def step():
        uh = fft.forward(u)
        fft.backward(uh, back)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # the round trip must reproduce the input
    err = 0.0
    ut, bt = as_tensor(u), as_tensor(back)
    rows = max(1, min(ut.shape[0], (1 << 27) // (int(np.prod(ut.shape[1:])) or 1)))
    for lo in range(0, ut.shape[0], rows):
        err = max(err, float((bt[lo:lo + rows] - ut[lo:lo + rows]).abs().max().item()))
    err, = allmax([err])
    assert err < (1e-11 if f64 else 1e-4), "round-trip error %g" % err

    clocks = Clocks(int(os.environ.get('LOCAL_RANK', 0)))
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    # per-direction marks inside the same timed region (t_fwd, t_bwd)
    mid = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    end = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev[0].record(stream)
    for k in range(args.steps):
        uh = fft.forward(u)
        mid[k].record(stream)
        fft.backward(uh, back)
        end[k].record(stream)
    ev[1].record(stream)
    barrier()
    t_ms = ev[0].elapsed_time(ev[1])
    launches = _lib.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    starts = [ev[0]] + end[:-1]
    fwd = float(np.median([a.elapsed_time(b) for a, b in zip(starts, mid)]))
    bwd = float(np.median([a.elapsed_time(b) for a, b in zip(mid, end)]))
    t_ms, fwd, bwd = allmax([t_ms, fwd, bwd])
    ms_per_step = t_ms / args.steps
    value = 2.0 * npts / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel timing for the roofline: the kernels the timed step runs --------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    kern = []
    reps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_local(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    merged = fft.forward._merged
    how = None
    if world == 1 and merged is not None:
        desc = merged.plan().describe().strip().split('\n')
        rot = [ln for ln in desc if 'rotating' in ln]
        classic = [ln for ln in desc if 'rotating' not in ln]
        if rot and os.environ.get('B2F_ROTATE', '0') not in ('0', ''):
            how = "each kernel of the rotating schedule alone (option rot_step_mask), arrays of the timed step"
            for k, ln in enumerate(rot):
                _lib.set_option('rot_step_mask', 1 << k)
                ms = time_local(lambda: merged.execute(u, uh, 1.0))
                kern.append(dict(stage=k, ms=ms, gbs=2.0 * u.nbytes / (ms * 1e-3) / 1e9, plan=ln.strip()))
            _lib.set_option('rot_step_mask', 7)
        else:
            how = "each single-axis kernel of the stage alone, arrays of the timed step"
            for k, st in enumerate(fft.xfftn):
                ms = time_local(lambda: st.forward.run(u if k == 0 else uh, uh))
                kern.append(dict(stage=k, axes=list(st.axes), ms=ms, gbs=2.0 * u.nbytes / (ms * 1e-3) / 1e9,
                                 plan=classic[k].strip() if k < len(classic) else ''))
    else:
        # stages timed INSIDE a forward: events after every stage (+ its redistribution) on the plan's
        # stream, median of `reps` forwards, max over ranks -- these are the fused / pipelined variants
        how = ("stages timed inside a forward (events between stages; a stage's time includes the redistribution "
               "fused into it), max over ranks")
        acc = None
        for _ in range(reps):
            fft.forward._marks = []
            fft.forward(u)
            marks = fft.forward._marks
            fft.forward._marks = None
            torch.cuda.synchronize()
            ms = [marks[i][1].elapsed_time(marks[i + 1][1]) for i in range(len(marks) - 1)]
            acc = [ms] if acc is None else acc + [ms]
        labels = [m[0] for m in marks[1:]]
        med = allmax(np.median(np.array(acc), axis=0))
        for lab, ms in zip(labels, med):
            sts = [fft.xfftn[i] for i in lab]
            nbytes = sum(int(np.prod(s.forward.input_shape)) * np.dtype(s.forward.input_dtype).itemsize +
                         int(np.prod(s.forward.output_shape)) * np.dtype(s.forward.output_dtype).itemsize for s in sts)
            kern.append(dict(stage=list(lab), axes=[list(s.axes) for s in sts], ms=ms, gbs=nbytes / (ms * 1e-3) / 1e9,
                             plan=' | '.join(s.fwd.plan().describe().strip().split('\n')[0] for s in sts)))
    worst = max(kern, key=lambda k: k['ms'] / (len(k['stage']) if isinstance(k['stage'], list) else 1))
    traffic = None
    if world == 1:
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
            if prof.get('workload') == name:
                traffic = prof.get('dram_bytes_per_launch')
        except Exception:
            pass
    itemsize_in = np.dtype(fft.dtype(False)).itemsize
    itemsize_out = np.dtype(fft.dtype(True)).itemsize
    bytes_dir = sum(int(np.prod(s.forward.input_shape)) * np.dtype(s.forward.input_dtype).itemsize +
                    int(np.prod(s.forward.output_shape)) * np.dtype(s.forward.output_dtype).itemsize for s in fft.xfftn)
    roofline = {"bound": "hbm", "achieved": worst['gbs'], "peak": peak, "unit": "GB/s",
                "frac": worst['gbs'] / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": worst['plan'], "how": how, "per_stage": kern,
                "algorithmic_bytes_per_direction_per_gpu": bytes_dir,
                "whole_step_gbs_per_gpu": 2.0 * bytes_dir / (ms_per_step * 1e-3) / 1e9}

    # ---- end to end with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        # every rank stages its own block through pinned host memory; whether the box has the
        # room is decided collectively (a rank that bails out alone would strand the others)
        h_in = h_out = None
        why = ''
        try:
            import psutil
            need = 2 * u.nbytes * world          # all ranks share this host
            avail = psutil.virtual_memory().available
            if avail < need * 1.5:
                raise MemoryError("host has %.0f GiB available, pinned staging needs %.0f GiB"
                                  % (avail / 2 ** 30, need / 2 ** 30))
            # NUMA-local staging: pages are placed on the node of the allocating thread, so the rank
            # allocates (and first touches) its pinned blocks while bound to the CPUs next to its GPU
            # -- with 8 ranks on a two-socket host half the copies otherwise cross the socket link
            before = os.sched_getaffinity(0)
            local = gpu_local_cpus(torch, torch.cuda.current_device())
            if local:
                os.sched_setaffinity(0, local)
            try:
                h_in = pinned_empty(u.shape, fft.dtype(False))
                h_out = pinned_empty(u.shape, fft.dtype(False))
                h_in[...] = 0.5
                h_out[...] = 0
            finally:
                os.sched_setaffinity(0, before)
            numa_note = "allocated on the GPU's NUMA node (cpus %d..%d)" % (min(local), max(local)) if local else \
                "single NUMA domain or no sysfs information"
        except Exception as exc:   # e.g. not enough host RAM for pinned staging buffers
            why = repr(exc)[:200]
            h_in = h_out = None
        flag = torch.tensor([1.0 if h_in is not None else 0.0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) < 0.5:
            e2e = {"value": None, "unit": "GPoints/s", "error": why or "another rank could not stage its block"}
        else:
            e2e_steps = max(1, min(args.steps, 3))

            def e2e_step():
                uh = fft.forward(h_in)           # host -> device copy inside
                fft.backward(uh, h_out)          # device -> host copy inside
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            dt = (time.perf_counter() - t0) / e2e_steps
            dt, = allmax([dt])
            good = abs(h_out[(0,) * h_out.ndim] - 0.5) < (1e-12 if f64 else 1e-5)
            e2e = {"value": 2.0 * npts / dt / 1e9 if good else None, "unit": "GPoints/s",
                   "h2d_bytes_per_step": int(h_in.nbytes) * world, "d2h_bytes_per_step": int(h_out.nbytes) * world,
                   "ms_per_step": dt * 1e3, "steps": e2e_steps, "host_memory": "pinned, one block per rank; " + numa_note}
            if not good:
                e2e["error"] = "round trip through host buffers did not reproduce the input"
        del h_in, h_out
        try:
            torch._C._host_emptyCache()      # give the pinned staging blocks back before the CPU leg
        except Exception:
            pass

    # bytes every GPU pushes over NVLink per step (both directions of the transform): the share
    # (p-1)/p of the block each redistribution moves; step time bounds the achieved rate from below
    nvlink = None
    if world > 1:
        sent = 0
        for t in fft.transfer:
            p = t.comm.Get_size()
            if p > 1:
                sent += 2 * int(np.prod(t.subshapeA)) * t.dtype.itemsize * (p - 1) // p
        nvlink = {"bytes_per_gpu_per_step": int(sent), "min_gbs_per_gpu": sent / (ms_per_step * 1e-3) / 1e9,
                  "reference_gbs": 770.0, "note": "peer-copy reference of this pool (B200_PROFILING.md); "
                  "redistributions are fused into the stages, so their NVLink time is not separable"}

    fused = [v for k, v in list(fft.forward._plan.items()) + list(fft.backward._plan.items())
             if isinstance(k, tuple) and k[0] == 'fused']
    p2p_name = ('stage kernels store into peer windows (fused, CUDA IPC over NVLink)' if fused and all(fused) else
                'fused where possible, else put kernel over CUDA-IPC windows' if any(fused) else
                'put kernel over CUDA-IPC windows')
    piped = [v for k, v in list(fft.forward._plan.items()) + list(fft.backward._plan.items())
             if isinstance(k, tuple) and k[0] == 'pipe' and v]
    if piped:
        p2p_name += '; %d redistribution(s) pipelined with their consumer in %d chunks' % (len(piped), len(piped[0]))
    modes = sorted(set(p2p_name if v is not None else 'pack + NCCL send/recv + unpack'
                       for v in fft._buffers.peers.values()))
    transfer_mode = ' / '.join(modes) if modes else ('none (single rank)' if world == 1 else 'pack + NCCL send/recv + unpack')
    grid = [c.Get_size() for c in fft.subcomm]
    local_shape = list(u.shape)
    nbytes_u = u.nbytes
    torch.cuda.synchronize()
    fft.destroy()          # collective: unmap the peers' windows, then release the own ones
    del u, back, uh
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # bounded sample of the same workload on the host cores: the workload itself (one timed
        # step) when the host has the memory, else the largest halved cube that fits
        try:
            cshape = (args.cpu_size,) * len(shape) if args.cpu_size else reference_shape(shape, dtype)
            gps, ms, info = cpu_reference(cshape, dtype, lambda n: reference_kwargs(args, n), 1, 1)
            cpu = {"value": gps, "unit": "GPoints/s", "cores": info['cores'], "kind": info['kind'],
                   "sample": info['sample'], "ms_per_step": ms, "host_cores": info['host_cores'],
                   "reduced_sample": tuple(cshape) != tuple(shape)}
        except Exception as exc:
            cpu = {"value": None, "unit": "GPoints/s", "error": repr(exc)[:200]}
        try:
            cpu["best_library"] = cpu_best_library(tuple(min(s, 512) for s in shape), dtype)
        except Exception as exc:
            cpu["best_library"] = {"value": None, "error": repr(exc)[:200]}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": "GPoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if f64 else "f32",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": name, "grid": grid, "local_shape": local_shape,
                       "l2": "inputs larger than L2 (%.1f GiB per array per GPU)" % (nbytes_u / 2 ** 30),
                       "transfer": transfer_mode,
                       "forward_ms_median": fwd, "backward_ms_median": bwd,
                       "forward_gpoints_s": npts / (fwd * 1e-3) / 1e9, "backward_gpoints_s": npts / (bwd * 1e-3) / 1e9,
                       "forward_max_err": ferr, "forward_check": "whole block of every rank vs the closed-form spectrum "
                       "of a %d-plane-wave input, tolerance %g" % (len(ks), tol_f),
                       "roundtrip_max_err": err},
            "clocks": clk, "roofline": roofline, "nvlink": nvlink, "e2e": e2e, "cpu_baseline": cpu,
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
