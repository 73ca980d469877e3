#!/usr/bin/env python
"""Benchmark of the PFFT.forward/backward hot path on B200 (contract: see the
task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl b200|reference]

One *step* = one `fft.forward(u)` + one `fft.backward(u_hat)` of a 3-D c2c
complex128 transform (BASELINE.json metric: 3D c2c 1024^3 fp64 GPoints/s),
GPoints/s = 2 * S^3 / t_step / 1e9.  For N > 1 the driver launches this file
under torchrun; the S^3 array is block-distributed over the N ranks (strong
scaling), time is the max over ranks of CUDA-event time on the stream the
kernels run on.

`value`      device-resident arrays (inputs already in HBM).
`e2e`        the same pair through the public API with HOST (pinned) arrays:
             host->device copy of the input block and device->host copy of the
             round-trip result inside the timed region.
`roofline`   the slowest axis kernel: algorithmic bytes (one read + one write of
             the local block = 2 * 16 B * points) / its CUDA-event time, against
             the measured copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` the reference's own PFFT (unmodified, staged under oracle/_ref,
             numpy/pocketfft serial backend because FFTW cannot be built here)
             on the host cores, thread-per-rank fake MPI, bounded sample.
`--impl reference` times only that CPU path and prints it in the same format.
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
    ap.add_argument('--size', type=int, default=int(os.environ.get('B2F_BENCH_SIZE', 1024)))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-size', type=int, default=int(os.environ.get('B2F_BENCH_CPU_SIZE', 256)))
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def workload_name(size):
    return "3D c2c %d^3 complex128 forward+backward" % size


# ---------------------------------------------------------------------------
# CPU arm: the reference's PFFT on host cores
# ---------------------------------------------------------------------------
def cpu_reference(size, steps, warmup, max_ranks=16):
    """Times the reference's own forward+backward on the host.  Returns
    (gpoints_per_s, ms_per_step, info dict)."""
    cores = len(os.sched_getaffinity(0))
    nranks = 1
    while nranks * 2 <= min(max_ranks, cores) and size % (nranks * 2) == 0:
        nranks *= 2
    shape = (size, size, size)
    ref_dir = os.path.join(ROOT, 'oracle', '_ref')
    if os.path.isdir(os.path.join(ref_dir, 'mpi4py_fft')):
        kind = 'reference'
        sys.path.insert(0, os.path.join(ROOT, 'oracle', 'fakempi'))
        sys.path.insert(0, ref_dir)
        from mpi4py import MPI
        from mpi4py_fft import PFFT, newDistArray
        times = []

        def body():
            comm = MPI.COMM_WORLD
            fft = PFFT(comm, shape, dtype='D', backend='numpy')
            u = newDistArray(fft, False)
            rng = np.random.default_rng(comm.Get_rank())
            u[:] = rng.random(u.shape) + 1j * rng.random(u.shape)
            out = []
            for it in range(warmup + steps):
                comm.Barrier()
                t0 = time.perf_counter()
                uh = fft.forward(u)
                ub = fft.backward(uh)
                comm.Barrier()
                out.append(time.perf_counter() - t0)
            return out[warmup:]
        res = MPI.run_ranks(nranks, body)
        per_step = np.max(np.array(res), axis=0)      # max over ranks per step
        t = float(np.mean(per_step))
        what = ("unmodified reference PFFT (numpy/pocketfft serial backend; FFTW not buildable here), "
                "%d thread-ranks over fake MPI" % nranks)
    else:
        kind = 'port'
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pfft_oracle as O
        orc = O.OraclePFFT(nranks, shape, dtype='D')
        g = np.random.default_rng(0).random(shape) + 0j
        blocks = orc.scatter(g)
        ts = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            orc.backward(orc.forward(blocks))
            ts.append(time.perf_counter() - t0)
        t = float(np.mean(ts[warmup:]))
        nranks = 1
        what = "numpy restatement oracle/pfft_oracle.py, one thread"
    gps = 2.0 * size ** 3 / t / 1e9
    return gps, t * 1e3, dict(kind=kind, cores=nranks, host_cores=cores,
                              sample="%d^3 complex128 fwd+bwd, %d steps; %s" % (size, steps, what))


def cpu_best_library(size, steps=3):
    """The strongest CPU library line of the box for context: pocketfft's
    threaded fftn / ifftn on the undistributed array, every host core (not the reference's path)."""
    import scipy.fft as sfft
    cores = len(os.sched_getaffinity(0))
    x = np.random.default_rng(0).random((size,) * 3) + 0j
    sfft.ifftn(sfft.fftn(x, workers=cores), workers=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = sfft.fftn(x, workers=cores)
        sfft.ifftn(y, workers=cores)
    t = (time.perf_counter() - t0) / steps
    return {"value": 2.0 * size ** 3 / t / 1e9, "unit": "GPoints/s", "cores": cores,
            "what": "scipy.fft.fftn + ifftn(workers=%d) on %d^3 complex128, undistributed" % (cores, size)}


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    size = args.cpu_size
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    gps, ms, info = cpu_reference(size, steps, warm)
    line = {
        "impl": "reference", "metric": "3D c2c fp64 forward+backward throughput", "value": gps, "unit": "GPoints/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.size), "sample": info['sample']},
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


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200 import _lib
    from mpi4py_fft_b200.devarray import pinned_empty

    comm = B.init()
    world, rank = comm.Get_size(), comm.Get_rank()
    if world != args.gpus and rank == 0:
        print("note: --gpus %d but world size %d (launch with torchrun for N>1)" % (args.gpus, world), file=sys.stderr)
    S = args.size
    shape = (S, S, S)
    fft = B.PFFT(comm, shape, dtype='D')
    u = B.newDistArray(fft, False)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1234 + rank)
    u.tensor.copy_(torch.view_as_complex(
        torch.rand(tuple(u.shape) + (2,), dtype=torch.float64, device='cuda', generator=gen)))
    back = B.newDistArray(fft, False)
    stream = torch.cuda.current_stream()
    local_points = int(np.prod(u.shape))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        uh = fft.forward(u)
        fft.backward(uh, back)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # correctness guard inside the bench: the round trip must reproduce the input
    err = float((back.tensor - u.tensor).abs().max().item())
    assert err < 1e-11, "round-trip error %g" % err

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
    tt = torch.tensor([t_ms, fwd, bwd], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, fwd, bwd = (float(x) for x in tt.tolist())
    ms_per_step = t_ms / args.steps
    value = 2.0 * S ** 3 / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel timing for the roofline (same arrays, same stream) ----------
    kern = []
    for i, st in enumerate(fft.xfftn):
        s_in = B.fftw.aligned(st.forward.input_shape, dtype=st.forward.input_dtype)
        s_out = B.fftw.aligned(st.forward.output_shape, dtype=st.forward.output_dtype)
        s_in.tensor.copy_(u.tensor.reshape(-1)[:s_in.size].reshape(s_in.shape))
        for _ in range(3):
            st.forward.run(s_in, s_out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, args.steps)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            st.forward.run(s_in, s_out)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = 2.0 * s_in.nbytes
        kern.append(dict(stage=i, axes=list(st.axes), ms=ms, gbs=nbytes / (ms * 1e-3) / 1e9,
                         plan=st.fwd.plan().describe().strip()))
        del s_in, s_out
    worst = max(kern, key=lambda k: k['ms'])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        traffic = prof.get('dram_bytes_per_launch')
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": worst['gbs'], "peak": peak, "unit": "GB/s",
                "frac": worst['gbs'] / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": worst['plan'], "per_stage": [{k: v for k, v in d.items()} for d in kern],
                "whole_step_gbs_per_gpu": 2 * 3 * 2 * 16.0 * local_points / (ms_per_step * 1e-3) / 1e9}

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
            h_in = pinned_empty(u.shape, 'D')
            h_out = pinned_empty(u.shape, 'D')
            h_in[...] = 0.5
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
            td = torch.tensor([dt], dtype=torch.float64, device='cuda')
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
            good = abs(h_out[(0,) * h_out.ndim] - 0.5) < 1e-12
            e2e = {"value": 2.0 * S ** 3 / dt / 1e9 if good else None, "unit": "GPoints/s",
                   "h2d_bytes_per_step": int(h_in.nbytes) * world, "d2h_bytes_per_step": int(h_out.nbytes) * world,
                   "ms_per_step": dt * 1e3, "steps": e2e_steps, "host_memory": "pinned, one block per rank"}
            if not good:
                e2e["error"] = "round trip through host buffers did not reproduce the input"
        del h_in, h_out

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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            gps, ms, info = cpu_reference(args.cpu_size, 2, 1)
            cpu = {"value": gps, "unit": "GPoints/s", "cores": info['cores'], "kind": info['kind'],
                   "sample": info['sample'], "ms_per_step": ms, "host_cores": info['host_cores']}
        except Exception as exc:
            cpu = {"value": None, "unit": "GPoints/s", "error": repr(exc)[:200]}
        try:
            cpu["best_library"] = cpu_best_library(args.cpu_size)
        except Exception as exc:
            cpu["best_library"] = {"value": None, "error": repr(exc)[:200]}

    fused = [v for k, v in list(fft.forward._plan.items()) + list(fft.backward._plan.items())
             if isinstance(k, tuple) and k[0] == 'fused']
    p2p_name = ('stage kernels store into peer windows (fused, CUDA IPC over NVLink)' if fused and all(fused) else
                'fused where possible, else put kernel over CUDA-IPC windows' if any(fused) else
                'put kernel over CUDA-IPC windows')
    piped = [v for k, v in list(fft.forward._plan.items()) + list(fft.backward._plan.items())
             if isinstance(k, tuple) and k[0] == 'pipe' and v]
    if piped:
        p2p_name += '; last redistribution of each direction pipelined with its consumer in %d chunks' % len(piped[0])
    modes = sorted(set(p2p_name if v is not None else 'pack + NCCL send/recv + unpack'
                       for v in fft._buffers.peers.values()))
    transfer_mode = ' / '.join(modes) if modes else ('none (single rank)' if world == 1 else 'pack + NCCL send/recv + unpack')
    if rank == 0:
        grid = [c.Get_size() for c in fft.subcomm]
        line = {
            "metric": "3D c2c fp64 forward+backward throughput", "value": value, "unit": "GPoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": workload_name(S), "grid": grid, "local_shape": list(u.shape),
                       "l2": "inputs larger than L2 (%.1f GiB per array per GPU)" % (u.nbytes / 2 ** 30),
                       "transfer": transfer_mode,
                       "forward_ms_median": fwd, "backward_ms_median": bwd,
                       "forward_gpoints_s": S ** 3 / (fwd * 1e-3) / 1e9, "backward_gpoints_s": S ** 3 / (bwd * 1e-3) / 1e9,
                       "roundtrip_max_err": err},
            "clocks": clk, "roofline": roofline, "nvlink": nvlink, "e2e": e2e, "cpu_baseline": cpu,
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    fft.destroy()          # collective: unmap the peers' windows, then release the own ones
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
