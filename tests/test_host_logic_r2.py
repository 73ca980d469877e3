"""Host-side decisions added in round 2, on virtual ranks (no device, no communication):
which chains of stages Transform hands to the library as ONE multi-axis plan
(mpifft.Transform._merge), the chunk-count heuristic of the pipelined redistribution, and the
closed-form input of bench.py (plane waves <-> spectrum) against numpy's FFT."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT


def _pfft(size, rank, shape, **kw):
    import mpi4py_fft_b200 as B
    from mpi4py_fft_b200.comm import virtual_world
    with virtual_world(size, rank):
        return B.PFFT(B.COMM_WORLD, shape, **kw)


def test_merge_only_chains_without_data_movement():
    # one rank, c2c: three single-axis stages become one 3-axis plan, in execution order
    fft = _pfft(1, 0, (64, 32, 16), dtype='D')
    assert len(fft.xfftn) == 3
    assert fft.forward._merged is not None and tuple(fft.forward._merged.axes) == (2, 1, 0)
    assert fft.backward._merged is not None and tuple(fft.backward._merged.axes) == (0, 1, 2)
    assert fft.forward._merged.kind == -1 and fft.backward._merged.kind == 1
    # r2c first stage, padded stages, distributed axes: stage by stage as the reference
    assert _pfft(1, 0, (64, 32, 16), dtype='d').forward._merged is None
    assert _pfft(1, 0, (64, 32, 16), dtype='D', padding=[1.5, 1.5, 1.5]).forward._merged is None
    assert _pfft(4, 1, (64, 32, 16), dtype='D').forward._merged is None
    # a 2-axis group keeps its own plan but still merges with the others
    fft = _pfft(1, 0, (16, 8, 8), dtype='D', axes=((0,), (1, 2)))
    assert fft.forward._merged is not None and sorted(fft.forward._merged.axes) == [0, 1, 2]


def test_merge_can_be_switched_off(monkeypatch):
    monkeypatch.setenv('B2F_MERGE', '0')
    assert _pfft(1, 0, (64, 32, 16), dtype='D').forward._merged is None


def test_pipeline_chunk_heuristic(monkeypatch):
    from mpi4py_fft_b200 import mpifft
    monkeypatch.delenv('B2F_PIPELINE', raising=False)
    monkeypatch.delenv('B2F_FLAG_BARRIER', raising=False)
    assert mpifft.pipeline_chunks(32 << 20) == 0                  # small blocks: not worth a barrier per chunk
    assert mpifft.pipeline_chunks(2 << 30) == 4                   # 8 GPUs, 1024^3
    assert mpifft.pipeline_chunks(4 << 30) == 4
    assert mpifft.pipeline_chunks(8 << 30) == 8                   # 2 GPUs
    assert mpifft.pipeline_chunks(64 << 30) == 8
    monkeypatch.setenv('B2F_FLAG_BARRIER', '0')
    assert mpifft.pipeline_chunks(2 << 30) == 2                   # NCCL barriers: the round-1 rule
    monkeypatch.setenv('B2F_PIPELINE', '6')
    assert mpifft.pipeline_chunks(1) == 6
    assert 30 <= mpifft.pipeline_producer_sms(2, 148) <= 118 and mpifft.pipeline_producer_sms(8, 148) < mpifft.pipeline_producer_sms(2, 148)


@pytest.mark.parametrize('shape,real', [((16, 12, 10), False), ((8, 6, 20), True), ((6, 4, 8, 10), False)])
def test_bench_plane_waves_have_the_claimed_spectrum(shape, real):
    """bench.py's closed form: forward-normalised FFT of the plane-wave sum = a_m at k_m (a_m / 2 in the
    stored half spectrum for real input), zero elsewhere -- checked here with numpy on a small grid"""
    import torch
    spec = importlib.util.spec_from_file_location('b2f_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ks, amps = bench.plane_waves(shape, real, count=3)
    u = torch.zeros(shape, dtype=torch.float64 if real else torch.complex128)
    full = tuple(slice(0, n) for n in shape)
    bench.fill_plane_waves(torch, u, full, shape, real, ks, amps, rows=3)
    x = u.numpy()
    spec_np = (np.fft.rfftn(x) if real else np.fft.fftn(x)) / x.size
    uh = torch.from_numpy(np.ascontiguousarray(spec_np))
    out_slices = tuple(slice(0, n) for n in spec_np.shape)
    assert bench.spectrum_error(torch, uh, out_slices, ks, amps, real) < 1e-14
    # and a wrong spectrum is noticed
    uh2 = torch.from_numpy(np.ascontiguousarray(spec_np))
    uh2[(0,) * len(shape)] += 1e-6
    assert bench.spectrum_error(torch, uh2, out_slices, ks, amps, real) > 5e-7


def _bench():
    spec = importlib.util.spec_from_file_location('b2f_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return bench


class _Args(object):
    def __init__(self, config='c3', size=0, shape=''):
        self.config, self.size, self.shape = config, size, shape


def test_bench_workloads_are_the_baseline_configurations():
    """BASELINE.json configs: c2 512^3 c128, c3 1024^3 c128 (headline), c4 2048^3 f32 r2c slab, c5 256^4 c128 grid (4, 2)"""
    bench = _bench()
    name, shape, dtype, kw = bench.workload(_Args('c3'), 8)
    assert shape == (1024, 1024, 1024) and dtype == 'D' and kw == {} and '1024^3 complex128' in name
    assert bench.workload(_Args('c2'), 1)[1] == (512, 512, 512)
    name, shape, dtype, kw = bench.workload(_Args('c4'), 8)
    assert shape == (2048, 2048, 2048) and dtype == 'f' and kw == dict(grid=(8,)) and 'slab' in name
    name, shape, dtype, kw = bench.workload(_Args('c5'), 8)
    assert shape == (256,) * 4 and dtype == 'D' and tuple(kw['grid']) == (4, 2)
    assert tuple(bench.workload(_Args('c5'), 4)[3]['grid']) == (2, 2)
    # the CPU arm keeps the decomposition KIND on its thread-ranks
    assert bench.reference_kwargs(_Args('c4'), 16) == dict(grid=(16,))
    assert bench.reference_kwargs(_Args('c5'), 16) == dict(grid=(4, 4))
    assert bench.reference_kwargs(_Args('c3'), 16) == {}
    assert bench.metric_name(_Args('c3')) == bench.metric_name(_Args('c2')) == "3D c2c fp64 forward+backward throughput"
    # a reduced shape is labelled as such
    a = _Args('c5', shape='64,64,64,64')
    name, shape, dtype, kw = bench.workload_with_overrides(a, 2)
    assert shape == (64, 64, 64, 64) and 'REDUCED' in name


def test_reference_sample_shrinks_only_when_the_host_cannot_hold_the_workload(monkeypatch):
    bench = _bench()
    import psutil

    class VM(object):
        def __init__(self, avail):
            self.available = avail
    monkeypatch.setattr(psutil, 'virtual_memory', lambda: VM(200 * 2 ** 30))
    assert bench.reference_shape((1024,) * 3, 'D') == (1024,) * 3          # 8.5 x 16 GiB fits in 200 GiB
    monkeypatch.setattr(psutil, 'virtual_memory', lambda: VM(60 * 2 ** 30))
    assert bench.reference_shape((1024,) * 3, 'D') == (512,) * 3
    monkeypatch.setattr(psutil, 'virtual_memory', lambda: VM(8 * 2 ** 30))
    assert bench.reference_shape((1024,) * 3, 'D') == (256,) * 3
