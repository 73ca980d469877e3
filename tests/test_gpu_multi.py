"""Multi-GPU parity (NCCL all-to-all inside Transfer): torchrun launches
tests/mp_worker.py with 2, 4 and 8 ranks when the box has that many GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


# every transfer mode on 2 GPUs; the larger worlds (a torchrun start-up each) keep to the modes that differ there
COMBOS = [(2, '1'), (2, 'nopipe'), (2, '0'), (2, 'put'), (2, 'ncclbar'), (4, '1'), (4, 'nopipe'), (4, '0'),
          (8, '1'), (8, 'nopipe'), (8, 'put'), (8, '0')]


@pytest.mark.parametrize('world,p2p', COMBOS)
def test_pfft_over_nccl(world, p2p):
    """p2p=1 (default): stages store straight into the peers' CUDA-IPC windows
    (fused) where they can, pipelined with the consuming stage where the geometry
    allows; nopipe: fused, one launch per stage; put: stage, then put kernel; ncclbar: as the
    default but the group barriers are NCCL all-reduces instead of flag kernels;
    p2p=0: pack -> NCCL send/recv -> unpack"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    only = os.environ.get('B2F_TEST_WORLD')
    if only and int(only) != world:
        pytest.skip("B2F_TEST_WORLD=%s" % only)
    port = 29500 + world + 20 * ['0', '1', 'put', 'nopipe', 'ncclbar'].index(p2p)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(ROOT, 'tests', 'mp_worker.py')]
    env = dict(os.environ, B2F_P2P='0' if p2p == '0' else '1', B2F_FUSED='0' if p2p == 'put' else '1',
               B2F_PIPELINE='0' if p2p == 'nopipe' else '4', B2F_FLAG_BARRIER='0' if p2p == 'ncclbar' else '1')
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and 'MULTI_OK' in r.stdout, r.stdout[-6000:] + r.stderr[-3000:]
    if p2p != '0':
        assert 'transfers=p2p' in r.stdout, r.stdout[-2000:]
