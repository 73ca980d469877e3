"""Multi-GPU parity (NCCL all-to-all inside Transfer): torchrun launches
tests/mp_worker.py with 2, 4 and 8 ranks when the box has that many GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('world', [2, 4, 8])
def test_pfft_over_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + world
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(ROOT, 'tests', 'mp_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'MULTI_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
