"""Data-parallel PRODUCT code against the single-GPU product code and the oracle (tests/dp_worker.py under torchrun).
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dp_parity.py -m gpu`); `bench.py` carries the same
check at the benchmark's own size in its `parity` key at every N."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import free_port  # noqa: E402


@pytest.mark.gpu
def test_data_parallel_model_matches_single_gpu_model():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs (gpurun --gpus 2)')
    n = 4 if torch.cuda.device_count() >= 4 else 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()),
           os.path.join(ROOT, 'tests', 'dp_worker.py')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    assert 'DP PARITY OK' in r.stdout, r.stdout[-4000:]
