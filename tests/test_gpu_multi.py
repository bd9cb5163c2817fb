"""Multi-GPU parity: launches tests/multi_gpu_check.py on 2 ranks when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('p2p', ['1', '0'])
def test_two_rank_row_sharded_parity(cuda_device, p2p):
    """p2p=1: NVLink peer-memory exchange fused into the T update; p2p=0: NCCL all-reduce"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '2951' + p2p,
                        os.path.join(ROOT, 'tests', 'multi_gpu_check.py')], env=dict(os.environ, RRI_P2P=p2p),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert 'MULTI_GPU_PARITY PASS' in r.stdout, r.stdout[-3000:]
