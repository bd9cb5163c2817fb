"""Multi-GPU parity: launches tests/multi_gpu_check.py on 2 ranks and on all GPUs of the box (dense, masked,
observed-entries and TF32 paths against the unsharded oracle; T replicas bit-identical on every rank)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('p2p', ['1', '0'])
@pytest.mark.parametrize('ranks', [2, 'all'])
def test_row_sharded_parity(cuda_device, p2p, ranks):
    """p2p=1: NVLink peer-memory exchange fused into the T update; p2p=0: NCCL all-reduce.
    ranks: 2, and every GPU of the box (4 or 8) when it has more than two."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    if ranks == 'all':
        if ngpu <= 2:
            pytest.skip('the box has 2 GPUs: covered by ranks=2')
        ranks = ngpu
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(ranks),
                        '--master-addr', '127.0.0.1', '--master-port', '295%d' % (10 * ranks + int(p2p)),
                        os.path.join(ROOT, 'tests', 'multi_gpu_check.py')], env=dict(os.environ, RRI_P2P=p2p),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert 'MULTI_GPU_PARITY PASS' in r.stdout, r.stdout[-3000:]
