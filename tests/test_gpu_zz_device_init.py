"""Device-resident NNDSVD initialisation on the GPU: same fit as with the host initialisation.
(The torch code itself is checked against the host implementation on CPU: tests/test_device_init_cpu.py.)"""
import numpy as np
import pytest
import torch

import rri_oracle as orc
from conftest import relfro

pytestmark = pytest.mark.gpu


def test_device_init_equals_host_init(cuda_device):
    import rri_nmf_b200 as R
    X, _, _ = orc.synth(600, 250, 8, 8, sigma=0.05, seed=9)
    kw = dict(max_iter=5, random_state=0, eps_stop=-1.0)
    host = R.nmf(X, 8, init_on_device=False, **kw)
    dev = R.nmf(X, 8, init_on_device=True, **kw)
    assert relfro(dev['W'], host['W']) < 1e-6 and relfro(dev['T'], host['T']) < 1e-6
    # a CUDA tensor is initialised where it lives by default and comes back as CUDA tensors
    out = R.nmf(torch.from_numpy(X).to(cuda_device), 8, **kw)
    assert isinstance(out['W'], torch.Tensor) and out['W'].is_cuda
    assert relfro(out['W'].cpu().numpy(), host['W']) < 1e-6 and relfro(out['T'].cpu().numpy(), host['T']) < 1e-6
    # masked: initialisation on W_mat o X (nmf.py:840-843)
    M = (np.random.RandomState(2).rand(600, 250) < 0.5).astype(np.float64)
    a = R.nmf(X, 8, W_mat=M, init_on_device=False, **kw)
    b = R.nmf(X, 8, W_mat=M, init_on_device=True, **kw)
    assert relfro(b['W'], a['W']) < 1e-6 and relfro(b['T'], a['T']) < 1e-6
