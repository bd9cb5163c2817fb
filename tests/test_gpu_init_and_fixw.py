"""GPU tests of the device-resident initialisation and of fix_W=True:

* device-resident NNDSVD initialisation -- same fit as with the host initialisation; the
  torch code itself is also checked against the host implementation on CPU (tests/test_device_init_cpu.py);
* fix_W=True (T-only sweeps with the W[:, t] *= nt1 rescaling of nmf.py:450-452) against the unmodified reference;
  the host-side sweep is also checked on CPU with the engine calls served by the oracle (tests/test_fixw_cpu.py)."""
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


@pytest.mark.parametrize('name,kw', [('plain', {}), ('simplex', dict(project_T_each_iter=True, t_row_sum=1.0)),
                                      ('reg', dict(reg_t_l1=0.01, reg_t_l2=0.05)), ('masked', dict(masked=True)),
                                      ('masked_ub', dict(masked=True, t_row_sum=1.0))])
def test_fix_W_matches_reference(cuda_device, name, kw):
    """fix_W=True: T-only sweeps + the W[:, t] *= nt1 rescaling of nmf.py:450-452, against the unmodified reference
    (tests/golden/fixW_f64.npz); statistic from rri_partials_T, solve on device vectors"""
    import scipy.sparse as sp
    import rri_nmf_b200 as R
    from conftest import golden
    X, W0, T0, M = orc.synth(120, 80, 5, 6, sigma=0.05, seed=13, mask_density=0.4)
    kw = dict(kw)
    masked = kw.pop('masked', False)
    g = golden('fixW_f64.npz')
    common = dict(W_in=W0, T_in=T0, max_iter=3, fix_W=True, reset_topic_method=None, eps_stop=-1.0, max_time=1e9)
    out = R.nmf(X, 6, W_mat=M if masked else None, **common, **kw)
    assert relfro(out['W'], g['W_' + name]) < 1e-9 and relfro(out['T'], g['T_' + name]) < 1e-9
    if masked:
        I, J = M.nonzero()
        out = R.nmf(sp.csr_matrix((X[I, J], (I, J)), shape=X.shape), 6, **common, **kw)
        assert relfro(out['W'], g['W_' + name]) < 1e-9 and relfro(out['T'], g['T_' + name]) < 1e-9


def test_device_init_known_answer_of_the_reference(cuda_device):
    """The reference's own known-answer test of the initialisation (tests/test_nmf.py:13-19, golden bytes in
    tests/conftest.py:8-19, decoded in SURVEY.md App. B.2) against the DEVICE initialisation, with the passes over
    X going through the engine's contraction kernel (rri_gemm_nt, IEEE fp64) in both handle kinds."""
    import rri_nmf_b200 as R
    from rri_nmf_b200._device_init import initialize_nmf_torch
    X = np.array([[1, 0], [.5, .5], [.25, .75]])
    Wg = np.array([[0.7731182278974053, 0], [0.6108880790649228, 0.2268147629447849],
                   [0.5297730046486816, 0.6538575655189827]])
    Tg = np.array([[0.9676003516472353, 0.5615202893907696], [0, 0.6920799467374485]])
    Xd = torch.from_numpy(X).to(cuda_device)
    for order in ('rri', 'hals'):
        eng = R.RRIEngine(Xd, 2, order=order)
        W, T = initialize_nmf_torch(Xd, 2, 'nndsvd', random_state=0, products=eng.products())
        eng.close()
        assert np.allclose(W.cpu().numpy(), Wg, atol=1e-12) and np.allclose(T.cpu().numpy(), Tg, atol=1e-12), order
    # and through the public call: one T-only-free sweep count of zero is not allowed, so compare the start of a fit
    out = R.nmf(X, 2, init='nndsvd', random_state=0, max_iter=1, init_on_device=True, reset_topic_method=None)
    ref = R.nmf(X, 2, W_in=Wg, T_in=Tg, max_iter=1, reset_topic_method=None)
    assert relfro(out['W'], ref['W']) < 1e-10 and relfro(out['T'], ref['T']) < 1e-10


def test_device_init_through_the_tf32_contraction(cuda_device):
    """float32 data, math='tf32': the 2*n_iter + 2 passes of the randomized SVD run through the tcgen05 contraction
    (k + 10 columns).  The subspace iteration is self-correcting, so the factors agree with the host (fp32 LAPACK)
    initialisation to TF32 accuracy and the fit started from them lands on the same relative error."""
    import rri_nmf_b200 as R
    from rri_nmf_b200._device_init import initialize_nmf_torch
    from rri_nmf_b200._host import initialize_nmf
    X, _, _ = orc.synth(3000, 1200, 16, 16, sigma=0.05, seed=17, dtype=np.float32)
    Wh, Th = initialize_nmf(X, 16, 'nndsvd', random_state=0)
    Xd = torch.from_numpy(X).to(cuda_device)
    eng = R.RRIEngine(Xd, 16, order='hals', math='tf32')
    launches0 = eng.stats()['kernel_launches']
    W, T = initialize_nmf_torch(Xd, 16, 'nndsvd', random_state=0, products=eng.products())
    assert eng.stats()['kernel_launches'] - launches0 >= 10          # the passes went through the engine's kernel
    eng.close()
    assert np.abs(W.cpu().numpy() - Wh).max() < 5e-3 * np.abs(Wh).max()
    assert np.abs(T.cpu().numpy() - Th).max() < 5e-3 * np.abs(Th).max()
    kw = dict(max_iter=5, random_state=0, eps_stop=-1.0, update_order='hals', math='tf32', reset_topic_method=None)
    a = R.nmf(X, 16, init_on_device=True, **kw)
    b = R.nmf(X, 16, init_on_device=False, **kw)
    ra = orc.rel_error(X.astype(np.float64), a['W'].astype(np.float64), a['T'].astype(np.float64))
    rb = orc.rel_error(X.astype(np.float64), b['W'].astype(np.float64), b['T'].astype(np.float64))
    assert abs(ra - rb) < 1e-4, (ra, rb)
