"""The device-resident NNDSVD initialisation (rri_nmf_b200/_device_init.py) is plain torch code,
so the very same functions are checked here on CPU tensors against the host implementation (`_host.initialize_nmf`,
itself pinned to initialization.py:80-163 and sklearn's randomized_svd)."""
import numpy as np
import pytest
import torch
from sklearn.utils.extmath import randomized_svd

import rri_oracle as orc
from rri_nmf_b200._device_init import initialize_nmf_torch, randomized_svd_torch
from rri_nmf_b200._host import initialize_nmf


def data(n, d, r, dtype=np.float64, seed=0):
    X, _, _ = orc.synth(n, d, r, r, sigma=0.05, seed=seed, dtype=dtype)
    return X


@pytest.mark.parametrize('n,d,k', [(300, 120, 6),      # n > d, k < 0.1 min(n, d): 7 power iterations
                                    (90, 400, 12),      # n < d: sklearn works on the transpose; 4 power iterations
                                    (64, 64, 5)])
def test_randomized_svd_matches_sklearn(n, d, k):
    X = data(n, d, k)
    U, s, Vt = randomized_svd(X, k, random_state=3)
    Ut, st, Vtt = randomized_svd_torch(torch.from_numpy(X), k, random_state=3)
    assert np.allclose(st.numpy(), s, rtol=1e-10)
    assert np.allclose(Ut.numpy(), U, atol=1e-8) and np.allclose(Vtt.numpy(), Vt, atol=1e-8)


@pytest.mark.parametrize('init', ['nndsvd', 'nndsvda', 'nndsvdar', 'random', 'smart_random', None])
@pytest.mark.parametrize('shape', [(300, 120), (90, 400)])
def test_initialize_matches_host(init, shape):
    n, d = shape
    X = data(n, d, 7, seed=1)
    W, T = initialize_nmf(X, 7, init, random_state=5)
    Wt, Tt = initialize_nmf_torch(torch.from_numpy(X), 7, init, random_state=5)
    assert Wt.shape == (n, 7) and Tt.shape == (7, d)
    assert np.allclose(Wt.numpy(), W, atol=1e-8) and np.allclose(Tt.numpy(), T, atol=1e-8)
    assert (Wt >= 0).all() and (Tt >= 0).all()


def test_float32_input_stays_float32_and_is_close():
    X = data(400, 150, 8, dtype=np.float32, seed=2)
    W, T = initialize_nmf(X, 8, 'nndsvd', random_state=0)
    Wt, Tt = initialize_nmf_torch(torch.from_numpy(X), 8, 'nndsvd', random_state=0)
    assert Wt.dtype == torch.float32 and Tt.dtype == torch.float32
    # fp32 factorisations: agreement to single-precision accuracy relative to the factor scale
    assert np.abs(Wt.numpy() - W).max() < 2e-3 * np.abs(W).max()
    assert np.abs(Tt.numpy() - T).max() < 2e-3 * np.abs(T).max()


def test_transposed_view_is_not_materialised():
    X = torch.from_numpy(data(50, 200, 4))
    base = X.data_ptr()
    randomized_svd_torch(X, 4, random_state=0)
    assert X.data_ptr() == base and X.is_contiguous()


def test_bad_init_name():
    with pytest.raises(ValueError):
        initialize_nmf_torch(torch.ones(5, 4, dtype=torch.float64), 2, 'nope')


@pytest.mark.parametrize('masked', [False, True])
def test_nmf_device_initialisation_step_matches_host_path(masked):
    """nmf()'s device-side initialisation step (W_mat o X, NNDSVD, row normalisations: nmf.py:840-850) against the
    host path it replaces, on CPU tensors"""
    from rri_nmf_b200._host import normalize
    from rri_nmf_b200.nmf import _initialize_on_device
    X = data(200, 90, 6, seed=4)
    M = (np.random.RandomState(1).rand(200, 90) < 0.4).astype(np.uint8) if masked else None
    W, T = initialize_nmf(M * X if masked else X, 6, 'nndsvd', random_state=2)
    T = normalize(T) * 1.0
    W = normalize(W) * 2.0
    Wm = torch.from_numpy(M) if masked else None
    Wt, Tt = _initialize_on_device(torch.from_numpy(X), Wm, 6, 'nndsvd', 2, 1.0, 2.0)
    assert Wt.dtype == torch.float64
    assert np.allclose(Wt.numpy(), W, atol=1e-8) and np.allclose(Tt.numpy(), T, atol=1e-8)
    assert np.allclose(Tt.sum(1).numpy(), 1.0) and np.allclose(Wt.sum(1).numpy(), 2.0)
