"""The reference's own estimator tests (tests/test_nmf.py:81-109) replayed against the device path,
on the same fixtures (stored, preprocessed, in tests/golden/)."""
import numpy as np
import pytest

import rri_oracle as orc
from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def R(cuda_device):
    import rri_nmf_b200
    return rri_nmf_b200


def test_convergence_RS_Estimator(R):
    """tests/test_nmf.py:81-88"""
    X = golden('recsys_wrri_f64.npz')['X']
    n, d = X.shape
    E = R.NMF_RS_Estimator(n, d, 5, random_state=0, max_iter=20)
    E = E.fit_from_Xtr(X)
    assert E.score(X) < 1.0
    assert E.components_ is E.T and E.n_components == 5 and E.reconstruction_err_ > 0
    Wn = E.transform(X)
    assert Wn.shape == (n, 5) and np.all(Wn >= 0)


@pytest.mark.parametrize('sparse', [False, True])
def test_RS_Estimator_early_stop_on_device_matches_reference(R, sparse):
    """the estimator's validation-RMSE early stop (sklearn_interface.py:85-93, nmf.py:381-407) runs on
    device tensors -- W, T and X do not return to the host between sweeps -- and stops where the unmodified
    reference stops on the same fixture (golden rs_estimator_f64.npz: 3 sweeps run, the third reverted)."""
    import torch
    g = golden('rs_estimator_f64.npz')
    X = golden('recsys_wrri_f64.npz')['X']
    n, d = X.shape
    big_d2h = []
    orig_cpu = torch.Tensor.cpu

    def counting_cpu(self, *a, **kw):
        if self.is_cuda and self.numel() >= min(n, d) * 5:
            big_d2h.append(tuple(self.shape))
        return orig_cpu(self, *a, **kw)

    torch.Tensor.cpu = counting_cpu
    try:
        E = R.NMF_RS_Estimator(n, d, 5, random_state=0, max_iter=20, sparse=sparse).fit_from_Xtr(X)
    finally:
        torch.Tensor.cpu = orig_cpu
    # exactly one device -> host copy of each factor (the returned W and T); X and the per-sweep states stay put
    assert sorted(big_d2h) == sorted([(n, 5), (5, d)]), big_d2h
    oh = np.array(E.nmf_outputs['obj_history'])
    assert oh.shape == g['obj_history'].shape and np.allclose(oh, g['obj_history'], rtol=1e-6)
    assert np.allclose(E.W, g['W'], rtol=1e-5, atol=1e-8) and np.allclose(E.T, g['T'], rtol=1e-5, atol=1e-8)
    assert abs(E.score(X) - float(g['score'][0])) < 1e-6
    # reconstruction_err_ = ||M o (X - WT)||_F over the training entries, from the engine
    Xtr_err = E.reconstruction_err_
    assert Xtr_err > 0 and np.isfinite(Xtr_err)
    # the returned objective calculator no longer pins the engine and can be pickled; it can still re-evaluate
    import pickle
    oc = pickle.loads(pickle.dumps(E.nmf_outputs['obj_calculator']))
    assert oc.engine is None and abs(oc.true_objective() - oh[-1]) < 1e-9 * oh[-1]


def test_convergence_TM_Estimator_and_resume(R):
    """tests/test_nmf.py:90-109: ||X-WT|| < ||X|| and fit(2) + 8 x one_iter (+ final projection) == fit(10)"""
    X = golden('text_tm_f64.npz')['X']
    n, d = X.shape
    M = R.NMF_TM_Estimator(n, d, 5, random_state=0, max_iter=10).fit(X)
    assert np.linalg.norm(X - np.dot(M.W, M.T), 'fro') < np.linalg.norm(X, 'fro')
    assert abs(M.reconstruction_err_ - np.linalg.norm(X - np.dot(M.W, M.T), 'fro')) < 1e-9
    M2 = R.NMF_TM_Estimator(n, d, 5, random_state=0, max_iter=2, do_final_project_W=False).fit(X)
    M2.max_iter = 10
    for _ in range(8):
        M2 = M2.one_iter(X)
    M2.W = orc.proj_mat_to_simplex(M2.W)
    assert np.allclose(M2.T, M.T)
    assert np.allclose(M2.W, M.W)
    r2 = M.score(X)
    assert 0.0 < r2 < 1.0
    assert M.transform(X[:10]).shape == (10, 5)


@pytest.mark.parametrize('params', [{'k': 25}, {'k': 15, 'reg_t_l2': 0.1}, {'k': 15, 'reg_t_l2': -0.1},
                                    {'k': 15, 'reg_w_l2': 0.1}])
def test_convergence_tm_setting(R, params):
    """tests/test_nmf.py:22-42 (NNDSVD initialisation on the host, sweeps on the device)"""
    X = golden('text_tm_f64.npz')['X']
    soln = R.nmf(X, max_iter=15, w_row_sum=1.0, random_state=0, eps_stop=1e-4, project_T_each_iter=True,
                 project_W_each_iter=True, compute_obj_each_iter=True, t_row_sum=1.0, early_stop=False, **params)
    oh = soln['obj_history']
    assert np.all(np.diff(oh) <= 0)
    W, T = soln['W'], soln['T']
    assert np.all(W >= -1e-13) and np.all(T >= -1e-13)
    assert np.sum(np.abs(W.sum(1) - 1)) + np.sum(np.abs(T.sum(1) - 1)) <= 1e-12


@pytest.mark.parametrize('params', [{}, {'reg_w_l1': 0.1, 'reg_t_l1': 0.1}, {'reg_w_l1': 0.1}, {'reg_t_l1': 0.1}])
def test_convergence_rs_setting(R, params):
    """tests/test_nmf.py:57-78, incl. the replayed reference objective values"""
    X = golden('recsys_wrri_f64.npz')['X']
    Wm = (X != 0).astype(np.float64)
    soln = R.nmf(X, max_iter=15, random_state=0, W_mat=Wm, compute_obj_each_iter=True, reset_topic_method=None,
                 early_stop=False, k=7, project_T_each_iter=False, t_row_sum=1.0, project_W_each_iter=False,
                 w_row_sum=None, **params)
    oh = soln['obj_history']
    assert np.all(np.diff(oh) <= 1e-9)
    rep = golden('replay_scalars.npz')
    if not params:
        assert np.allclose(oh, rep['rs_plain'], rtol=1e-6)
    if params == {'reg_w_l1': 0.1, 'reg_t_l1': 0.1}:
        assert np.allclose(oh, rep['rs_l1both'], rtol=1e-6)
