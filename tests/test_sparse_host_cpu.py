"""Host-side logic of the observed-entries input path (no GPU): CSR canonicalisation, entry-weight alignment,
what counts as observed in the recommender estimator (sklearn_interface.py:78-83, :100-102), sparse NNDSVD
initialisation, and that nothing falls back to the CPU."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import rri_nmf_b200 as R
from rri_nmf_b200 import _lib
from rri_nmf_b200._host import initialize_nmf
from rri_nmf_b200.nmf import _is_sparse, _sparse_to_device

warnings.filterwarnings('ignore', message='Sparse')


def test_csr_is_canonicalised_sorted_and_duplicates_summed():
    # unsorted triples with a duplicate (2, 1): coo_matrix(...).toarray() sums duplicates, so must we
    I = np.array([2, 0, 2, 1, 2])
    J = np.array([1, 3, 0, 2, 1])
    V = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    X = sp.coo_matrix((V, (I, J)), shape=(3, 4))
    Xc, w = _sparse_to_device(X, None, torch.device('cpu'), torch.float32)
    assert w is None and Xc.layout == torch.sparse_csr and Xc.dtype == torch.float32
    assert Xc.crow_indices().tolist() == [0, 1, 2, 4]
    assert Xc.col_indices().tolist() == [3, 2, 0, 1]
    assert Xc.values().tolist() == [2.0, 4.0, 3.0, 6.0]
    assert np.array_equal(Xc.to_dense().numpy(), X.toarray().astype(np.float32))
    assert _is_sparse(X) and _is_sparse(Xc) and not _is_sparse(X.toarray())


def test_entry_weights_must_share_the_structure():
    X = sp.random(20, 15, 0.3, random_state=0, format='csr')
    Wm = X.copy()
    Wm.data = np.arange(1, X.nnz + 1, dtype=np.float64)
    Xc, w = _sparse_to_device(X.tocoo(), Wm.tocsc(), torch.device('cpu'), torch.float64)
    assert np.array_equal(w.numpy(), Wm.data)              # aligned with the CSR order of X
    other = sp.random(20, 15, 0.3, random_state=1, format='csr')
    with pytest.raises(ValueError):
        _sparse_to_device(X, other, torch.device('cpu'), torch.float64)
    with pytest.raises(ValueError):
        _sparse_to_device(X, np.ones((20, 15)), torch.device('cpu'), torch.float64)


def test_csr_tensor_helper_moves_components_only():
    t = R.RRIEngine.csr_tensor(np.array([0, 2, 3]), np.array([0, 3, 1]), np.array([1.0, 2.0, 3.0]), (2, 4), 'cpu',
                               torch.float32)
    assert t.layout == torch.sparse_csr and t.dtype == torch.float32 and tuple(t.shape) == (2, 4)
    assert t.crow_indices().dtype == torch.int64 and t.col_indices().dtype == torch.int64


def test_estimator_observed_matches_reference_densification():
    """sklearn_interface.py:78-83 + :100-102: X = coo(...).toarray(); W_mat = (X != 0)"""
    rs = np.random.RandomState(3)
    ij = np.column_stack([rs.randint(0, 12, 60), rs.randint(0, 9, 60)])
    r = rs.randint(0, 4, 60).astype(float)                 # zeros and duplicates on purpose
    E = R.NMF_RS_Estimator(12, 9, 3, sparse=True)
    M = E._observed(ij, r)
    dense = sp.coo_matrix((r, (ij[:, 0], ij[:, 1])), shape=(12, 9)).toarray()
    assert np.array_equal(M.toarray(), dense)
    assert M.nnz == int((dense != 0).sum())                # stored entries == observed entries
    assert M.has_sorted_indices


def test_sparse_initialisation_never_densifies_and_agrees():
    X = sp.random(60, 40, 0.2, random_state=2, format='csr')
    Ws, Ts = initialize_nmf(X, 5, 'nndsvd', random_state=0)
    Wd, Td = initialize_nmf(X.toarray(), 5, 'nndsvd', random_state=0)
    assert np.allclose(Ws, Wd, atol=1e-12) and np.allclose(Ts, Td, atol=1e-12)
    Ws, Ts = initialize_nmf(X, 5, 'random', random_state=0)
    Wd, Td = initialize_nmf(X.toarray(), 5, 'random', random_state=0)
    assert np.array_equal(Ws, Wd) and np.array_equal(Ts, Td)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour on a CPU-only host')
def test_sparse_input_has_no_cpu_fallback():
    X = sp.random(30, 20, 0.2, random_state=0, format='csr')
    with pytest.raises(_lib.RriError):
        R.nmf(X, 3, W_in=np.ones((30, 3)), T_in=np.ones((3, 20)), max_iter=1)
    with pytest.raises(_lib.RriError):
        R.RRIEngine(R.RRIEngine.csr_tensor(X.indptr, X.indices, X.data, X.shape, 'cpu'), 3)
