"""The observed-entries WRRI scheme (two residual copies + one pending rank-one record per copy, as orchestrated
in rri_nmf_b200/csrc/api.cu: sp_T_step / sp_W_step / sp_sweeps) checked on the CPU through its NumPy model
(oracle/sparse_scheme.py) against the golden vectors of the unmodified reference."""
import numpy as np
import pytest

import rri_oracle as orc
from conftest import golden, relfro
from sparse_scheme import SparseWRRI


@pytest.mark.parametrize('name,regs', [('plain', {}), ('l1both', {'reg_w_l1': 0.1, 'reg_t_l1': 0.1}),
                                        ('l1w', {'reg_w_l1': 0.1}), ('l1t', {'reg_t_l1': 0.1})])
def test_scheme_matches_reference_recsys(name, regs):
    g = golden('recsys_wrri_f64.npz')
    X = g['X']
    I, J = X.nonzero()
    S = SparseWRRI(I, J, X[I, J], *X.shape)
    W, T = np.maximum(g['W0'], 0).copy(), np.maximum(g['T0'], 0).copy()
    S.sweeps(W, T, 15, ub_t=1.0, check_copies=True, **regs)
    assert relfro(W, g['W_' + name]) < 1e-9 and relfro(T, g['T_' + name]) < 1e-9
    half_sq, _ = S.objective_terms(W, T)
    obj = half_sq + regs.get('reg_w_l1', 0) * np.abs(W).sum() + regs.get('reg_t_l1', 0) * np.abs(T).sum()
    assert abs(obj - g['obj_' + name][-1]) <= 1e-9 * abs(obj)


def test_scheme_entry_weights_both_orders():
    X, W0, T0, Mb = orc.synth(120, 90, 5, 6, sigma=0.05, seed=5, mask_density=0.3)
    Mw = Mb * np.random.RandomState(9).rand(120, 90) * 2.0
    I, J = Mb.nonzero()
    S = SparseWRRI(I, J, X[I, J], 120, 90, weights=Mw[I, J])
    g = golden('weighted_wrri_f64.npz')
    W, T = np.maximum(W0, 0).copy(), np.maximum(T0, 0).copy()
    S.sweeps(W, T, 8, check_copies=True)
    assert relfro(W, g['W']) < 1e-9 and relfro(T, g['T']) < 1e-9
    g = golden('weighted_wrri_hals_f64.npz')
    W, T = np.maximum(W0, 0).copy(), np.maximum(T0, 0).copy()
    S.sweeps(W, T, 5, order='hals')
    assert relfro(W, g['W']) < 1e-9 and relfro(T, g['T']) < 1e-9


def test_scheme_fix_T_matches_oracle():
    X, W0, T0, M = orc.synth(80, 60, 4, 5, sigma=0.05, seed=3, mask_density=0.25)
    o = orc.nmf_oracle(X, 5, W0, T0, max_iter=4, W_mat=M, fix_T=True)
    I, J = M.nonzero()
    S = SparseWRRI(I, J, X[I, J], 80, 60)
    W, T = np.maximum(W0, 0).copy(), np.maximum(T0, 0).copy()
    S.sweeps(W, T, 4, fix_T=True)
    assert relfro(W, o['W']) < 1e-9 and np.array_equal(T, np.maximum(T0, 0))


@pytest.mark.parametrize('seed', range(12))
def test_scheme_random_structures_match_oracle(seed):
    """ragged structures (empty rows / columns, single-entry segments), k = 1, entry weights, both orders, bounds and
    regularisers drawn at random: the scheme equals the oracle's masked iteration on the densified data"""
    rs = np.random.RandomState(100 + seed)
    n, d = int(rs.randint(3, 40)), int(rs.randint(3, 40))
    k = int(rs.choice([1, 2, 5]))
    X, W0, T0, M = orc.synth(n, d, max(1, k), k, sigma=0.1, seed=seed, mask_density=float(rs.choice([0.1, 0.4, 0.9])))
    M[rs.randint(n), :] = 0
    M[:, rs.randint(d)] = 0
    weighted = bool(rs.randint(2))
    Mw = M * (rs.rand(n, d) + 0.5) if weighted else M
    order = str(rs.choice(['rri', 'hals']))
    kw = dict(reg_w_l1=float(rs.choice([0, 0.05])), reg_t_l1=float(rs.choice([0, 0.05])),
              reg_w_l2=float(rs.choice([0, 0.1])), reg_t_l2=float(rs.choice([0, 0.1])))
    ub_t = float(rs.choice([0, 1.0])) or None
    # with no l2 term an unobserved row/column has a zero denominator: the vector-c branch returns 0 there
    try:
        o = orc.nmf_oracle(X, k, W0, T0, max_iter=3, W_mat=Mw, order=order, t_row_sum=ub_t, **kw)
    except AssertionError:
        pytest.skip('a W column collapses on this draw: the reference itself asserts (nmf.py:476)')
    I, J = M.nonzero()
    S = SparseWRRI(I, J, X[I, J], n, d, weights=Mw[I, J] if weighted else None)
    W, T = np.maximum(W0, 0).copy(), np.maximum(T0, 0).copy()
    S.sweeps(W, T, 3, order=order, ub_t=ub_t, check_copies=(order == 'rri'), **kw)
    assert relfro(W, o['W']) < 1e-9 and relfro(T, o['T']) < 1e-9
