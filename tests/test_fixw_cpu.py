"""fix_W=True (T-only sweeps with the reference's W[:, t] *= nt1 rescaling, nmf.py:450-452): the oracle is pinned
to the unmodified reference (tests/golden/fixW_f64.npz), and the product's host-side sweep (`nmf._sweep_fix_W`:
rri_partials_T + device vector ops) is run here on CPU tensors with the engine's two calls served by the oracle."""
import numpy as np
import pytest
import torch

import rri_oracle as orc
from conftest import golden, relfro
from rri_nmf_b200.nmf import _qf_min_device, _sweep_fix_W

CASES = [('plain', {}), ('simplex', dict(project_T_each_iter=True, t_row_sum=1.0)),
         ('reg', dict(reg_t_l1=0.01, reg_t_l2=0.05)), ('masked', dict(masked=True)),
         ('masked_ub', dict(masked=True, t_row_sum=1.0))]


def inputs():
    return orc.synth(120, 80, 5, 6, sigma=0.05, seed=13, mask_density=0.4)


@pytest.mark.parametrize('name,kw', CASES)
def test_oracle_fix_W_matches_reference(name, kw):
    X, W0, T0, M = inputs()
    kw = dict(kw)
    Wm = M if kw.pop('masked', False) else None
    o = orc.nmf_oracle(X, 6, W0, T0, max_iter=3, fix_W=True, W_mat=Wm, **kw)
    g = golden('fixW_f64.npz')
    assert relfro(o['W'], g['W_' + name]) < 1e-12 and relfro(o['T'], g['T_' + name]) < 1e-12


class OracleBackedEngine(object):
    """stands in for RRIEngine: the two C-ABI calls the fix_W sweep makes, answered by the oracle"""

    def __init__(self, X, M):
        self.X, self.M = X, M

    def partials_T(self, W, T, t):
        wR, nw = orc.update_T_stats(self.X, W.numpy(), T.numpy(), t, self.M)
        return torch.from_numpy(np.asarray(wR)), torch.from_numpy(np.atleast_1d(np.asarray(nw, dtype=np.float64)))

    def project_rows_simplex(self, A, s=1.0):
        A.copy_(torch.from_numpy(orc.proj_mat_to_simplex(A.numpy(), s)))
        return A


@pytest.mark.parametrize('name,kw', CASES)
def test_product_fix_W_sweep_matches_reference(name, kw):
    X, W0, T0, M = inputs()
    kw = dict(kw)
    Wm = M if kw.pop('masked', False) else None
    a = dict(k=6, t_row_sum=None, project_T_each_iter=False, reg_w_l1=0, reg_w_l2=0, reg_t_l1=0, reg_t_l2=0)
    a.update(kw)
    eng = OracleBackedEngine(X, Wm)
    W, T = torch.from_numpy(np.maximum(W0, 0).copy()), torch.from_numpy(np.maximum(T0, 0).copy())
    if a['project_T_each_iter']:
        eng.project_rows_simplex(T, a['t_row_sum'])          # nmf.py:875-878 (done by nmf._solve before the sweeps)
    for _ in range(3):
        _sweep_fix_W(eng, W, T, a)
    g = golden('fixW_f64.npz')
    assert relfro(W.numpy(), g['W_' + name]) < 1e-9 and relfro(T.numpy(), g['T_' + name]) < 1e-9


def test_qf_min_device_branches_match_oracle():
    rs = np.random.RandomState(0)
    w = rs.randn(40)
    eng = OracleBackedEngine(None, None)
    for c, s, ub in [(2.5, None, None), (2.5, 1.0, None), (2.5, None, 0.3), (-1.0, None, 0.7), (0.0, 1.0, None),
                     (rs.rand(40) + 0.1, None, None), (rs.rand(40) - 0.3, None, 0.5), (rs.rand(40) + 0.1, 2.0, None)]:
        xo, no = orc.qf_min(w, c, s=s, ub=ub)
        cd = torch.from_numpy(np.atleast_1d(np.asarray(c, dtype=np.float64)))
        xd, nd = _qf_min_device(torch.from_numpy(-w), cd, s, ub, eng)
        assert np.allclose(xd.numpy(), xo, atol=1e-14) and abs(nd - no) < 1e-12
    with pytest.raises(ValueError):
        _qf_min_device(torch.from_numpy(-w), torch.tensor([-1.0], dtype=torch.float64), None, None, eng)
    with pytest.raises(ValueError):
        _qf_min_device(torch.from_numpy(-w), torch.from_numpy(rs.rand(40) - 0.5), None, None, eng)
