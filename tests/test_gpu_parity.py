"""GPU parity tests: the CUDA sweep (through the C-ABI of include/rri_b200.h, via rri_nmf_b200.nmf /
RRIEngine) against the golden vectors of the unmodified reference and against the NumPy oracle on
the same seeded inputs.

Tolerances (BASELINE.json north_star): FP64 -- relative Frobenius <= 1e-9 per compared state;
FP32 -- final relative reconstruction error within 1e-4 of the reference after the same sweeps.
"""
import numpy as np
import pytest
import torch

import rri_oracle as orc
from conftest import golden, relfro

pytestmark = pytest.mark.gpu

F64_TOL = 1e-9
F32_RELERR_TOL = 1e-4


@pytest.fixture(scope='module')
def R(cuda_device):
    import rri_nmf_b200
    return rri_nmf_b200


def run(R, X, k, W0, T0, **kw):
    kw.setdefault('reset_topic_method', None)
    kw.setdefault('max_time', 1e9)
    kw.setdefault('eps_stop', -1.0)
    return R.nmf(X, k, W_in=W0, T_in=T0, **kw)


# ------------------------------------------------------------------------------------------- cfg1
@pytest.mark.parametrize('order,gname', [('rri', 'cfg1_rri_f64.npz'), ('hals', 'cfg1_hals_f64.npz')])
def test_cfg1_fp64_snapshots(R, order, gname):
    g = golden(gname)
    X, W0, T0 = orc.synth(500, 300, 10, 10, sigma=0.0, seed=0)
    W, T, done = W0, T0, 0
    for c in [int(c) for c in g['counts']]:
        out = run(R, X, 10, W, T, max_iter=c - done, update_order=order)
        W, T, done = out['W'], out['T'], c
        assert relfro(W, g['W_%d' % c]) < F64_TOL, (order, c)
        assert relfro(T, g['T_%d' % c]) < F64_TOL, (order, c)


def test_cfg1_fp64_objective_history(R):
    g = golden('cfg1_rri_f64.npz')
    X, W0, T0 = orc.synth(500, 300, 10, 10, sigma=0.0, seed=0)
    out = run(R, X, 10, W0, T0, max_iter=30, compute_obj_each_iter=True)
    assert np.allclose(out['obj_history'], g['obj_history'][:30], rtol=1e-9)
    assert np.all(np.diff(out['obj_history']) <= 0)          # reference tests/test_nmf.py:40
    assert not np.shares_memory(out['W'], W0)                 # inputs are not mutated (nmf.py:867-868)
    assert abs(out['obj_calculator'].true_objective() - out['obj_history'][-1]) < 1e-9 * out['obj_history'][-1]


def test_cfg1_fp32_rri(R):
    g = golden('cfg1_rri_f32.npz')
    X, W0, T0 = orc.synth(500, 300, 10, 10, sigma=0.0, seed=0, dtype=np.float32)
    out = run(R, X, 10, W0, T0, max_iter=10)
    assert out['W'].dtype == np.float32                       # dtype follows the inputs
    re = orc.rel_error(X.astype(np.float64), out['W'].astype(np.float64), out['T'].astype(np.float64))
    assert abs(re - float(g['relerr_10'])) < F32_RELERR_TOL
    assert relfro(out['W'], g['W_10']) < 1e-3


@pytest.mark.parametrize('math', ['ieee', 'tf32'])
def test_cfg1_fp32_hals(R, math):
    g = golden('cfg1_hals_f64.npz')
    X, W0, T0 = orc.synth(500, 300, 10, 10, sigma=0.0, seed=0)
    ref_re = orc.rel_error(X, g['W_10'], g['T_10'])
    out = run(R, X.astype(np.float32), 10, W0.astype(np.float32), T0.astype(np.float32), max_iter=10,
              update_order='hals', math=math)
    re = orc.rel_error(X, out['W'].astype(np.float64), out['T'].astype(np.float64))
    assert abs(re - ref_re) < F32_RELERR_TOL, (math, re, ref_re)


# -------------------------------------------------------------------------- regularised / partials
def test_regularised_fp64(R):
    g = golden('reg_rri_f64.npz')
    X, W0, T0 = orc.synth(257, 131, 6, 7, sigma=0.05, seed=3)
    rw1, rw2, rt1, rt2 = g['regs']
    out = run(R, X, 7, W0, T0, max_iter=6, compute_obj_each_iter=True, reg_w_l1=rw1, reg_w_l2=rw2,
              reg_t_l1=rt1, reg_t_l2=rt2)
    assert relfro(out['W'], g['W']) < F64_TOL and relfro(out['T'], g['T']) < F64_TOL
    assert np.allclose(out['obj_history'], g['obj_history'], rtol=1e-9)


def test_row_shard_partials(R, cuda_device):
    """rri_partials_T == nmf.py:680-686 on each shard; shards add up to the full statistic."""
    g = golden('partials_f64.npz')
    X, W0, T0 = orc.synth(257, 131, 6, 7, sigma=0.05, seed=3)
    s = int(g['split'][0])
    dev = cuda_device
    Td = torch.from_numpy(T0).to(dev)
    for t in (0, 3, 6):
        tot_w, tot_n = 0, 0
        for nm, sl in (('a', slice(0, s)), ('b', slice(s, 257))):
            eng = R.RRIEngine(torch.from_numpy(X[sl]).to(dev), 7, order='rri')
            wR, nw = eng.partials_T(torch.from_numpy(W0[sl]).to(dev).contiguous(), Td, t)
            assert relfro(wR.cpu().numpy(), g['numer_%s_%d' % (nm, t)]) < 1e-12
            assert abs(float(nw[0]) - float(g['denom_%s_%d' % (nm, t)])) < 1e-11
            tot_w, tot_n = tot_w + wR.cpu().numpy(), tot_n + float(nw[0])
            eng.close()
        assert relfro(tot_w, g['numer_full_%d' % t]) < 1e-12
        assert abs(tot_n - float(g['denom_full_%d' % t])) < 1e-11


@pytest.mark.parametrize('order', ['rri', 'hals'])
def test_fix_T_transform_path(R, order):
    g = golden('fixT_f64.npz')
    X, W0, T0 = orc.synth(257, 131, 6, 7, sigma=0.05, seed=3)
    out = run(R, X, 7, W0, T0, max_iter=4, fix_T=True, update_order=order)
    assert relfro(out['W'], g['W']) < F64_TOL
    assert np.array_equal(out['T'], g['T'])


# ------------------------------------------------------------------------------------ masked WRRI
@pytest.mark.parametrize('name,regs', [('plain', {}), ('l1both', {'reg_w_l1': 0.1, 'reg_t_l1': 0.1}),
                                        ('l1w', {'reg_w_l1': 0.1}), ('l1t', {'reg_t_l1': 0.1})])
@pytest.mark.parametrize('mask_kind', ['real', 'u8'])
def test_recsys_masked_wrri(R, name, regs, mask_kind):
    g = golden('recsys_wrri_f64.npz')
    X = g['X']
    M = (X != 0)
    Wm = M.astype(np.float64) if mask_kind == 'real' else torch.from_numpy(M.astype(np.uint8))
    out = run(R, X, 7, g['W0'], g['T0'], max_iter=15, W_mat=Wm, compute_obj_each_iter=True, t_row_sum=1.0,
              **regs)
    assert relfro(out['W'], g['W_' + name]) < F64_TOL and relfro(out['T'], g['T_' + name]) < F64_TOL
    assert np.allclose(out['obj_history'], g['obj_' + name], rtol=1e-9)
    assert np.all(np.diff(out['obj_history']) <= 1e-9)        # reference tests/test_nmf.py:78
    assert out['T'].max() <= 1.0                              # ub clip (optimization.py:82-83)


def test_weighted_wrri_real_weights_both_orders(R):
    g = golden('weighted_wrri_f64.npz')
    X, W0, T0, Mb = orc.synth(120, 90, 5, 6, sigma=0.05, seed=5, mask_density=0.3)
    Mw = Mb * np.random.RandomState(9).rand(120, 90) * 2.0
    out = run(R, X, 6, W0, T0, max_iter=8, W_mat=Mw, compute_obj_each_iter=True)
    assert relfro(out['W'], g['W']) < F64_TOL and relfro(out['T'], g['T']) < F64_TOL
    assert np.allclose(out['obj_history'], g['obj_history'], rtol=1e-9)
    gh = golden('weighted_wrri_hals_f64.npz')
    out = run(R, X, 6, W0, T0, max_iter=5, W_mat=Mw, update_order='hals')
    assert relfro(out['W'], gh['W']) < F64_TOL and relfro(out['T'], gh['T']) < F64_TOL


def test_masked_fp32_relerr(R):
    X, W0, T0, M = orc.synth(300, 200, 8, 8, sigma=0.05, seed=21, mask_density=0.2)
    o = orc.nmf_oracle(X, 8, W0, T0, max_iter=10, W_mat=M)
    out = run(R, X.astype(np.float32), 8, W0.astype(np.float32), T0.astype(np.float32), max_iter=10,
              W_mat=M.astype(np.float32))
    re_o = orc.rel_error(X, o['W'], o['T'], M)
    re_g = orc.rel_error(X, out['W'].astype(np.float64), out['T'].astype(np.float64), M)
    assert abs(re_o - re_g) < F32_RELERR_TOL


@pytest.mark.parametrize('order', ['rri', 'hals'])
@pytest.mark.parametrize('mask_kind', ['u8', 'real'])
def test_masked_tensor_core_path(R, order, mask_kind):
    """fp32 + math='tf32' with a mask: the W T tiles come from tcgen05; same 1e-4 criterion, and the
    TF32 result stays close to the IEEE fp32 kernels on the same inputs."""
    n, d, k = 700, 520, 13
    X, W0, T0, M = orc.synth(n, d, k, k, sigma=0.05, seed=23, mask_density=0.15)
    if mask_kind == 'real':
        Mw = M * (0.5 + np.random.RandomState(3).rand(n, d))
        Mg = Mw.astype(np.float32)
    else:
        Mw = M
        Mg = torch.from_numpy(M.astype(np.uint8))
    o = orc.nmf_oracle(X, k, W0, T0, max_iter=6, W_mat=Mw, order=order, t_row_sum=1.0)
    kw = dict(max_iter=6, W_mat=Mg, update_order=order, t_row_sum=1.0)
    g = run(R, X.astype(np.float32), k, W0.astype(np.float32), T0.astype(np.float32), math='tf32', **kw)
    gi = run(R, X.astype(np.float32), k, W0.astype(np.float32), T0.astype(np.float32), math='ieee', **kw)
    re_o = orc.rel_error(X, o['W'], o['T'], Mw)
    re_g = orc.rel_error(X, g['W'].astype(np.float64), g['T'].astype(np.float64), Mw)
    assert abs(re_o - re_g) < F32_RELERR_TOL, (re_o, re_g)
    assert relfro(g['W'], gi['W']) < 2e-2 and relfro(g['T'], gi['T']) < 2e-2
    assert g['T'].max() <= 1.0


# ---------------------------------------------------------------------- topic-model setting (f1)
def test_text_topic_model_simplex(R):
    g = golden('text_tm_f64.npz')
    X = g['X']
    out = run(R, X, 15, g['W0'], g['T0'], max_iter=15, w_row_sum=1.0, project_T_each_iter=True,
              project_W_each_iter=True, compute_obj_each_iter=True, t_row_sum=1.0, reg_t_l2=0.1)
    assert relfro(out['W'], g['W_tm']) < F64_TOL and relfro(out['T'], g['T_tm']) < F64_TOL
    assert np.allclose(out['obj_history'], g['obj_tm'], rtol=1e-9)
    assert np.all(np.diff(out['obj_history']) <= 0)
    cv = np.sum(np.abs(out['W'].sum(1) - 1)) + np.sum(np.abs(out['T'].sum(1) - 1))
    assert cv <= 1e-12                                         # reference tolerance 1e-13 per test_nmf.py:10
    out = run(R, X, 15, g['W0'], g['T0'], max_iter=10, w_row_sum=1.0, project_T_each_iter=True,
              t_row_sum=1.0)
    assert relfro(out['W'], g['W_est']) < F64_TOL and relfro(out['T'], g['T_est']) < F64_TOL


def test_simplex_projection_rows(R, cuda_device):
    rs = np.random.RandomState(5)
    for rows, cols, s in ((7, 5, 1.0), (64, 300, 2.5), (3, 20000, 1.0), (1000, 64, 1.0)):
        A = rs.randn(rows, cols)
        A[0, :] = 0.0                                          # all-zero row -> uniform
        ref = np.stack([orc.euclidean_proj_simplex(A[i], s) for i in range(rows)])
        eng = R.RRIEngine(torch.zeros(4, 4, dtype=torch.float64, device=cuda_device), 2)
        got = eng.project_rows_simplex(torch.from_numpy(A).to(cuda_device), s).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-12
        assert np.abs(got.sum(1) - s).max() < 1e-12


# -------------------------------------------------------------------------- determinism / resume
@pytest.mark.parametrize('order', ['rri', 'hals'])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_reentrant_bit_exact(R, order, dtype):
    """N sweeps == N x 1 sweep, bit for bit (reference tests/test_nmf.py:97-109 asserts allclose)."""
    X, W0, T0 = orc.synth(403, 257, 9, 9, sigma=0.05, seed=8, dtype=dtype)
    a = run(R, X, 9, W0, T0, max_iter=5, update_order=order)
    W, T = W0, T0
    for _ in range(5):
        o = run(R, X, 9, W, T, max_iter=1, update_order=order)
        W, T = o['W'], o['T']
    assert np.array_equal(a['W'], W) and np.array_equal(a['T'], T)
    b = run(R, X, 9, W0, T0, max_iter=5, update_order=order)
    assert np.array_equal(a['W'], b['W']) and np.array_equal(a['T'], b['T'])     # run-to-run


# ------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize('n,d,k', [(1, 7, 1), (5, 3, 2), (33, 1, 1), (70, 1030, 3), (129, 65, 33),
                                   (64, 64, 64), (40, 50, 70)])
@pytest.mark.parametrize('order', ['rri', 'hals'])
def test_ragged_shapes_fp64(R, n, d, k, order):
    rs = np.random.RandomState(n * 1000 + d)
    X, W0, T0 = rs.rand(n, d) + 0.1, rs.rand(n, k) + 0.1, rs.rand(k, d) + 0.1
    try:
        o = orc.nmf_oracle(X, k, W0, T0, max_iter=2, order=order)
    except ValueError:
        with pytest.raises(ValueError):
            run(R, X, k, W0, T0, max_iter=2, update_order=order)
        return
    out = run(R, X, k, W0, T0, max_iter=2, update_order=order)
    assert relfro(out['W'], o['W']) < 1e-8 and relfro(out['T'], o['T']) < 1e-8


def test_strided_X_and_torch_io(R, cuda_device):
    rs = np.random.RandomState(4)
    big = torch.from_numpy(rs.rand(90, 77)).to(cuda_device)
    Xv = big[:, :70]                                           # row stride 77 != d
    W0, T0 = rs.rand(90, 4), rs.rand(4, 70)
    o = orc.nmf_oracle(Xv.cpu().numpy(), 4, W0, T0, max_iter=3)
    out = run(R, Xv, 4, torch.from_numpy(W0), torch.from_numpy(T0), max_iter=3)
    assert isinstance(out['W'], torch.Tensor) and out['W'].is_cuda
    assert relfro(out['W'].cpu().numpy(), o['W']) < F64_TOL


def test_zero_topic_raises_like_reference(R):
    rs = np.random.RandomState(33 + 57)
    X, W0, T0 = rs.rand(33, 57), rs.rand(33, 8), rs.rand(8, 57)
    with pytest.raises(ValueError):
        orc.nmf_oracle(X, 8, W0, T0, max_iter=5)
    with pytest.raises(ValueError):
        run(R, X, 8, W0, T0, max_iter=5)


def test_topic_reset_matches_reference(R):
    """default reset_topic_method='max_resid_document': a topic that empties in the first sweep is re-seeded
    from the document with the largest positive residual, exactly where the reference does it
    (nmf.py:762-783); golden = unmodified reference on the same inputs."""
    g = golden('reset_rri_f64.npz')
    rs = np.random.RandomState(33 + 57)
    X, W0, T0 = rs.rand(33, 57), rs.rand(33, 8), rs.rand(8, 57)
    out = R.nmf(X, 8, W_in=W0, T_in=T0, max_iter=6, compute_obj_each_iter=True, eps_stop=-1.0)
    assert np.allclose(out['obj_history'], g['obj_history'], rtol=1e-9)
    assert relfro(out['W'], g['W']) < F64_TOL and relfro(out['T'], g['T']) < F64_TOL


def test_argument_errors(R):
    X = np.random.RandomState(0).rand(20, 10)
    with pytest.raises(ValueError):
        R.nmf(X, 3, W_in=np.ones((19, 3)), T_in=np.ones((3, 10)))         # nmf.py:853-854
    with pytest.raises(ValueError):
        R.nmf(X, 3, W_in=np.ones((20, 3)), T_in=np.ones((3, 11)))         # nmf.py:858-859
    out = R.nmf(X, 3, reg_t_l2=-0.1)                                       # unbounded: nmf.py:292-303
    assert out['obj_history'] == [-np.inf] and np.all(out['T'] == 1e6)


def test_objective_and_relerr(R, cuda_device):
    X, W0, T0, M = orc.synth(150, 130, 5, 6, sigma=0.1, seed=2, mask_density=0.5)
    for Mm in (None, M):
        eng = R.RRIEngine(torch.from_numpy(X).to(cuda_device), 6,
                          W_mat=None if Mm is None else torch.from_numpy(Mm).to(cuda_device))
        Wd, Td = torch.from_numpy(W0).to(cuda_device), torch.from_numpy(T0).to(cuda_device)
        regs = dict(reg_w_l1=0.1, reg_w_l2=0.2, reg_t_l1=0.3, reg_t_l2=0.4)
        assert abs(eng.objective(Wd, Td, **regs) / orc.objective(X, W0, T0, Mm, **regs) - 1) < 1e-12
        assert abs(eng.rel_error(Wd, Td) / orc.rel_error(X, W0, T0, Mm) - 1) < 1e-12
        eng.close()


# ----------------------------------------------------------------------------- config-2 scale
@pytest.mark.parametrize('order', ['rri', 'hals'])
def test_cfg2_scale_fp64_every_sweep(R, order):
    """20k x 5k, k=32, fp64 (BASELINE.json configs[1]), both update orders; 10 sweeps, compared with the oracle
    after EVERY sweep, the device continuing from its own state."""
    X, W0, T0 = orc.synth(20000, 5000, 32, 32, sigma=0.05, seed=0)
    Wo, To = np.maximum(W0, 0), np.maximum(T0, 0)
    dev = torch.device('cuda:0')
    eng = R.RRIEngine(torch.from_numpy(X).to(dev), 32, order=order)
    W, T = torch.from_numpy(Wo).to(dev), torch.from_numpy(To).to(dev)
    p = eng.params()
    for s in range(10):
        orc.sweep(X, Wo, To, order=order)
        assert eng.sweeps(W, T, 1, p) == 0
        assert relfro(W.cpu().numpy(), Wo) < F64_TOL and relfro(T.cpu().numpy(), To) < F64_TOL, (order, s)
    eng.close()


def test_cfg3_shape_fp32_tf32_relerr(R):
    """config-3 columns (d=20000, k=64) on a row subsample the CPU oracle can hold: FP32 IEEE rri and
    TF32 hals both land within 1e-4 of the oracle's relative error after the same sweeps."""
    n, d, k = 2048, 20000, 64
    X, W0, T0 = orc.synth(n, d, k, k, sigma=0.05, seed=1)
    Xf, Wf, Tf = X.astype(np.float32), W0.astype(np.float32), T0.astype(np.float32)
    for order, math in (('rri', 'ieee'), ('hals', 'ieee'), ('hals', 'tf32')):
        o = orc.nmf_oracle(X, k, W0, T0, max_iter=3, order=order)
        out = run(R, Xf, k, Wf, Tf, max_iter=3, update_order=order, math=math)
        re_o = orc.rel_error(X, o['W'], o['T'])
        re_g = orc.rel_error(X, out['W'].astype(np.float64), out['T'].astype(np.float64))
        assert abs(re_o - re_g) < F32_RELERR_TOL, (order, math, re_o, re_g)


def test_cfg5_shape_k128_relerr(R):
    """config-5 columns and rank (d=20000, k=128) on a row subsample: TF32 block order within 1e-4."""
    n, d, k = 1536, 20000, 128
    X, W0, T0 = orc.synth(n, d, k, k, sigma=0.05, seed=2)
    o = orc.nmf_oracle(X, k, W0, T0, max_iter=2, order='hals')
    out = run(R, X.astype(np.float32), k, W0.astype(np.float32), T0.astype(np.float32), max_iter=2,
              update_order='hals', math='tf32')
    re_o = orc.rel_error(X, o['W'], o['T'])
    re_g = orc.rel_error(X, out['W'].astype(np.float64), out['T'].astype(np.float64))
    assert abs(re_o - re_g) < F32_RELERR_TOL, (re_o, re_g)


def test_cfg4_shape_masked_relerr(R):
    """config-4 columns, rank and mask density (d=20000, k=50, 5 % observed, ub_t=1) on a row subsample:
    the tensor-core masked path and the IEEE fp32 path both land within 1e-4 of the fp64 oracle."""
    n, d, k = 768, 20000, 50
    X, W0, T0, M = orc.synth(n, d, k, k, sigma=0.05, seed=4, mask_density=0.05)
    o = orc.nmf_oracle(X, k, W0, T0, max_iter=1, W_mat=M, t_row_sum=1.0)
    re_o = orc.rel_error(X, o['W'], o['T'], M)
    for math in ('tf32', 'ieee'):
        out = run(R, X.astype(np.float32), k, W0.astype(np.float32), T0.astype(np.float32), max_iter=1,
                  W_mat=torch.from_numpy(M.astype(np.uint8)), t_row_sum=1.0, math=math)
        re_g = orc.rel_error(X, out['W'].astype(np.float64), out['T'].astype(np.float64), M)
        assert abs(re_o - re_g) < F32_RELERR_TOL, (math, re_o, re_g)


def test_gemm_nt_tf32_vs_torch(R, cuda_device):
    torch.manual_seed(0)
    for M, N, K in ((128, 64, 64), (1000, 64, 2000), (4096, 128, 777 * 4), (300, 10, 500), (257, 50, 1028)):
        A = torch.rand(M, K, device=cuda_device)
        B = torch.rand(N, K, device=cuda_device)
        eng = R.RRIEngine(torch.zeros(8, 8, device=cuda_device), N, order='hals', math='tf32')
        Cg = eng.gemm_nt(A, B)
        Cr = (A.double() @ B.double().t())
        rel = float((Cg.double() - Cr).norm() / Cr.norm())
        assert rel < 1e-4, (M, N, K, rel)                      # RN-rounded TF32 operands on [0,1) data: ~1e-5
        eng.close()


@pytest.mark.parametrize('M,N,K', [(2048, 64, 200000), (1024, 128, 125000), (20000, 64, 20000)])
def test_gemm_nt_tf32_long_k_bias(R, cuda_device, M, N, K):
    """Guard of the long-K accumulation: the tensor-core accumulator truncates, and one TMEM accumulation chain
    over K = 200 000 showed a systematic -6.5e-4 relative bias (all products positive) until the kernel began to
    flush TMEM into round-to-nearest fp32 registers every 1024 K.  K here is the row count of config 3 / a config-5
    shard -- the X'W contraction of the T half-step -- which no subsampled parity test reaches."""
    torch.manual_seed(1)
    A = torch.rand(M, K, device=cuda_device)
    B = torch.rand(N, K, device=cuda_device)
    eng = R.RRIEngine(torch.zeros(8, 8, device=cuda_device), N, order='hals', math='tf32')
    Cg = eng.gemm_nt(A, B).double()
    Cr = A.double() @ B.double().t()
    fro = float((Cg - Cr).norm() / Cr.norm())
    bias = float(((Cg - Cr) / Cr).mean())
    eng.close()
    assert abs(bias) < 5e-5, (M, N, K, bias)
    assert fro < 1e-4, (M, N, K, fro)


@pytest.mark.parametrize('order', ['rri', 'hals'])
@pytest.mark.parametrize('form', ['dense', 'sparse'])
def test_masked_sum_constrained_T(R, order, form):
    """qf_min's vector-c branch WITH the sum constraint (optimization.py:75-87: clip to ub, then x <- s x / sum x),
    i.e. nmf(W_mat=..., project_T_each_iter=True, t_row_sum=1): on the device for the dense masked and the
    observed-entries engines, fp64 against the oracle (which equals the unmodified reference bit for bit here)."""
    import scipy.sparse as sp
    X, W0, T0, M = orc.synth(220, 170, 6, 6, sigma=0.05, seed=13, mask_density=0.35)
    kw = dict(max_iter=4, t_row_sum=1.0, project_T_each_iter=True)
    o = orc.nmf_oracle(X, 6, W0, T0, W_mat=M, order=order, **kw)
    if form == 'dense':
        out = run(R, X, 6, W0, T0, W_mat=M, update_order=order, **kw)
    else:
        out = run(R, sp.csr_matrix(X * M), 6, W0, T0, update_order=order, **kw)
    assert relfro(out['W'], o['W']) < F64_TOL and relfro(out['T'], o['T']) < F64_TOL
    assert np.allclose(out['T'].sum(1), 1.0, atol=1e-12)


def test_objective_through_the_contraction(R, cuda_device):
    """nmf.py:71-94 evaluated as ||X||^2 - 2<X T', W> + <W'W, T T'> from the block-order sweep's own contraction (no
    pass over X per evaluation): fp64 agrees with the oracle's explicit objective to 1e-9 relative per sweep, IEEE
    fp32 with the explicit device pass to 1e-4; with TF32 operands the difference of large terms keeps ~1e-3."""
    X, W0, T0 = orc.synth(1500, 900, 12, 12, sigma=0.05, seed=31)
    o = orc.nmf_oracle(X, 12, W0, T0, max_iter=6, order='hals', compute_obj_each_iter=True, eps_stop=-1.0)
    g = run(R, X, 12, W0, T0, max_iter=6, update_order='hals', compute_obj_each_iter=True)       # 'auto' -> contraction
    assert g['obj_calculator'].via_contraction
    assert np.allclose(g['obj_history'], o['obj_history'], rtol=1e-9)
    e = run(R, X, 12, W0, T0, max_iter=6, update_order='hals', compute_obj_each_iter=True, objective='exact')
    assert not e['obj_calculator'].via_contraction and np.allclose(e['obj_history'], o['obj_history'], rtol=1e-10)
    # engine level, fp32: explicit pass vs contraction form, with and without reuse of the sweep's contraction
    Xf = torch.from_numpy(X.astype(np.float32)).to(cuda_device)
    for math, tol in (('ieee', 1e-4), ('tf32', 2e-2)):
        eng = R.RRIEngine(Xf, 12, order='hals', math=math)
        W = torch.from_numpy(W0.astype(np.float32)).to(cuda_device)
        T = torch.from_numpy(T0.astype(np.float32)).to(cuda_device)
        eng.sweeps(W, T, 3, eng.params())
        exact = eng.objective(W, T)
        reuse = eng.objective(W, T, via_contraction=True, reuse_last_sweep=True)
        fresh = eng.objective(W, T, via_contraction=True)
        assert abs(reuse / exact - 1) < tol and abs(fresh / exact - 1) < tol, (math, exact, reuse, fresh)
        assert abs(reuse / fresh - 1) < 1e-6 * (1 if math == 'ieee' else 1e3)
        eng.close()
    with pytest.raises(ValueError):
        run(R, X, 12, W0, T0, max_iter=1, update_order='rri', compute_obj_each_iter=True, objective='contraction')


def test_pageable_host_input_is_staged_exactly(R, cuda_device):
    """Plain (pageable) NumPy input above 64 MB goes to the device through the threaded pinned-chunk staging of
    nmf._pageable_to_device (53 GB/s against 11 GB/s for tensor.to()): the bytes must arrive unchanged, also when the
    size is not a multiple of the chunk and for float64."""
    import importlib
    N = importlib.import_module('rri_nmf_b200.nmf')
    rs = np.random.RandomState(5)
    for shape, dt in (((9001, 2503), np.float32), ((4099, 2053), np.float64)):
        a = rs.rand(*shape).astype(dt)
        assert a.nbytes >= (64 << 20)
        t = N._to_device(a, cuda_device, torch.float32 if dt == np.float32 else torch.float64)
        torch.cuda.synchronize()
        assert t.shape == shape and bool((t.cpu() == torch.from_numpy(a)).all())
