"""CPU oracle: NumPy restatement of the reference's RRI / WRRI sweep arithmetic.

TEST INFRASTRUCTURE ONLY.  Imported by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- never by the product package
`rri_nmf_b200/` (which fails loudly when its CUDA library is missing; there is no CPU fallback).

Parity status: PINNED.  `oracle/make_golden.py` runs the UNMODIFIED reference (through
`oracle/refshim.py`) in the build container and stores its outputs under `tests/golden/`;
`tests/test_oracle.py` checks this restatement against every one of those vectors (and, when
/root/reference is present, against the live reference on random shapes).

Every function cites the reference lines it restates (paths relative to /root/reference).
Notation follows the reference: X[n,d] documents x features, W[n,k] doc-topic, T[k,d] topic-feature,
W_mat (here `M`) optional elementwise weights.
"""
import numpy as np

# src/rri_nmf/nmf.py:52 and src/rri_nmf/optimization.py:5 -- added to every denominator
EPS_DIV_BY_ZERO = float(np.spacing(10))          # 1.7763568394002505e-15
ZERO_TOPIC_TOL = 1e-10                           # nmf.py:758, :794


# ----------------------------------------------------------------------------------------------
# projected 1-D solves  (src/rri_nmf/optimization.py:12-88)
# ----------------------------------------------------------------------------------------------
def euclidean_proj_simplex(v, s=1.0):
    """Duchi et al. sort-based projection onto {w>=0, sum w = s}.  matrixops.py:5-69."""
    v = np.asarray(v)
    shape = v.shape
    v = v.reshape(-1)
    n = v.size
    if v.sum() == s and np.all(v >= 0):          # matrixops.py:53-55
        return v.reshape(shape)
    u = np.sort(v)[::-1]                          # matrixops.py:58
    cssv = np.cumsum(u)                           # :59
    rho = np.nonzero(u * np.arange(1, n + 1) > (cssv - s))[0][-1]   # :61
    theta = (cssv[rho] - s) / (rho + 1.0)         # :63
    return (v - theta).clip(min=0).reshape(shape)  # :65


def proj_mat_to_simplex(Wm, s=1.0):
    """Row-wise simplex projection, in place.  matrixops.py:72-100 (axis=1)."""
    if np.isscalar(s):
        for i in range(Wm.shape[0]):
            Wm[i, :] = euclidean_proj_simplex(Wm[i, :], s)
    else:
        s = np.asarray(s).reshape(-1)
        for i in range(Wm.shape[0]):
            Wm[i, :] = euclidean_proj_simplex(Wm[i, :], s[i])
    return Wm


def qf_min(w, c, s=None, ub=None):
    """min w'x + 0.5 x'diag(c)x  s.t. 0<=x(<=ub), (sum x = s).  optimization.py:12-88.

    Returns (x, nx) with nx = sum(x) before the optional scaling/projection to sum s.
    Branches restated: scalar c>0 (:51-59), scalar c<=0 (:60-74), vector c (:75-87).
    """
    d = w.size
    if s:                                         # :43-49
        if ub:
            ub = min(ub, s)
            assert d * ub >= s
        else:
            ub = s
    if np.isscalar(c) or np.ndim(c) == 0:
        c = float(c)
        if c > 0:                                 # :53-59  (note: `ub` is IGNORED on this branch)
            x = np.maximum(-w, 0) / (c + EPS_DIV_BY_ZERO)
            nx = x.sum()
            if s is not None:
                x = euclidean_proj_simplex(x, s)
        else:                                     # :60-74
            x = np.zeros_like(w)
            if s is None:
                I = np.argwhere(w + c < 0)
                if ub:
                    x[I] = ub
                else:
                    raise ValueError('Minimum objective is unbounded.')    # :105-107
            elif s == 1.0:
                x[np.argmin(w)] = 1.0
            else:
                raise NotImplementedError('s={} is not yet implemented'.format(s))
            nx = 1.0
    else:                                         # :75-87
        if np.any(c < 0) and (s is None and ub is None):
            raise ValueError('Minimum objective is unbounded.')
        I = np.argwhere(c > 0).ravel()
        x = np.zeros_like(w)
        x[I] = np.maximum(-w[I], 0) / (c[I] + EPS_DIV_BY_ZERO)
        if ub is not None:
            x = np.minimum(x, ub)
        nx = x.sum()
        if s is not None:
            x = s * x / x.sum()
    return x, nx


# ----------------------------------------------------------------------------------------------
# sufficient statistics of one half-step  (src/rri_nmf/nmf.py:633-747)
# ----------------------------------------------------------------------------------------------
def update_T_stats(X, W, T, t, M=None, rows=None):
    """(wR, nw) for row T[t,:].  nmf.py:670-676 (unmasked), :687-701 (masked WRRI).

    `rows`: optional row subset -> the partial statistic of nmf.py:680-686 / :706-713 (what one
    row shard of a multi-GPU run contributes; partials over disjoint subsets add up).
    """
    if rows is not None:
        X = X[rows, :]
        W = W[rows, :]
        if M is not None:
            M = M[rows, :]
    w = W[:, t]
    if M is None:
        wX = w.T.dot(X)                           # :672
        wW = w.T.dot(W)                           # :673
        wW[t] = 0                                 # :674
        wR = wX - wW.dot(T)                       # :675
        nw = (w ** 2).sum()                       # :676
    else:
        Wz = W.copy()
        Wz[:, t] = 0                              # :690-693
        Rt = M * (X - Wz.dot(T))                  # :692, :695-698
        wR = w.T.dot(Rt).ravel()                  # :700
        nw = (w ** 2).dot(M).ravel()              # :701
    return wR, nw


def update_W_stats(X, W, T, t, M=None):
    """(Rt, nt) for column W[:,t].  nmf.py:728-734 (unmasked), :735-746 (masked WRRI)."""
    if M is None:
        Xt = X.dot(T[t, :].T)                     # :729
        Tt = T.dot(T[t, :].T)                     # :730
        Tt[t] = 0                                 # :732
        Rt = Xt - W.dot(Tt)                       # :733
        nt = (T[t, :] ** 2).sum()                 # :734
    else:
        Wz = W.copy()
        Wz[:, t] = 0                              # :736-739
        R = M * (X - Wz.dot(T))                   # :738, :740-743
        Rt = R.dot(T[t, :].T).ravel()             # :745
        nt = M.dot(T[t, :] ** 2).ravel()          # :746
    return Rt, nt


def step_T(X, W, T, t, M=None, reg_t_l1=0.0, reg_t_l2=0.0, s=None, ub=None, rows=None):
    """T-step for topic t, in place on T.  nmf.py:420-447.  Returns nt1 = sum of new row."""
    wR, nw = update_T_stats(X, W, T, t, M, rows)
    numer = wR - reg_t_l1                         # :437
    denom = nw + reg_t_l2                         # :438
    T[t, :], nt1 = qf_min(-numer, denom, s=s, ub=ub)    # :447
    return nt1


def step_W(X, W, T, t, M=None, reg_w_l1=0.0, reg_w_l2=0.0, ub=None):
    """W-step for topic t, in place on W.  nmf.py:462-469.  Returns nw1 = sum of new column."""
    Rt, nt = update_W_stats(X, W, T, t, M)
    numer = Rt - reg_w_l1                         # :464
    denom = nt + reg_w_l2                         # :465
    W[:, t], nw1 = qf_min(-numer, denom, s=None, ub=ub)  # :469
    return nw1


# ----------------------------------------------------------------------------------------------
# objective  (src/rri_nmf/nmf.py:71-94)
# ----------------------------------------------------------------------------------------------
def objective(X, W, T, M=None, reg_w_l1=0.0, reg_w_l2=0.0, reg_t_l1=0.0, reg_t_l2=0.0):
    R = (X - np.dot(W, T)) ** 2                   # :77
    if M is not None:
        R = M * R                                 # :78-79
    return (0.5 * np.sum(R) + 0.5 * reg_w_l2 * np.sum(W ** 2) + 0.5 * reg_t_l2 * np.sum(T ** 2)
            + reg_t_l1 * np.sum(np.abs(T)) + reg_w_l1 * np.sum(np.abs(W)))   # :83-91


def rel_error(X, W, T, M=None):
    """||M^(1/2) o (X - WT)||_F / ||M^(1/2) o X||_F  -- the FP32-mode acceptance figure
    (BASELINE.json north_star; the reference itself only compares Frobenius norms,
    tests/test_nmf.py:96)."""
    R = (X - np.dot(W, T)) ** 2
    X2 = X ** 2
    if M is not None:
        R = M * R
        X2 = M * X2
    return float(np.sqrt(np.sum(R, dtype=np.float64) / np.sum(X2, dtype=np.float64)))


def universal_stopping_condition(obj_history, eps_stop=1e-4):
    """optimization.py:284-291."""
    if len(obj_history) < 2:
        return False
    d1 = abs(obj_history[0] - obj_history[1])
    de = abs(obj_history[-1] - obj_history[-2])
    return de <= eps_stop * d1


# ----------------------------------------------------------------------------------------------
# sweeps
# ----------------------------------------------------------------------------------------------
class ZeroTopic(Exception):
    """A row of T / column of W summed to <= 1e-10 (nmf.py:757-758, :793-794) while resets were
    requested; the restatement does not implement the reset action (out of scope)."""


def sweep(X, W, T, M=None, order='rri', fix_W=False, fix_T=False,
          reg_w_l1=0.0, reg_w_l2=0.0, reg_t_l1=0.0, reg_t_l2=0.0,
          t_row_sum=None, w_row_sum=None, project_T_each_iter=False, check_zero=False):
    """One sweep over all k topics, in place on W and T.

    order='rri' : the reference's interleaved order, nmf.py:415-476 -- for each t: T-step then W-step.
    order='hals': block order -- all T-steps (W frozen), then all W-steps (T frozen); this is the
                  reference's own _compute_update_T/_compute_update_W/qf_min driven in block order
                  (SURVEY.md F2/F5), NOT nmf(fix_W=True).
    Returns (sum_T[k], sum_W[k]) -- the per-topic sums the reference uses for zero-topic detection.
    """
    k = W.shape[1]
    sT = np.full(k, np.nan)
    sW = np.full(k, np.nan)
    s = t_row_sum if project_T_each_iter else None            # :442-445
    noreg = (abs(reg_w_l1) + abs(reg_w_l2) + abs(reg_t_l1) + abs(reg_t_l2)) == 0

    def do_T(t):
        nt1 = step_T(X, W, T, t, M, reg_t_l1, reg_t_l2, s=s, ub=t_row_sum)
        if noreg and order == 'rri':
            # nmf.py:450-452.  Harmless when the W-step follows (column t is overwritten without
            # being read, :732); only observable with fix_W=True.
            W[:, t] = W[:, t] * nt1
        sT[t] = np.sum(T[t, :])                                # :757
        if check_zero and sT[t] <= ZERO_TOPIC_TOL:
            raise ZeroTopic('T', t)
        if t_row_sum and project_T_each_iter and abs(sT[t] - t_row_sum) > 1e-15:   # :759-761
            T[t, :] = euclidean_proj_simplex(T[t, :], s=t_row_sum)

    def do_W(t):
        step_W(X, W, T, t, M, reg_w_l1, reg_w_l2, ub=w_row_sum)
        sW[t] = np.sum(W[:, t])                                # :793
        if check_zero and sW[t] <= ZERO_TOPIC_TOL:
            raise ZeroTopic('W', t)
        assert np.all(W[:, t] >= 0), 'W contains negative entries'     # :475
        assert np.sum(W[:, t]) > 0, 'W[:, t] sums to 0'                 # :476

    if order == 'rri':
        for t in range(k):                                     # :415
            if not fix_T:
                do_T(t)
            if not fix_W:
                do_W(t)
    elif order == 'hals':
        if not fix_T:
            for t in range(k):
                do_T(t)
        if not fix_W:
            for t in range(k):
                do_W(t)
    else:
        raise ValueError(order)
    return sT, sW


def nmf_oracle(X, k, W_in, T_in, max_iter=200, W_mat=None, order='rri', fix_W=False, fix_T=False,
               compute_obj_each_iter=False, eps_stop=1e-4,
               reg_w_l1=0.0, reg_w_l2=0.0, reg_t_l1=0.0, reg_t_l2=0.0,
               t_row_sum=None, w_row_sum=None, project_T_each_iter=False,
               project_W_each_iter=False, do_final_project_W=True, early_stop=None,
               snapshots=None):
    """Restatement of the sweep driver nmf.py:351-560 for explicit W_in/T_in, resets disabled
    (reset_topic_method=None), no w_row, no DP noise.  `snapshots`: optional iterable of sweep
    counts at which (W,T) copies are recorded -> returned under 'snapshots'."""
    n, d = X.shape
    if np.shape(W_in) != (n, k):
        raise ValueError('W_in has wrong dimensions, must be n*k')     # :853-854
    if np.shape(T_in) != (k, d):
        raise ValueError('T_in has wrong dimensions, must be k*d')     # :858-859
    if project_T_each_iter and np.any([reg_w_l1, reg_t_l1]):            # :280-285
        project_T_each_iter = False
    W = np.maximum(W_in, 0)                                             # :867
    T = np.maximum(T_in, 0)                                             # :868
    if project_W_each_iter and not fix_W and w_row_sum is not None:     # :870-873
        W = proj_mat_to_simplex(W, w_row_sum)
    if project_T_each_iter and not fix_T and t_row_sum is not None:     # :875-878
        T = proj_mat_to_simplex(T, t_row_sum)
    regs = dict(reg_w_l1=reg_w_l1, reg_w_l2=reg_w_l2, reg_t_l1=reg_t_l1, reg_t_l2=reg_t_l2)
    obj_history = []
    snaps = {}
    if early_stop:
        last_score = np.inf
        W_prev, T_prev = W.copy(), T.copy()
    for it in range(max_iter):                                          # :377
        if early_stop:                                                  # :381-407
            this_score = early_stop(X, W, T)
            if this_score > last_score:
                W, T = W_prev, T_prev
                obj_history = obj_history[:-1]
                break
            last_score = this_score
            W_prev, T_prev = W.copy(), T.copy()
        sweep(X, W, T, W_mat, order, fix_W, fix_T, t_row_sum=t_row_sum, w_row_sum=w_row_sum,
              project_T_each_iter=project_T_each_iter, **regs)
        if project_W_each_iter and not fix_W and w_row_sum is not None:  # :481-484
            W = proj_mat_to_simplex(W, w_row_sum)
        if compute_obj_each_iter:                                       # :488-489
            obj_history.append(objective(X, W, T, W_mat, **regs))
        if snapshots is not None and (it + 1) in snapshots:
            snaps[it + 1] = (W.copy(), T.copy())
        if compute_obj_each_iter and universal_stopping_condition(obj_history, eps_stop):  # :510
            break
    if (not project_W_each_iter and w_row_sum is not None and not fix_W
            and do_final_project_W):                                    # :519-529
        W = proj_mat_to_simplex(W, w_row_sum)
    out = {'W': W, 'T': T}
    if compute_obj_each_iter:
        out['obj_history'] = obj_history
    if snapshots is not None:
        out['snapshots'] = snaps
    return out


# ----------------------------------------------------------------------------------------------
# synthetic workloads -- one generator for the CPU and the GPU side
# ----------------------------------------------------------------------------------------------
def synth(n, d, r, k, sigma=0.0, seed=0, dtype=np.float64, mask_density=None, mask_seed=7):
    """X = U V + sigma*mean(UV)*E with U,V,E ~ U[0,1); W0,T0 ~ U[0,1).  Returns X, W0, T0[, M]."""
    rs = np.random.RandomState(seed)
    U = rs.rand(n, r)
    V = rs.rand(r, d)
    X = U.dot(V)
    if sigma:
        X = X + sigma * X.mean() * rs.rand(n, d)
    W0 = rs.rand(n, k)
    T0 = rs.rand(k, d)
    out = [X.astype(dtype), W0.astype(dtype), T0.astype(dtype)]
    if mask_density is not None:
        M = (np.random.RandomState(mask_seed).rand(n, d) < mask_density).astype(dtype)
        out.append(M)
    return tuple(out)
