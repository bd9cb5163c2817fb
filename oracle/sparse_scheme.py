"""TEST INFRASTRUCTURE ONLY -- NumPy model of the observed-entries WRRI scheme of
rri_nmf_b200/csrc/sparse_kernels.cu + api.cu (sp_T_step / sp_W_step / sp_sweeps).

It mirrors the device orchestration step for step (two residual copies, one pending rank-one record per copy,
packed gather records, refresh at the start of every sweep) so that the *scheme* can be checked against the
reference iteration (nmf.py:415-476 with the masked branches :687-701, :735-746) on the CPU, where there is no
GPU to run the kernels.  The product never imports this file.
"""
import numpy as np

EPS = float(np.spacing(10))          # optimization.py:5


def solve_vector_c(numer, denom, eps, ub):
    """optimization.py:75-84 (no sum constraint)"""
    x = np.zeros_like(numer)
    pos = denom > 0
    x[pos] = np.maximum(numer[pos], 0) / (denom[pos] + eps)
    if ub is None and np.any(denom < 0):
        raise ValueError('Minimum objective is unbounded.')
    if ub is not None:
        x = np.minimum(x, ub)
    return x


class Side(object):
    """one orientation: segments own contiguous entry ranges"""

    def __init__(self, seg, idx, x, wgt, nseg):
        order = np.lexsort((idx, seg))
        self.seg, self.idx = seg[order], idx[order]
        self.x = x[order]
        self.wgt = None if wgt is None else wgt[order]
        self.nseg = nseg
        self.E = np.zeros_like(self.x)

    def residual(self, A, B):
        """E = x - sum_l A[seg,l] B[idx,l]   (sp_residual_kernel)"""
        self.E = self.x - np.einsum('el,el->e', A[self.seg], B[self.idx])

    def half_step(self, pend_other, pend_own, oth_old, oth_new, own_cur):
        """sp_pass_kernel: apply the pending record, return (numer, denom) per segment"""
        if pend_own is not None:
            opo, opn = pend_own
            qa, qb = pend_other
            self.E = self.E + (qa[self.idx] * opo[self.seg] - qb[self.idx] * opn[self.seg])
        eh = self.E + own_cur[self.seg] * oth_old[self.idx]
        wn = oth_new[self.idx] if self.wgt is None else self.wgt * oth_new[self.idx]
        numer = np.bincount(self.seg, weights=wn * eh, minlength=self.nseg)
        denom = np.bincount(self.seg, weights=wn * oth_new[self.idx], minlength=self.nseg)
        return numer, denom


class SparseWRRI(object):
    def __init__(self, rows, cols, vals, n, d, weights=None, allreduce=None):
        """allreduce: optional callable summing a vector over the row shards (the multi-GPU exchange of sp_T_step:
        [numer(d) | denom(d)] of every T-step; W-steps are shard-local)"""
        self.n, self.d = n, d
        self.allreduce = allreduce
        self.csr = Side(rows, cols, vals, weights, n)
        self.csc = Side(cols, rows, vals, weights, d)

    def _T_step(self, W, T, t, S, reg_l1, reg_l2, ub):
        wt = W[:, t].copy()
        pend = S['csc']
        told = T[t].copy()
        numer, denom = self.csc.half_step(None if pend is None else (pend[0], pend[1]),
                                          None if pend is None else (pend[2], pend[3]), wt, wt, told)
        if self.allreduce is not None:
            both = self.allreduce(np.concatenate([numer, denom]))
            numer, denom = both[:self.d], both[self.d:]
        T[t] = solve_vector_c(numer - reg_l1, denom + reg_l2, EPS, ub)
        S['csc'] = (wt, wt, told, T[t].copy())
        S['told_cur'] = told

    def _W_step(self, W, T, t, S, reg_l1, reg_l2, ub):
        trow = T[t].copy()
        told = S['told_cur'] if S['told_cur'] is not None else trow
        pend = S['csr']
        wold = W[:, t].copy()
        numer, denom = self.csr.half_step(None if pend is None else (pend[2], pend[3]),
                                          None if pend is None else (pend[0], pend[1]), told, trow, wold)
        W[:, t] = solve_vector_c(numer - reg_l1, denom + reg_l2, EPS, ub)
        S['csr'] = (wold, W[:, t].copy(), told, trow)
        if S['told_cur'] is not None:
            S['csc'] = S['csr']
        S['told_cur'] = None

    def sweeps(self, W, T, n_sweeps, order='rri', fix_T=False, reg_w_l1=0.0, reg_w_l2=0.0, reg_t_l1=0.0,
               reg_t_l2=0.0, ub_w=None, ub_t=None, check_copies=False):
        k = W.shape[1]
        new = lambda: {'csr': None, 'csc': None, 'told_cur': None}
        for _ in range(n_sweeps):
            if not fix_T and order == 'rri':
                S = new()
                self.csc.residual(T.T.copy(), W)
                self.csr.residual(W, T.T.copy())
                for t in range(k):
                    self._T_step(W, T, t, S, reg_t_l1, reg_t_l2, ub_t)
                    self._W_step(W, T, t, S, reg_w_l1, reg_w_l2, ub_w)
                    if check_copies:
                        self._assert_copies_agree(S)
                continue
            if not fix_T:
                S = new()
                self.csc.residual(T.T.copy(), W)
                for t in range(k):
                    self._T_step(W, T, t, S, reg_t_l1, reg_t_l2, ub_t)
                    S['told_cur'] = None
            S = new()
            self.csr.residual(W, T.T.copy())
            for t in range(k):
                self._W_step(W, T, t, S, reg_w_l1, reg_w_l2, ub_w)
        return W, T

    def _assert_copies_agree(self, S):
        """after a full topic both copies hold the same values and owe the same pending record"""
        a = np.zeros((self.n, self.d))
        b = np.zeros((self.n, self.d))
        a[self.csr.seg, self.csr.idx] = self.csr.E
        b[self.csc.idx, self.csc.seg] = self.csc.E
        # the CSR copy absorbed the previous topic's record during this topic's W-step; the CSC copy during this
        # topic's T-step: identical arithmetic per entry -> identical bits
        assert np.array_equal(a, b), 'the two residual copies diverged'
        assert S['csr'] is S['csc']

    def objective_terms(self, W, T):
        self.csr.residual(W, T.T.copy())
        m = 1.0 if self.csr.wgt is None else self.csr.wgt
        return 0.5 * float(np.sum(m * self.csr.E ** 2)), float(np.sum(m * self.csr.x ** 2))
