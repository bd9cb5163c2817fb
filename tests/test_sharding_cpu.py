"""World-size-2 (gloo, CPU) check of the row-sharded algorithm the multi-GPU path implements:
every rank holds X_i, W_i and a replica of T; per topic the only exchange is an all-reduce of the shard
statistic [w_t'X_i (d) | w_t'W_i (k)] (reference nmf.py:680-686); the W-step is local.  Here the oracle
plays the role of the per-rank kernels, so this pins the *host-side* protocol (what is reduced, in what
order) against the unsharded oracle.  The observed-entries (sparse) path is covered the same way through its
NumPy scheme model (oracle/sparse_scheme.py): the exchange is [numer(d) | denom(d)] per T-step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rri_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_sweep_rri(Xi, Wi, T, regs=None):
    """interleaved order with one all-reduce of (d + k) numbers per topic"""
    k, d = T.shape
    for t in range(k):
        w = Wi[:, t]
        stat = np.concatenate([w.dot(Xi), w.dot(Wi)])
        ts = torch.from_numpy(stat)
        dist.all_reduce(ts)
        p, g = ts.numpy()[:d], ts.numpy()[d:].copy()
        nw = g[t]
        g[t] = 0
        T[t, :] = np.maximum(p - g.dot(T), 0) / (nw + orc.EPS_DIV_BY_ZERO)
        orc.step_W(Xi, Wi, T, t)                       # row-local


def _sharded_sweep_hals(Xi, Wi, T):
    """block order: ONE all-reduce of [W_i'X_i (k x d) | W_i'W_i (k x k)] per sweep"""
    k, d = T.shape
    stat = torch.from_numpy(np.concatenate([Wi.T.dot(Xi).ravel(), Wi.T.dot(Wi).ravel()]))
    dist.all_reduce(stat)
    P = stat.numpy()[:k * d].reshape(k, d)
    G = stat.numpy()[k * d:].reshape(k, k)
    for t in range(k):
        g = G[t].copy()
        g[t] = 0
        T[t, :] = np.maximum(P[t] - g.dot(T), 0) / (G[t, t] + orc.EPS_DIV_BY_ZERO)
    for t in range(k):
        orc.step_W(Xi, Wi, T, t)


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    X, W0, T0 = orc.synth(101, 60, 5, 6, sigma=0.05, seed=4)
    bounds = np.linspace(0, 101, world + 1).astype(int)
    sl = slice(bounds[rank], bounds[rank + 1])
    out = {}
    for order, fn in (('rri', _sharded_sweep_rri), ('hals', _sharded_sweep_hals)):
        Wi, T = W0[sl].copy(), T0.copy()
        for _ in range(4):
            fn(Xi=X[sl], Wi=Wi, T=T)
        gathered = [None] * world
        dist.all_gather_object(gathered, Wi)
        out[order] = (np.vstack(gathered), T)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_row_sharded_protocol_world2():
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    X, W0, T0 = orc.synth(101, 60, 5, 6, sigma=0.05, seed=4)
    for order in ('rri', 'hals'):
        ref = orc.nmf_oracle(X, 6, W0, T0, max_iter=4, order=order)
        W, T = out[order]
        assert np.linalg.norm(W - ref['W']) / np.linalg.norm(ref['W']) < 1e-12
        assert np.linalg.norm(T - ref['T']) / np.linalg.norm(ref['T']) < 1e-12


def _sparse_worker(rank, world, port, q):
    """observed-entries path, row-sharded: every rank owns the entries of its rows (both orientations of them);
    the only exchange is the all-reduce of [numer(d) | denom(d)] in each T-step (api.cu: sp_T_step)"""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sparse_scheme import SparseWRRI
    X, W0, T0, M = orc.synth(90, 70, 5, 6, sigma=0.05, seed=6, mask_density=0.3)
    bounds = np.linspace(0, 90, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]
    I, J = M[lo:hi].nonzero()

    def allreduce(v):
        t = torch.from_numpy(np.ascontiguousarray(v))
        dist.all_reduce(t)
        return t.numpy()

    out = {}
    for order in ('rri', 'hals'):
        S = SparseWRRI(I, J, X[lo:hi][I, J], hi - lo, 70, allreduce=allreduce)
        Wi, T = np.maximum(W0[lo:hi], 0).copy(), np.maximum(T0, 0).copy()
        S.sweeps(Wi, T, 3, order=order, ub_t=1.0, reg_t_l1=0.01)
        gathered = [None] * world
        dist.all_gather_object(gathered, (Wi, T))
        out[order] = gathered
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_row_sharded_sparse_protocol_world2():
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    X, W0, T0, M = orc.synth(90, 70, 5, 6, sigma=0.05, seed=6, mask_density=0.3)
    for order in ('rri', 'hals'):
        ref = orc.nmf_oracle(X, 6, W0, T0, max_iter=3, W_mat=M, order=order, t_row_sum=1.0, reg_t_l1=0.01)
        (Wa, Ta), (Wb, Tb) = out[order]
        assert np.array_equal(Ta, Tb)                       # replicas of T stay identical
        W = np.vstack([Wa, Wb])
        assert np.linalg.norm(W - ref['W']) / np.linalg.norm(ref['W']) < 1e-12
        assert np.linalg.norm(Ta - ref['T']) / np.linalg.norm(ref['T']) < 1e-12


def test_shard_bounds_helper():
    from rri_nmf_b200.sharding import shard_bounds
    assert shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    b = shard_bounds(200000, 8)
    assert b[0] == (0, 25000) and b[-1] == (175000, 200000)


def test_row_sharded_call_requires_both_factors():
    """an initialisation per shard would give every rank its own T replica; nmf() refuses before touching
    the device (the check is host logic, so it runs here without a GPU)"""
    import rri_nmf_b200 as R

    class FakeComm(object):
        world, rank = 2, 0

    X = np.random.RandomState(0).rand(20, 12)
    with pytest.raises(ValueError, match='row-sharded'):
        R.nmf(X, 3, comm=FakeComm())
    with pytest.raises(ValueError, match='row-sharded'):
        R.nmf(X, 3, comm=FakeComm(), T_in=np.ones((3, 12)))
