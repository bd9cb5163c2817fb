"""Row-sharded multi-GPU parity (run under torch.distributed.run, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py                     (N = 2, 4 or 8: the world size is whatever was launched)

Every rank builds the same seeded problem, keeps its row shard, runs the device sweep with the engine's NCCL
all-reduce, and rank 0 compares the gathered factors with the unsharded NumPy oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import rri_oracle as orc                      # noqa: E402
import rri_nmf_b200 as R                      # noqa: E402
from rri_nmf_b200.engine import NcclComm      # noqa: E402
from rri_nmf_b200.sharding import shard_bounds  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = NcclComm(local)
    ok = True

    def gather_rows(Wl, n):
        parts = [None] * world
        dist.all_gather_object(parts, Wl)
        return np.vstack(parts)

    n, d, k = 1003, 517, 9
    X, W0, T0, M = orc.synth(n, d, k, k, sigma=0.05, seed=3, mask_density=0.3)
    b, e = shard_bounds(n, world)[rank]
    cases = [('rri', None, np.float64, 1e-9), ('hals', None, np.float64, 1e-9), ('rri', M, np.float64, 1e-9),
             ('hals', M, np.float64, 1e-9)]
    for order, Mm, dt, tol in cases:
        o = orc.nmf_oracle(X, k, W0, T0, max_iter=4, order=order, W_mat=Mm, compute_obj_each_iter=True, eps_stop=-1.0)
        g = R.nmf(X[b:e], k, W_in=W0[b:e], T_in=T0, max_iter=4, update_order=order, reset_topic_method=None,
                  W_mat=None if Mm is None else Mm[b:e], comm=comm, device='cuda:%d' % local,
                  compute_obj_each_iter=True, eps_stop=-1.0)
        Wg = gather_rows(g['W'], n)
        rw = np.linalg.norm(Wg - o['W']) / np.linalg.norm(o['W'])
        rt = np.linalg.norm(g['T'] - o['T']) / np.linalg.norm(o['T'])
        ro = np.max(np.abs(np.array(g['obj_history']) / np.array(o['obj_history']) - 1))
        good = rw < tol and rt < tol and ro < 1e-9
        ok = ok and good
        if rank == 0:
            print('%-5s masked=%-5s relW=%.2e relT=%.2e relObj=%.2e %s' % (order, Mm is not None, rw, rt, ro,
                                                                            'ok' if good else 'FAIL'), flush=True)
    def replicas_identical(Tnp):
        """T is replicated: every rank must hold the same bits as rank 0"""
        tt = torch.from_numpy(np.ascontiguousarray(Tnp)).cuda()
        t0 = tt.clone()
        dist.broadcast(t0, src=0)
        flag = torch.tensor([1 if torch.equal(t0, tt) else 0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    same = replicas_identical(g['T'])
    ok = ok and same
    # observed-entries (sparse CSR) path on row shards: the T-step statistic [numer | denom] is all-reduced
    # (api.cu sp_T_step), W-steps are local; fp64 against the unsharded oracle on the same masked problem
    import scipy.sparse as sp
    Xs = sp.csr_matrix(X * M)
    for order in ('rri', 'hals'):
        o = orc.nmf_oracle(X, k, W0, T0, max_iter=3, order=order, W_mat=M, t_row_sum=1.0)
        g = R.nmf(Xs[b:e], k, W_in=W0[b:e], T_in=T0, max_iter=3, update_order=order, reset_topic_method=None,
                  t_row_sum=1.0, comm=comm, device='cuda:%d' % local)
        Wg = gather_rows(g['W'], n)
        rw = np.linalg.norm(Wg - o['W']) / np.linalg.norm(o['W'])
        rt = np.linalg.norm(g['T'] - o['T']) / np.linalg.norm(o['T'])
        sm = replicas_identical(g['T'])
        good = rw < 1e-9 and rt < 1e-9 and sm
        ok = ok and good
        if rank == 0:
            print('sparse %-5s relW=%.2e relT=%.2e replicas identical: %s %s' % (order, rw, rt, sm, 'ok' if good else 'FAIL'),
                  flush=True)
    # dense masked fp32 with the tensor-core statistics (math='tf32') on row shards: relative error within 1e-4
    n3, d3, k3 = 1536, 1024, 20
    X3, W3, T3, M3 = orc.synth(n3, d3, k3, k3, sigma=0.05, seed=9, mask_density=0.2)
    o = orc.nmf_oracle(X3, k3, W3, T3, max_iter=2, W_mat=M3, t_row_sum=1.0)
    b3, e3 = shard_bounds(n3, world)[rank]
    g = R.nmf(X3[b3:e3].astype(np.float32), k3, W_in=W3[b3:e3].astype(np.float32), T_in=T3.astype(np.float32), max_iter=2,
              W_mat=torch.from_numpy(M3[b3:e3].astype(np.uint8)), t_row_sum=1.0, math='tf32', reset_topic_method=None,
              comm=comm, device='cuda:%d' % local)
    Wg = gather_rows(g['W'], n3).astype(np.float64)
    re_o = orc.rel_error(X3, o['W'], o['T'], M3)
    re_g = orc.rel_error(X3, Wg, g['T'].astype(np.float64), M3)
    sm = replicas_identical(g['T'])
    good = abs(re_o - re_g) < 1e-4 and sm
    ok = ok and good
    if rank == 0:
        print('masked tf32 relerr oracle=%.6f gpu=%.6f replicas identical: %s %s' % (re_o, re_g, sm, 'ok' if good else 'FAIL'),
              flush=True)
    # fp32 / tf32 block order: relative error within 1e-4 of the oracle
    n2, d2, k2 = 4096, 2048, 32
    X2, W2, T2 = orc.synth(n2, d2, k2, k2, sigma=0.05, seed=5)
    o = orc.nmf_oracle(X2, k2, W2, T2, max_iter=3, order='hals')
    b2, e2 = shard_bounds(n2, world)[rank]
    g = R.nmf(X2[b2:e2].astype(np.float32), k2, W_in=W2[b2:e2].astype(np.float32), T_in=T2.astype(np.float32),
              max_iter=3, update_order='hals', math='tf32', reset_topic_method=None, comm=comm,
              device='cuda:%d' % local)
    Wg = gather_rows(g['W'], n2).astype(np.float64)
    re_o = orc.rel_error(X2, o['W'], o['T'])
    re_g = orc.rel_error(X2, Wg, g['T'].astype(np.float64))
    same = same and replicas_identical(g['T'])
    good = abs(re_o - re_g) < 1e-4 and same
    ok = ok and good
    if rank == 0:
        print('tf32 hals relerr oracle=%.6f gpu=%.6f %s ; T replicas identical: %s' % (re_o, re_g, 'ok' if good else 'FAIL', same),
              flush=True)
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm.destroy()
    dist.destroy_process_group()
    if rank == 0:
        print('MULTI_GPU_PARITY %s (world=%d)' % ('PASS' if int(flag.item()) else 'FAIL', world), flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == '__main__':
    main()
